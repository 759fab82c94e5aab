"""GPU parity tests proper: libnanowrap.so (through the C ABI via the ctypes shim) against the CPU
oracle on the same seeded inputs.  Tolerances are stated per check."""
import copy

import numpy as np
import pytest

from conftest import make_case

pytestmark = pytest.mark.gpu


def _clone(mesh):
    return copy.deepcopy(mesh)


def _oracle(mesh, pts):
    from oracle import nanowrap_oracle as orc
    return orc.OracleConjGrad(mesh, pts)


def _gpu(mesh, pts):
    from ch_shrinkwrap_b200.mesh_conj_grad import ShrinkwrapMeshConjGrad
    cg = ShrinkwrapMeshConjGrad(mesh, pts)
    mesh.cg = cg
    return cg


def _weights_parity(mesh, pts):
    mo, mg = _clone(mesh), _clone(mesh)
    oc = _oracle(mo, pts)
    oc.f = oc.vertices.copy().ravel()
    v_idx_o, w_o = oc.compute_weights(oc.f)
    g = _gpu(mg, pts)
    v_idx_g, w_g = g.compute_weights()
    face_g = g.nearest_face
    # nearest-face index: bit-exact (fp64 ties -> any index at the same fp64 distance is accepted)
    diff = np.flatnonzero(face_g != oc.nearest)
    if len(diff):
        fv = oc.f.reshape(-1, 3)
        cen = fv[mo.faces].mean(1).astype(np.float64)
        p64 = pts.astype(np.float64)
        d_g = ((p64[diff] - cen[face_g[diff]]) ** 2)
        d_o = ((p64[diff] - cen[oc.nearest[diff]]) ** 2)
        d_g = (d_g[:, 0] + d_g[:, 1]) + d_g[:, 2]
        d_o = (d_o[:, 0] + d_o[:, 1]) + d_o[:, 2]
        assert np.array_equal(d_g, d_o), 'nearest-face mismatch that is not an exact fp64 tie'
    ties = len(diff)
    same = face_g == oc.nearest
    assert np.array_equal(v_idx_g[same], v_idx_o[same])
    assert np.array_equal(w_g[same], w_o[same]), 'weights must be bit-identical'      # tolerance: 0
    assert np.array_equal(g.d[:, 0][same], oc.d[:, 0][same]), 'distances must be bit-identical (fp64)'
    return ties


def test_nearest_face_and_weights_f32():
    mesh, pts, sig = make_case(n_points=20000, n_geo=10, seed=11)
    assert _weights_parity(mesh, pts) == 0


def test_nearest_face_and_weights_f64_points():
    mesh, pts, sig = make_case(n_points=6000, n_geo=7, seed=12, dtype=np.float64)
    assert pts.dtype == np.float64
    assert _weights_parity(mesh, pts) == 0


def test_nearest_face_far_and_inside_points():
    # points far outside, at the centre (worst case for pruning) and on vertices
    mesh, pts, sig = make_case(n_points=3000, n_geo=5, seed=13)
    extra = np.array([[0, 0, 0], [1e4, -2e4, 3e4], [-1e5, 0, 0]], np.float32)
    pts = np.concatenate([pts, extra, mesh.vertices[:50].copy()], 0)
    _weights_parity(mesh, pts)


def test_nearest_face_exact_after_the_mesh_moved_without_a_new_upload():
    """The octree cells are keyed once per topology upload; the cell-clearance early-out of the search has to stay exact
    when the vertices have since moved by many cell widths (its slack is re-measured at every refit)."""
    mesh, pts, sig = make_case(n_points=20000, n_geo=9, seed=41)
    s = (1.0 / sig.ravel()).astype(np.float32)
    g = _gpu(mesh, pts)
    g.search(pts, lams=[5.0], num_iters=6, sigma_inv=s)           # moves every vertex; no re-upload afterwards
    moved = np.abs(mesh._vertices['position'] - make_case(n_points=10, n_geo=9, seed=41)[0]._vertices['position']).max()
    assert moved > 1.0
    # shift a patch of vertices by hand as well: a large, local displacement
    mesh._vertices['position'][:40] += np.float32(15.0)
    g.compute_weights()                                            # nw_set_positions + refit on the OLD tables
    face_g, d_g = g.nearest_face.copy(), g.d[:, 0].copy()
    mo = _clone(mesh)
    oc = _oracle(mo, pts)
    oc.f = oc.vertices.copy().ravel()
    oc.compute_weights(oc.f)
    assert np.array_equal(d_g, oc.d[:, 0]), 'nearest distances must be bit-identical (fp64)'
    diff = np.flatnonzero(face_g != oc.nearest)                    # only exact fp64 ties may differ, and they resolve to the lowest index
    assert len(diff) == 0


def test_forward_and_adjoint():
    mesh, pts, sig = make_case(n_points=8000, n_geo=6, seed=14)
    mo, mg = _clone(mesh), _clone(mesh)
    oc = _oracle(mo, pts)
    oc.f = oc.vertices.copy().ravel()
    oc.w = oc.compute_weights(oc.f)
    oc._prev_loopcount = oc.loopcount
    g = _gpu(mg, pts)
    g.compute_weights()
    rng = np.random.default_rng(0)
    x = rng.standard_normal(3 * oc.M).astype(np.float32)
    r = rng.standard_normal(3 * len(pts)).astype(np.float32)
    # A: same float32 operations in the same order -> bit-identical
    assert np.array_equal(g.Afunc(x), oc.Afunc(x))
    # AH: reference sums float32 sequentially in point order; ours is an exact fixed-point sum rounded once.
    # tolerance: |diff| <= 1e-5 * sum_p |w r| per vertex component (float32 reassociation bound)
    ah_g, ah_o = g.Ahfunc(r), oc.Ahfunc(r)
    v_idx, w = oc.w
    absum = np.zeros((oc.M, 3))
    for j in range(3):
        np.add.at(absum, v_idx[:, j], np.abs(w[:, j][:, None] * r.reshape(-1, 3)))
    assert np.all(np.abs(ah_g - ah_o).reshape(-1, 3) <= 1e-5 * absum + 1e-30)
    # exact check against a float64 scatter: rounding once must be within half an ulp-ish
    ex = np.zeros((oc.M, 3))
    for j in range(3):
        np.add.at(ex, v_idx[:, j], (w[:, j][:, None] * r.reshape(-1, 3)).astype(np.float32).astype(np.float64))
    assert np.allclose(ah_g.reshape(-1, 3), ex, rtol=2e-7, atol=1e-9 * np.abs(r).max())
    # adjointness <Ax, r> == <x, AHr> (fp64 accumulate), relative 1e-5
    lhs = np.dot(g.Afunc(x).astype(np.float64), r.astype(np.float64))
    rhs = np.dot(x.astype(np.float64), ah_g.astype(np.float64))
    assert abs(lhs - rhs) <= 1e-5 * max(abs(lhs), abs(rhs), 1.0)
    # point influence, tolerance rel 1e-5
    pi_o = oc.point_influence() if hasattr(oc, 'res') else None
    oc.res = np.zeros(3 * len(pts), np.float32)
    pi_o = oc.point_influence()
    assert np.allclose(g.point_influence(), pi_o, rtol=1e-5, atol=1e-7)


@pytest.mark.parametrize('n_geo,n_points', [(1, 20000), (3, 20000), (6, 3000)])
def test_adjoint_group_shapes(n_geo, n_points):
    """The warp-level sums of the adjoint (csrc/sweep.cu: warp_adjoint_scatter) at their extremes: a coarse mesh puts all 32
    points of a warp on one or two faces (one group, a 32-term byte column), a sparse cloud gives every lane its own face
    (more than 16 groups: the second pass of the product).  Residuals are signed and span 12 orders of magnitude, so the
    bias / byte split of negative and tiny terms is exercised.  Tolerance: the result is the exact sum of the fixed-point
    terms rounded once to float32: |diff| <= 2^-23 |exact| + n_terms * 2^-37 * max|w r|  (terms are kept to 2^-38 of the
    largest possible one, k_shift_final)."""
    mesh, pts, sig = make_case(n_points=n_points, n_geo=n_geo, seed=40 + n_geo)
    mo, mg = _clone(mesh), _clone(mesh)
    oc = _oracle(mo, pts)
    oc.f = oc.vertices.copy().ravel()
    oc.w = oc.compute_weights(oc.f)
    g = _gpu(mg, pts)
    g.compute_weights()
    rng = np.random.default_rng(n_geo)
    r = (rng.standard_normal(3 * len(pts)) * 10.0 ** rng.uniform(-8, 4, 3 * len(pts))).astype(np.float32)
    ah_g = g.Ahfunc(r).reshape(-1, 3).astype(np.float64)
    v_idx, w = oc.w
    prod = [(w[:, j][:, None] * r.reshape(-1, 3)).astype(np.float32).astype(np.float64) for j in range(3)]
    ex = np.zeros((oc.M, 3))
    cnt = np.zeros(oc.M)
    for j in range(3):
        np.add.at(ex, v_idx[:, j], prod[j])
        np.add.at(cnt, v_idx[:, j], 1.0)
    tol = 2.0 ** -23 * np.abs(ex) + (cnt[:, None] + 1.0) * 2.0 ** -37 * np.abs(r).max()
    assert np.all(np.abs(ah_g - ex) <= tol)
    # bitwise repeatable, and independent of the order the points were given in (integer sums)
    assert np.array_equal(g.Ahfunc(r), g.Ahfunc(r))
    perm = rng.permutation(len(pts))
    g2 = _gpu(_clone(mesh), pts[perm])
    g2.compute_weights()
    assert np.array_equal(g2.Ahfunc(r.reshape(-1, 3)[perm].ravel()), g.Ahfunc(r))


def test_determinism_adjoint_bitwise():
    mesh, pts, sig = make_case(n_points=30000, n_geo=6, seed=15)
    g = _gpu(mesh, pts)
    g.compute_weights()
    r = np.random.default_rng(1).standard_normal(3 * len(pts)).astype(np.float32)
    a = g.Ahfunc(r)
    for _ in range(3):
        assert np.array_equal(a, g.Ahfunc(r))


def test_ncc_prior():
    mesh, pts, sig = make_case(n_points=8000, n_geo=6, seed=16)
    mo, mg = _clone(mesh), _clone(mesh)
    oc = _oracle(mo, pts)
    oc.f = oc.vertices.copy().ravel()
    oc.w = oc.compute_weights(oc.f)
    oc.res = np.zeros(3 * len(pts), np.float32)
    fd_o = oc.ncc()
    g = _gpu(mg, pts)
    g.compute_weights()
    fd_g = g._ncc()
    # tolerance: relative 1e-6 of the coordinate scale (point influence differs at float32 rounding level)
    assert np.allclose(fd_g, fd_o, rtol=0, atol=1e-6 * np.abs(fd_o).max())


@pytest.mark.parametrize('n_iters', [1, 2, 7])
def test_search_fixed_iterations(n_iters):
    mesh, pts, sig = make_case(n_points=10000, n_geo=6, seed=17)
    mo, mg = _clone(mesh), _clone(mesh)
    s = (1.0 / sig.ravel()).astype(np.float32)
    oc = _oracle(mo, pts)
    vo = oc.search(pts, lams=[10.0], num_iters=n_iters, sigma_inv=s)
    g = _gpu(mg, pts)
    vg = g.search(pts, lams=[10.0], num_iters=n_iters, sigma_inv=s)
    # final mesh: max vertex displacement (upper bound on symmetric Hausdorff) <= 0.01 nm
    disp = np.sqrt(((vg.astype(np.float64) - vo) ** 2).sum(1)).max()
    assert disp <= 1e-2, disp
    assert np.array_equal(mg._vertices['position'], vg)
    assert g.loopcount == n_iters
    assert np.allclose(np.array(g.tests, np.float64), np.array(oc.tests, np.float64), rtol=1e-3, atol=1e-5)
    assert np.allclose(g.ress, oc.ress, rtol=1e-4)
    assert np.allclose(g.S, oc.S, rtol=1e-3, atol=1e-3 * np.abs(oc.S).max())
    assert np.allclose(g.res, oc.res, rtol=1e-4, atol=1e-4 * np.abs(oc.res).max())


def test_search_scalar_sigma_and_continue():
    # scalar sigma is passed through un-inverted (_membrane_mesh.pyx:1460-1461); calling search twice continues
    mesh, pts, sig = make_case(n_points=5000, n_geo=5, seed=18)
    mo, mg = _clone(mesh), _clone(mesh)
    oc = _oracle(mo, pts)
    g = _gpu(mg, pts)
    for _ in range(2):
        vo = oc.search(pts, lams=[5.0], num_iters=3, sigma_inv=10.0)
        vg = g.search(pts, lams=[5.0], num_iters=3, sigma_inv=10.0)
    disp = np.sqrt(((vg.astype(np.float64) - vo) ** 2).sum(1)).max()
    assert disp <= 1e-2, disp
    assert len(g.tests) == 6


def test_search_with_zero_weights_mask():
    mesh, pts, sig = make_case(n_points=5000, n_geo=5, seed=19)
    mo, mg = _clone(mesh), _clone(mesh)
    s = (1.0 / sig.ravel()).astype(np.float32)
    wts = s.copy()
    wts[::7] = 0.0
    oc = _oracle(mo, pts)
    vo = oc.search(pts, lams=[5.0], num_iters=3, sigma_inv=s, weights=wts)
    g = _gpu(mg, pts)
    vg = g.search(pts, lams=[5.0], num_iters=3, sigma_inv=s, weights=wts)
    disp = np.sqrt(((vg.astype(np.float64) - vo) ** 2).sum(1)).max()
    assert disp <= 1e-2, disp


def test_deleted_vertices_stay_put():
    mesh, pts, sig = make_case(n_points=4000, n_geo=5, seed=20)
    # append two unused (deleted) vertex rows: halfedge == -1, no neighbours (mesh_conj_grad.py:44)
    from ch_shrinkwrap_b200.minimesh import VERTEX_DTYPE
    extra = np.zeros(2, VERTEX_DTYPE)
    extra['halfedge'] = -1
    extra['neighbors'] = -1
    extra['position'] = [[1., 2., 3.], [4., 5., 6.]]
    mesh._vertices = np.concatenate([mesh._vertices, extra])
    mo, mg = _clone(mesh), _clone(mesh)
    s = (1.0 / sig.ravel()).astype(np.float32)
    vo = _oracle(mo, pts).search(pts, lams=[5.0], num_iters=3, sigma_inv=s)
    vg = _gpu(mg, pts).search(pts, lams=[5.0], num_iters=3, sigma_inv=s)
    assert np.array_equal(vg[-2:], extra['position'])
    assert np.sqrt(((vg.astype(np.float64) - vo) ** 2).sum(1)).max() <= 1e-2
    # side effect of search() (:289): the library writes f into the records row by row, valid rows only
    assert np.array_equal(mg._vertices['position'][:-2], vg[:-2])
    assert np.array_equal(mg._vertices['position'][-2:], extra['position'])
    assert np.array_equal(mg._vertices['normal'], mesh._vertices['normal'])      # neighbouring fields untouched
    assert np.array_equal(mg._vertices['halfedge'], mesh._vertices['halfedge'])


def test_curvature_sphere_and_plane_known_answers():
    # the reference's own pins: tests/test_membrane_mesh.py:50,64,73,88
    from ch_shrinkwrap_b200 import minimesh
    from ch_shrinkwrap_b200.membrane_mesh import curvature_grad
    for R, n in [(37.0, 8), (100.0, 16)]:
        m = minimesh.sphere_mesh(R, n)
        c = curvature_grad(m)
        np.testing.assert_almost_equal(np.nanmean(c['H']), 1.0 / R, decimal=2)
        np.testing.assert_almost_equal(np.nanmean(c['K']), 1.0 / R ** 2, decimal=4)
    pl = minimesh.planar_mesh(17.0, 4)
    c = curvature_grad(pl)
    assert abs(np.nanmean(c['H'])) < 1e-6 and abs(np.nanmedian(c['K'])) < 1e-6


def test_curvature_vs_oracle():
    from ch_shrinkwrap_b200 import minimesh, synth
    from ch_shrinkwrap_b200.membrane_mesh import curvature_grad
    from oracle import nanowrap_oracle as orc
    shape = synth.two_lobed()
    m = synth.star_mesh(shape, 12, scale=1.0)
    nv = int((m._vertices['halfedge'] != -1).sum())
    u = np.random.default_rng(5).random(3 * nv)
    o = orc.curvature_grad(m, jitter_u=u)
    g = curvature_grad(m, jitter_u=u)
    for k in ('k0', 'k1', 'e0', 'e1', 'H', 'K', 'E', 'dE_neighbors'):
        assert np.array_equal(g[k], o[k], equal_nan=True), k          # IEEE ops only: bit-exact
    for k in ('dH', 'dK', 'pE', 'dEdN'):                                # atan2/sin/cos/exp: rel 1e-5
        assert np.allclose(g[k], o[k], rtol=1e-5, atol=1e-5 * np.nanmax(np.abs(o[k])), equal_nan=True), k


def test_ring_regularisers_bitwise():
    from oracle import nanowrap_oracle as orc
    mesh, pts, sig = make_case(n_points=100, n_geo=5, seed=21)
    g = _gpu(mesh, pts)
    nb = g.vertex_neighbors
    rng = np.random.default_rng(2)
    f = rng.standard_normal(3 * g.M).astype(np.float32)
    ref = mesh.vertices.astype(np.float32).ravel().copy()
    g.f = ref
    assert np.array_equal(g.Lfunc(f), orc.l_func(f, nb))
    assert np.array_equal(g.Lhfunc(f), orc.lh_func(f, nb))
    assert np.array_equal(g.Lfunc3(f), orc.lw_func(f, nb, ref))
    assert np.array_equal(g.Lhfunc3(f), orc.lhw_func(f, nb, ref))
    assert np.array_equal(g.wfunc(f), f * orc.vertex_area_weights(ref, nb))


def test_final_mesh_at_moderate_scale():
    """200k localisations, 9 002 vertices, 6 iterations in two blocks (topology re-uploaded, seeds from foot points).
    Tolerance: mean vertex displacement <= 0.01 nm, 99.9th percentile <= 0.2 nm, max <= 3 nm -- the reference's own
    float32-vs-float64 Gram-sum variants differ by that much."""
    from ch_shrinkwrap_b200 import synth
    shape = synth.two_lobed()
    pts, sig = synth.smlm_cloud(shape, 200000, seed=91)
    mesh = synth.star_mesh(shape, 30, scale=1.15)
    mo, mg = _clone(mesh), _clone(mesh)
    s = (1.0 / sig.ravel()).astype(np.float32)
    for blk in range(2):
        oc = _oracle(mo, pts)
        mo.cg = oc
        vo = oc.search(pts, lams=[5.0], num_iters=3, sigma_inv=s)
        mo.update_geometry()
        g = _gpu(mg, pts)
        vg = g.search(pts, lams=[5.0], num_iters=3, sigma_inv=s)
        mg.update_geometry()
    d = np.sqrt(((vg.astype(np.float64) - vo) ** 2).sum(1))
    assert d.mean() <= 1e-2 and np.percentile(d, 99.9) <= 0.2 and d.max() <= 3.0, (d.mean(), np.percentile(d, 99.9), d.max())


def test_search_with_the_wfunc_regulariser():
    """Lfuncs = Lhfuncs = ["wfunc"] (mesh_conj_grad.py:39,725-736) read from the solver object like the reference does;
    anything else the reference cannot run with is refused."""
    mesh, pts, sig = make_case(n_points=10000, n_geo=6, seed=23)
    from oracle import nanowrap_oracle as orc
    mo, m64, mg = _clone(mesh), _clone(mesh), _clone(mesh)
    s = (1.0 / sig.ravel()).astype(np.float32)
    lam = 10.0       # the area weights are ~1/sqrt(6 edge^2) ~ 1e-2: a weak prior, so the reference's float32 sgemm noise is large here
    oc = _oracle(mo, pts)
    oc.Lfuncs, oc.Lhfuncs = ["wfunc"], ["wfunc"]
    vo = oc.search(pts, lams=[lam], num_iters=5, sigma_inv=s)
    o64 = orc.OracleConjGrad64(m64, pts)                  # the same algorithm with float64 Gram sums (what the GPU forms)
    o64.Lfuncs, o64.Lhfuncs = ["wfunc"], ["wfunc"]
    v64 = o64.search(pts, lams=[lam], num_iters=5, sigma_inv=s)
    g = _gpu(mg, pts)
    g.Lfuncs, g.Lhfuncs = ["wfunc"], ["wfunc"]
    vg = g.search(pts, lams=[lam], num_iters=5, sigma_inv=s)
    d64 = np.sqrt(((vg.astype(np.float64) - v64) ** 2).sum(1)).max()
    d32 = np.sqrt(((vg.astype(np.float64) - vo) ** 2).sum(1)).max()
    noise = np.sqrt(((v64 - vo) ** 2).sum(1)).max()      # the reference's own float32-sgemm noise on this input
    assert d64 <= 1e-2, (d64, d32, noise)                 # final mesh, nm
    assert d32 <= max(1e-2, 3 * noise), (d64, d32, noise)
    assert np.allclose(np.array(g.tests, np.float64), np.array(oc.tests, np.float64), rtol=1e-3, atol=1e-5)
    assert np.allclose([float(p[0]) for p in g.prefs], [float(p[0]) for p in oc.prefs], rtol=1e-4)
    # differs from the default regulariser (the selector is live) ...
    vi = _gpu(_clone(mesh), pts).search(pts, lams=[lam], num_iters=5, sigma_inv=s)
    assert np.abs(vi - vg).max() > 1e-2
    # ... and the operators the reference itself cannot run with are refused
    g2 = _gpu(_clone(mesh), pts)
    g2.Lfuncs, g2.Lhfuncs = ["Lfunc3"], ["Lhfunc3"]
    with pytest.raises(NotImplementedError):
        g2.search(pts, lams=[10.0], num_iters=1, sigma_inv=s)
