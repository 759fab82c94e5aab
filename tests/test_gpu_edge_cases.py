"""Edge cases of the GPU path: empty and tiny inputs, ragged warps, tiny meshes, duplicates, large offsets."""
import copy

import numpy as np
import pytest

from conftest import make_case

pytestmark = pytest.mark.gpu


def _pair(mesh, pts, **kw):
    from ch_shrinkwrap_b200.mesh_conj_grad import ShrinkwrapMeshConjGrad
    from oracle import nanowrap_oracle as orc
    mo, mg = copy.deepcopy(mesh), copy.deepcopy(mesh)
    oc = orc.OracleConjGrad(mo, pts)
    cg = ShrinkwrapMeshConjGrad(mg, pts)
    mg.cg = cg
    return oc, cg


@pytest.mark.parametrize('n_points', [1, 2, 31, 32, 33, 257])
def test_ragged_point_counts(n_points):
    mesh, pts, sig = make_case(n_points=max(n_points, 40), n_geo=3, seed=70 + n_points)
    pts, sig = pts[:n_points].copy(), sig[:n_points].copy()
    oc, cg = _pair(mesh, pts)
    s = (1.0 / sig.ravel()).astype(np.float32)
    vo = oc.search(pts, lams=[5.0], num_iters=3, sigma_inv=s)
    vg = cg.search(pts, lams=[5.0], num_iters=3, sigma_inv=s)
    assert np.sqrt(((vg.astype(np.float64) - vo) ** 2).sum(1)).max() <= 1e-2
    assert np.array_equal(cg.w[0], oc.w[0]) or np.mean(np.all(cg.w[0] == oc.w[0], 1)) > 0.9


def test_zero_points_moves_only_by_the_prior():
    # with no localisations the data term vanishes; S0 = 0 makes the reference's test statistic NaN and the subspace
    # matrix singular -> numpy raises LinAlgError; here: AssertionError (non-finite / singular), never a crash
    from ch_shrinkwrap_b200.mesh_conj_grad import ShrinkwrapMeshConjGrad
    mesh, pts, sig = make_case(n_points=40, n_geo=3, seed=80)
    cg = ShrinkwrapMeshConjGrad(mesh, pts[:0].copy())
    with pytest.raises(AssertionError):
        cg.search(cg.points, lams=[5.0], num_iters=1, sigma_inv=10.0)


def test_tetrahedron_mesh_fewer_faces_than_a_leaf():
    from ch_shrinkwrap_b200.minimesh import MiniMesh
    v = np.array([[1, 1, 1], [1, -1, -1], [-1, 1, -1], [-1, -1, 1]], np.float64) * 100
    f = np.array([[0, 1, 2], [0, 3, 1], [0, 2, 3], [1, 3, 2]], np.int32)
    mesh = MiniMesh(v, f)
    rng = np.random.default_rng(3)
    pts = (rng.standard_normal((500, 3)) * 60).astype(np.float32)
    oc, cg = _pair(mesh, pts)
    oc.f = oc.vertices.copy().ravel()
    vo, wo = oc.compute_weights(oc.f)
    vg, wg = cg.compute_weights()
    assert np.array_equal(vg, vo) and np.array_equal(wg, wo)
    a = cg.search(pts, lams=[1.0], num_iters=2, sigma_inv=0.1)
    b = oc.search(pts, lams=[1.0], num_iters=2, sigma_inv=0.1)
    # 4 vertices, 3 nearly dependent directions: cond(H) ~ 1e8 in the second step, so the reference's float32 sgemm noise
    # (conj_grad.py:202,208) is amplified to ~0.03 nm.  Against the same algorithm with float64 Gram sums we agree to
    # 1e-3 nm; against the float32 reference path only to its own noise level.
    from oracle import nanowrap_oracle as orc
    o64 = orc.OracleConjGrad64(copy.deepcopy(mesh), pts)
    c = o64.search(pts, lams=[1.0], num_iters=2, sigma_inv=0.1)
    assert o64.cond > 1e6
    assert np.sqrt(((a.astype(np.float64) - c) ** 2).sum(1)).max() <= 1e-3
    assert np.sqrt(((a.astype(np.float64) - b) ** 2).sum(1)).max() <= 0.1


def test_duplicate_points_and_points_on_vertices():
    mesh, pts, sig = make_case(n_points=300, n_geo=4, seed=81)
    pts = np.concatenate([pts, pts[:100], mesh.vertices[:60].astype(np.float32)], 0)   # duplicates; d == 0 -> clamp 1e-6 (:503)
    oc, cg = _pair(mesh, pts)
    oc.f = oc.vertices.copy().ravel()
    vo, wo = oc.compute_weights(oc.f)
    vg, wg = cg.compute_weights()
    same = np.all(vg == vo, 1)
    assert same.mean() > 0.99            # a point exactly on a shared vertex is an exact fp64 tie between its faces
    assert np.array_equal(wg[same], wo[same])
    assert np.isfinite(wg).all()


def test_large_coordinate_offset():
    # localisations far from the origin (typical PYME coordinates are 1e4-1e5 nm): exactness of the search must hold
    mesh, pts, sig = make_case(n_points=3000, n_geo=5, seed=82)
    off = np.array([61234.5, -40321.25, 1500.0], np.float32)
    pts = pts + off
    mesh._vertices['position'] += off
    oc, cg = _pair(mesh, pts)
    oc.f = oc.vertices.copy().ravel()
    vo, wo = oc.compute_weights(oc.f)
    vg, wg = cg.compute_weights()
    assert np.array_equal(vg, vo) and np.array_equal(wg, wo)
    assert np.array_equal(cg.d, oc.d)


def test_open_mesh_with_boundary():
    from ch_shrinkwrap_b200 import minimesh
    mesh = minimesh.planar_mesh(200.0, 12)
    rng = np.random.default_rng(5)
    pts = np.stack([rng.uniform(0, 200, 4000), rng.uniform(0, 200, 4000), rng.normal(15, 5, 4000)], 1).astype(np.float32)
    oc, cg = _pair(mesh, pts)
    vo = oc.search(pts, lams=[2.0], num_iters=4, sigma_inv=0.2)
    vg = cg.search(pts, lams=[2.0], num_iters=4, sigma_inv=0.2)
    assert np.sqrt(((vg.astype(np.float64) - vo) ** 2).sum(1)).max() <= 1e-2


def test_repeated_topology_uploads_reuse_seeds_correctly():
    # second block on a DIFFERENT mesh: seeds come from the previous block's foot points and must not affect exactness
    from ch_shrinkwrap_b200 import synth
    from ch_shrinkwrap_b200.mesh_conj_grad import ShrinkwrapMeshConjGrad
    from oracle import nanowrap_oracle as orc
    shape = synth.Sphere(500.0)
    pts, sig = synth.smlm_cloud(shape, 20000, seed=83)
    m1 = synth.star_mesh(shape, 6, scale=1.15)
    m2 = synth.star_mesh(shape, 9, scale=1.05)
    cg1 = ShrinkwrapMeshConjGrad(m1, pts)
    cg1.search(pts, lams=[5.0], num_iters=2, sigma_inv=0.1)
    m2_host = copy.deepcopy(m2)
    m2._nw_session = m1._nw_session                 # same device session, new topology (what a remesh does)
    cg2 = ShrinkwrapMeshConjGrad(m2, pts)
    vg, wg = cg2.compute_weights()
    oc = orc.OracleConjGrad(m2_host, pts)
    oc.f = oc.vertices.copy().ravel()
    vo, wo = oc.compute_weights(oc.f)
    assert np.array_equal(vg, vo) and np.array_equal(wg, wo)


def test_large_mesh_uploads_and_write_back_take_the_threaded_paths():
    """> 2 MB uploads go through the pinned multi-lane uploader (the half-edge 'vertex' field gathered out of its 28-byte
    records) and >= 65 536 vertices are written back into the 120-byte records by several threads; the result must be
    what the packed single-threaded paths give: positions round-trip bit-exactly, deleted rows untouched."""
    import ctypes
    from ch_shrinkwrap_b200 import synth, _lib
    from ch_shrinkwrap_b200.mesh_conj_grad import ShrinkwrapMeshConjGrad
    mesh = synth.star_mesh(synth.Sphere(300.0), 82, scale=1.0)          # 10*82^2+2 = 67 242 vertices
    assert len(mesh._vertices) >= 65536 and mesh._halfedges['vertex'].strides[0] == 28
    rng = np.random.default_rng(5)
    pts = (mesh.vertices[::7] * np.float32(0.98) + rng.normal(0, 2.0, (len(mesh.vertices[::7]), 3))).astype(np.float32)   # jitter: no exact NN ties
    cg = ShrinkwrapMeshConjGrad(mesh, pts)
    cg._ensure_ready()                                                  # nw_set_points + nw_set_topology_records (strided half-edges)
    h = cg._h
    M = len(mesh._vertices)
    # neighbour table as seen by the device == the host lookup of mesh_conj_grad.py:50-52
    ref_nb = cg.vertex_neighbors
    got = np.empty((M, 3), np.float32)
    h.call('nw_get_positions', _lib.fptr(got))
    assert np.array_equal(got, mesh._vertices['position'])
    # strided write-back: mark two rows deleted on the device by re-uploading, then scatter new positions
    new_pos = (mesh._vertices['position'] * np.float32(1.5)).astype(np.float32)
    h.call('nw_set_positions', _lib.fptr(np.ascontiguousarray(new_pos)))
    rec = mesh._vertices.copy()
    h.call('nw_get_positions_strided', ctypes.c_void_p(rec['position'].ctypes.data), int(rec['position'].strides[0]), 1)
    assert np.array_equal(rec['position'], new_pos)
    for name in rec.dtype.names:
        if name != 'position':
            assert np.array_equal(rec[name], mesh._vertices[name]), name
    # the curvature prior reads the device's neighbour table (built from the gathered half-edge field): compare it with
    # the oracle's, which uses the host lookup
    import copy
    from oracle import nanowrap_oracle as orc
    h.call('nw_set_positions', _lib.fptr(np.ascontiguousarray(mesh._vertices['position'])))
    oc = orc.OracleConjGrad(copy.deepcopy(mesh), pts)
    oc.f = oc.vertices.copy().ravel()
    oc.w = oc.compute_weights(oc.f)
    oc.res = np.zeros(3 * len(pts), np.float32)
    fd_o = oc.ncc()
    cg.compute_weights()
    fd_g = cg._ncc()
    assert ref_nb.shape == (M, 20)
    assert np.allclose(fd_g, fd_o, rtol=0, atol=1e-6 * np.abs(fd_o).max()), np.abs(fd_g - fd_o).max()
