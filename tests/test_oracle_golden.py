"""Pins the CPU oracle (oracle/) to fixtures written by the UNMODIFIED reference (oracle/make_golden.py).
Runs anywhere (no GPU, no /root/reference)."""
import hashlib
import os

import numpy as np
import pytest

from conftest import make_case

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def _digest(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return np.frombuffer(h.digest()[:8], dtype=np.uint64)[0]


def golden_case(name):
    from ch_shrinkwrap_b200 import synth
    g = np.load(os.path.join(GOLD, name + '.npz'))
    if name == 'search_sphere_f32':
        mesh, pts, sig = make_case(n_points=3000, n_geo=5, seed=101)
    else:
        mesh, pts, sig = make_case(n_points=2500, n_geo=6, seed=102, dtype=np.float64, shape=synth.two_lobed())
    assert _digest(pts, sig, mesh.faces) == g['input_digest'], 'seeded inputs changed: regenerate the fixtures'
    s = (1.0 / sig.ravel()).astype(pts.dtype) if str(g['sigma_mode']) == 'array' else 10.0
    return g, mesh, pts, s


@pytest.mark.parametrize('name', ['search_sphere_f32', 'search_lobed_f64'])
def test_oracle_search_matches_reference_fixture(name):
    from oracle import nanowrap_oracle as orc
    g, mesh, pts, s = golden_case(name)
    oc = orc.OracleConjGrad(mesh, pts)
    v = oc.search(pts, lams=[float(g['lam'])], num_iters=int(g['n_iters']), sigma_inv=s)
    # the oracle is the same arithmetic as the reference: bit-exact
    assert np.array_equal(v, g['vertices'])
    assert np.array_equal(oc.w[0], g['v_idx']) and np.array_equal(oc.w[1], g['w'])
    assert np.array_equal(oc.d[:, 0], g['d'])
    assert np.array_equal(np.asarray(oc.res), g['res'])
    assert np.array_equal(oc.S, g['S'])
    assert np.array_equal(np.array(oc.tests, np.float64), g['tests'])
    assert np.allclose(oc.ress, g['ress'], rtol=0, atol=0)
    assert float(oc.cpred) == float(g['cpred']) and float(oc.wpreds[0]) == float(g['wpred'])


def test_oracle_curvature_matches_reference_fixture():
    from ch_shrinkwrap_b200 import minimesh
    from oracle import nanowrap_oracle as orc
    g = np.load(os.path.join(GOLD, 'curvature_sphere.npz'))
    m = minimesh.sphere_mesh(float(g['radius']), int(g['n_geo']))
    o = orc.curvature_grad(m, jitter_u=g['jitter_u'])
    for k in orc.CURV_SCALARS + orc.CURV_VECTORS:
        assert np.array_equal(o[k], g[k], equal_nan=True), k
    # the reference's own pins (tests/test_membrane_mesh.py:64,88): H ~ 1/R (2 decimals), K ~ 1/R^2 (4 decimals)
    np.testing.assert_almost_equal(np.nanmean(o['H']), 1.0 / 50.0, decimal=2)
    np.testing.assert_almost_equal(np.nanmean(o['K']), 1.0 / 2500.0, decimal=4)


def test_oracle_curvature_plane_known_answer():
    # tests/test_membrane_mesh.py:50,73
    from ch_shrinkwrap_b200 import minimesh
    from oracle import nanowrap_oracle as orc
    for a, n in [(1.0, 1), (37.0, 3), (100.0, 5)]:
        o = orc.curvature_grad(minimesh.planar_mesh(a, n))
        assert abs(np.nanmean(o['H'])) < 1e-6
        assert abs(np.nanmedian(o['K'])) < 1e-6


def test_oracle_ring_ops_match_reference_fixture():
    from oracle import nanowrap_oracle as orc
    g = np.load(os.path.join(GOLD, 'ring_ops.npz'))
    mesh, pts, _ = make_case(n_points=50, n_geo=4, seed=int(g['seed']))
    nb = mesh.neighbor_vertices()
    f = g['f']
    ref = np.ascontiguousarray(mesh.vertices, dtype=np.float32).ravel()
    assert np.array_equal(orc.l_func(f, nb), g['l'])
    assert np.array_equal(orc.lh_func(f, nb), g['lh'])
    assert np.array_equal(orc.lw_func(f, nb, ref), g['lw'])
    assert np.array_equal(orc.lhw_func(f, nb, ref), g['lhw'])
    assert np.array_equal(orc.vertex_area_weights(ref, nb), g['area_w'])


def test_oracle_adjointness_and_stop_rule():
    from oracle import nanowrap_oracle as orc
    mesh, pts, sig = make_case(n_points=800, n_geo=4, seed=5)
    oc = orc.OracleConjGrad(mesh, pts)
    oc.f = oc.vertices.copy().ravel()
    oc.w = oc.compute_weights(oc.f)
    oc._prev_loopcount = oc.loopcount
    rng = np.random.default_rng(0)
    x = rng.standard_normal(3 * oc.M).astype(np.float32)
    r = rng.standard_normal(3 * len(pts)).astype(np.float32)
    lhs = np.dot(oc.Afunc(x).astype(np.float64), r)
    rhs = np.dot(x.astype(np.float64), oc.Ahfunc(r))
    assert abs(lhs - rhs) <= 1e-4 * max(abs(lhs), 1.0)
    oc.tests = [5e-7, 4e-7, 3e-7]
    assert oc._stop_cond()
    oc.tests = [5e-7, 4e-7, 4.5e-7]
    assert not oc._stop_cond()


def _quality_case():
    from ch_shrinkwrap_b200 import synth
    g = np.load(os.path.join(GOLD, 'quality.npz'))
    shape = synth.two_lobed()
    mesh = synth.star_mesh(shape, int(g['n_geo']), scale=1.0)
    pts, _ = synth.smlm_cloud(shape, int(g['cloud_n']), seed=int(g['cloud_seed']))
    return g, mesh, pts


def _lexsorted(a):
    return a[np.lexsort((a[:, 2], a[:, 1], a[:, 0]))]


def test_oracle_quality_metrics_match_reference_fixture():
    """points_from_mesh / average_squared_distance of the unmodified reference (evaluation_utils.py:35-180): bit-exact."""
    from oracle import nanowrap_oracle as orc
    g, mesh, pts = _quality_case()
    d = orc.points_from_mesh_samples(mesh, int(g['dx_min']))
    assert np.array_equal(_lexsorted(d), g['samples_sorted'])
    assert np.array_equal(np.array(orc.average_squared_distance(d, pts.astype(np.float64))), g['msd'])
    assert np.array_equal(np.array(orc.average_squared_distance(d.astype(np.float32), pts)), g['msd32'])


def test_oracle_holepunch_pairs_match_reference_fixture():
    """c_holepunch_pair_candidate_faces of the unmodified reference (membrane_mesh_utils.c:1301-1379): index-exact."""
    from oracle import nanowrap_oracle as orc
    g, mesh, pts = _quality_case()
    assert np.array_equal(orc.holepunch_pairs(mesh, g['candidates']), g['pairs'])
    assert (g['pairs'] != -1).sum() > 50
