"""SURVEY 8(f) rows 2 and 4 on the GPU against the oracle and the reference's fixtures: mesh sampling, nearest-point
queries against raw point sets (average_squared_distance, hole-punch candidate search), candidate pairing."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def _lexsorted(a):
    return a[np.lexsort((a[:, 2], a[:, 1], a[:, 0]))]


def _quality_case():
    from ch_shrinkwrap_b200 import synth
    g = np.load(os.path.join(GOLD, 'quality.npz'))
    shape = synth.two_lobed()
    mesh = synth.star_mesh(shape, int(g['n_geo']), scale=1.0)
    pts, _ = synth.smlm_cloud(shape, int(g['cloud_n']), seed=int(g['cloud_seed']))
    return g, mesh, pts


def test_points_from_mesh_bitwise_vs_reference_fixture_and_oracle():
    from ch_shrinkwrap_b200 import evaluation_utils as ev, synth
    from oracle import nanowrap_oracle as orc
    g, mesh, pts = _quality_case()
    d = ev.mesh_samples(mesh, int(g['dx_min']))
    assert d.dtype == np.float64
    assert np.array_equal(_lexsorted(d), g['samples_sorted'])            # the reference's samples, bit for bit
    # generation order too (face by face, y outer / x inner), on a finer mesh and a non-representable spacing
    m2 = synth.star_mesh(synth.two_lobed(), 9, scale=1.1)
    for dx in (5, 7.3, 0.9 * 11):
        assert np.array_equal(ev.mesh_samples(m2, dx), orc.points_from_mesh_samples(m2, dx)), dx
    # public function: same permutation as the reference for the same numpy seed (np.random.choice, :137)
    np.random.seed(11)
    out = ev.points_from_mesh(m2, dx_min=5)
    np.random.seed(11)
    ref = orc.points_from_mesh_samples(m2, 5)
    assert np.array_equal(out, ref[np.random.choice(np.arange(len(ref)), size=len(ref), replace=False)])


def test_points_from_mesh_degenerate_and_tiny_triangles():
    from ch_shrinkwrap_b200 import evaluation_utils as ev, minimesh
    from oracle import nanowrap_oracle as orc
    m = minimesh.sphere_mesh(40.0, 2)
    m._vertices['position'][3] = m._vertices['position'][4]              # zero-area triangles (dropped, :66-72)
    m._vertices['position'][10] *= np.float32(1.0000001)
    assert np.array_equal(ev.mesh_samples(m, 5), orc.points_from_mesh_samples(m, 5))
    assert ev.mesh_samples(m, 1e4).shape[1] == 3                          # spacing far larger than the mesh: few or no samples
    assert np.array_equal(ev.mesh_samples(m, 1e4), orc.points_from_mesh_samples(m, 1e4))
    out, nrm = ev.points_from_mesh(m, dx_min=5, p=0.5, return_normals=True)
    assert out.shape == nrm.shape and len(out) == int(0.5 * len(ev.mesh_samples(m, 5)))


@pytest.mark.parametrize('dt0, dt1', [(np.float64, np.float64), (np.float32, np.float32), (np.float64, np.float32)])
def test_average_squared_distance_bitwise(dt0, dt1):
    """Both directions, float32 and float64 clouds: the per-point nearest distances are the float64 numbers cKDTree
    returns, so the two means are bit-identical (evaluation_utils.py:172-180)."""
    from ch_shrinkwrap_b200 import evaluation_utils as ev
    from oracle import nanowrap_oracle as orc
    rng = np.random.default_rng(7)
    a = (rng.standard_normal((30000, 3)) * 300).astype(dt0)
    b = (a[:20000].astype(np.float64) + rng.standard_normal((20000, 3)) * 2).astype(dt1)
    b[:5] = a[:5]                                                         # exact hits: distance 0
    assert ev.average_squared_distance(a, b) == orc.average_squared_distance(a, b)
    dist, idx = ev.nearest_point_distance(a, b, return_index=True)
    import scipy.spatial
    dd, ii = scipy.spatial.cKDTree(a).query(b, k=1)
    assert np.array_equal(dist, dd)
    assert np.array_equal(idx, ii) or np.all(np.linalg.norm(a[idx].astype(np.float64) - b, axis=1) == dd)


def test_average_squared_distance_vs_reference_fixture():
    from ch_shrinkwrap_b200 import evaluation_utils as ev
    g, mesh, pts = _quality_case()
    d = ev.mesh_samples(mesh, int(g['dx_min']))
    d = d[np.lexsort((d[:, 2], d[:, 1], d[:, 0]))]
    assert np.array_equal(d, g['samples_sorted'])
    assert np.array_equal(np.array(ev.average_squared_distance(d, pts.astype(np.float64))), g['msd'])
    assert np.array_equal(np.array(ev.average_squared_distance(d.astype(np.float32), pts)), g['msd32'])


def test_holepunch_searches():
    from ch_shrinkwrap_b200 import synth
    from ch_shrinkwrap_b200.membrane_mesh import MembraneMesh
    from oracle import nanowrap_oracle as orc
    g, base, pts = _quality_case()
    mesh = MembraneMesh(mesh=base)
    # pairing: the reference's fixture (every second face as candidate), index-exact
    cand = g['candidates']
    c, pi = mesh._holepunch_pair_candidate_faces(cand)
    pairs = g['pairs']
    keep = pairs != -1
    assert np.array_equal(c, cand[keep]) and np.array_equal(pi, (np.cumsum(keep) - 1)[pairs[keep]])
    # all faces of a finer mesh against the oracle (raw pairs)
    m2 = MembraneMesh(mesh=synth.star_mesh(synth.two_lobed(), 12, scale=1.0))
    allf = np.arange(len(m2._faces), dtype=np.int32)
    co, po = orc.holepunch_pair_candidate_faces(m2, allf)
    cg, pg = m2._holepunch_pair_candidate_faces(allf)
    assert np.array_equal(cg, co) and np.array_equal(pg, po) and len(cg) > 1000
    assert len(m2._holepunch_pair_candidate_faces(allf[:0])[0]) == 0
    # candidate search: faces with no localisation within eps of their centre (_membrane_mesh.pyx:877-887)
    cloud = pts[pts[:, 0] > 0]                                             # one lobe has no points at all
    for eps in (10.0, 60.0):
        assert np.array_equal(m2._holepunch_find_candidate_faces(cloud, eps), orc.holepunch_find_candidate_faces(m2, cloud, eps))
