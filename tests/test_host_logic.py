"""Host-side logic: mini-mesh invariants, synthetic generators, facade schedule, shim argument handling."""
import numpy as np
import pytest

from conftest import make_case


def test_minimesh_halfedge_invariants():
    from ch_shrinkwrap_b200 import minimesh
    m = minimesh.sphere_mesh(10.0, 5)
    he, v, f = m._halfedges, m._vertices, m._faces
    assert len(v) == 10 * 25 + 2 and len(f) == 20 * 25
    assert np.all(he['twin'][he['twin']] == np.arange(len(he)))          # closed surface: twins pair up
    assert np.all(he['next'][he['prev']] == np.arange(len(he)))
    # face corner order: prev.vertex, h.vertex, next.vertex == mesh.faces rows (membrane_mesh_utils.c:1271-1274)
    h = f['halfedge']
    rec = np.stack([he['vertex'][he['prev'][h]], he['vertex'][h], he['vertex'][he['next'][h]]], 1)
    assert np.array_equal(rec, m.faces)
    # neighbours are OUTGOING half-edges, ring ordered, -1 terminated; valence 5 or 6 on a geodesic sphere
    nb = v['neighbors']
    val = (nb != -1).sum(1)
    assert set(np.unique(val)) <= {5, 6} and np.array_equal(val, v['valence'])
    src = he['vertex'][he['prev']]
    for i in (0, 7, 100):
        hs = nb[i][nb[i] != -1]
        assert np.all(src[hs] == i)
        assert len(set(he['vertex'][hs])) == len(hs)
    assert abs(m.area() - 4 * np.pi * 100.0) / (4 * np.pi * 100.0) < 0.02
    assert np.allclose(np.linalg.norm(m.vertex_normals, axis=1), 1.0, atol=1e-5)
    assert np.all((m.vertex_normals * m.vertices).sum(1) > 0)              # outward


def test_minimesh_open_boundary():
    from ch_shrinkwrap_b200 import minimesh
    m = minimesh.planar_mesh(4.0, 4)
    assert (m._halfedges['twin'] == -1).sum() == 16
    assert np.all(m._vertices['valence'] >= 1)
    nv = m.neighbor_vertices()
    assert nv.shape == (25, 20) and nv.max() < 25


def test_geodesic_counts():
    from ch_shrinkwrap_b200 import minimesh
    for n in (1, 2, 3, 7):
        v, f = minimesh.geodesic_sphere(n)
        assert len(v) == 10 * n * n + 2 and len(f) == 20 * n * n
        assert np.allclose(np.linalg.norm(v, axis=1), 1.0)
        e = set()
        for a, b, c in f:
            e |= {(a, b), (b, c), (c, a)}
        assert all((b, a) in e for a, b in e)                              # consistently oriented, closed


def test_synth_cloud_is_seeded_and_on_surface():
    from ch_shrinkwrap_b200 import synth
    shape = synth.two_lobed()
    p1, s1 = synth.smlm_cloud(shape, 2000, seed=4)
    p2, s2 = synth.smlm_cloud(shape, 2000, seed=4)
    assert np.array_equal(p1, p2) and np.array_equal(s1, s2)
    d = np.abs(shape.sdf(p1.astype(np.float64)))
    assert np.median(d) < 15.0                       # jitter ~ localisation precision
    assert (d > 60).mean() > 0.03                    # ~10 % background
    assert 3.0 < np.median(s1[:, 0]) < 9.0 and np.median(s1[:, 2]) > 2.5 * np.median(s1[:, 0])
    v, f = __import__('ch_shrinkwrap_b200.minimesh', fromlist=['x']).geodesic_sphere(12)
    r = synth.radial_surface(shape, v)
    a, _ = synth.mesh_surface_cloud(v * r[:, None], f, 5000, seed=1, threads=1)
    b, _ = synth.mesh_surface_cloud(v * r[:, None], f, 5000, seed=1, threads=4)
    assert np.array_equal(a, b)


def test_facade_block_schedule_matches_reference_driver():
    """opt_conjugate_gradient block structure (_membrane_mesh.pyx:1430-1517) with the solver stubbed out."""
    from ch_shrinkwrap_b200 import membrane_mesh as mm
    from ch_shrinkwrap_b200 import minimesh
    calls = []

    class FakeCG:
        def __init__(self, mesh, points, **kw):
            calls.append(['new'])

        def search(self, points, lams, num_iters, sigma_inv, weights):
            calls[-1] += [num_iters, lams, np.isscalar(sigma_inv) or sigma_inv.shape]
    real = mm.ShrinkwrapMeshConjGrad
    mm.ShrinkwrapMeshConjGrad = FakeCG
    try:
        v, f = minimesh.geodesic_sphere(3)
        m = mm.MembraneMesh(v * 100, f, kc=1.0, step_size=20.0, remesh_frequency=5, delaunay_remesh_frequency=0, max_iter=12)
        pts = np.zeros((10, 3), np.float32)
        n = m.shrink_wrap(pts, np.full((10, 3), 2.0, np.float32), minimum_edge_length=5)
        assert n == 12 and [c[1] for c in calls] == [5, 5, 2]
        assert calls[0][2] == [10.0] and calls[0][3] == (30,)             # lam = step_size*kc/2 ; sigma (P,3) -> 1/sigma raveled
        calls.clear()
        m.remesh_frequency = 0
        m.shrink_wrap(pts, 7.0, max_iter=4)                                # scalar sigma passes through un-inverted (:1460)
        assert [c[1] for c in calls] == [4] and calls[0][3] is True
        with pytest.raises(ValueError):
            m.shrink_wrap(pts, np.ones((3, 2)), max_iter=2)
        calls.clear()
        m.truncate_at = 3
        m.shrink_wrap(pts, 7.0, max_iter=10)
        assert sum(c[1] for c in calls) == 3
    finally:
        mm.ShrinkwrapMeshConjGrad = real


def test_recipe_module_surface():
    from ch_shrinkwrap_b200.recipe_modules.surface_fitting import ShrinkwrapMembrane
    mod = ShrinkwrapMembrane()
    # trait names and defaults of the reference module (surface_fitting.py:13-42)
    assert mod.max_iters == 39 and mod.curvature_weight == 20.0 and mod.remesh_frequency == 5
    assert mod.neck_threshold_low == -1e-3 and mod.neck_threshold_high == 1e-2 and mod.neck_first_iter == 9
    assert mod.kc == 1.0 and mod.minimum_edge_length == 5 and mod.sigma_x == 'error_x'


def test_image_module_surface_and_voxel_conversion():
    from ch_shrinkwrap_b200.recipe_modules.surface_fitting import ImageShrinkwrapMembrane, image_to_weighted_points
    mod = ImageShrinkwrapMembrane()
    # defaults of the reference module (surface_fitting.py:252-273)
    assert mod.max_iters == 100 and mod.curvature_weight == 10.0 and mod.shrink_weight == 1.0 and mod.minimum_edge_length == -1.0
    assert mod.neck_threshold_low == -1e-4 and mod.cut_frequency == 0
    data = np.zeros((3, 4, 2))
    data[1, 2, 0] = 5.0
    data[2, 3, 1] = 7.0
    data[0, 0, 0] = -1.0
    pts, w, sigma = image_to_weighted_points(data, (10.0, 20.0, 30.0), (100.0, 200.0, 300.0))
    assert sigma == 10.0 and pts.shape == (2, 3) and w.shape == (6,)
    assert np.array_equal(pts, [[110.0, 240.0, 300.0], [120.0, 260.0, 330.0]])
    assert np.array_equal(w, [5, 5, 5, 7, 7, 7])


def test_solver_shim_chooses_the_record_upload_and_passes_the_half_edge_stride():
    """Host logic of ShrinkwrapMeshConjGrad without a GPU: a recording stand-in for the C ABI handle.  PYME-layout meshes
    go up as raw records with the half-edge 'vertex' field addressed in place (stride 28); search() asks the library to
    write the positions back into the records and keeps the reference's side effects (mesh_conj_grad.py:289-290)."""
    import ctypes
    from ch_shrinkwrap_b200 import mesh_conj_grad as mcg, synth
    from ch_shrinkwrap_b200.membrane_mesh import MembraneMesh

    calls = []

    class FakeHandle:
        h = 1

        def call(self, name, *args):
            calls.append((name, args))
            if name == 'nw_search':
                args[-1]._obj.value = args[1]                 # n_done = num_iters
            return 0

    class FakeSession:
        handle = FakeHandle()
        points_key = None
        P = 0

        def set_points(self, points, sigma_inv, weights):
            calls.append(('set_points', (points.shape, np.isscalar(sigma_inv), weights is None)))
            self.P = len(points)

    base = synth.star_mesh(synth.Sphere(100.0), 3, scale=1.1)
    mesh = MembraneMesh(mesh=base)
    mesh._nw_session = FakeSession()
    inited = []
    mesh._initialize_curvature_vectors = lambda: inited.append(1)
    pts = (np.asarray(mesh.vertices) * 0.9).astype(np.float32)
    cg = mcg.ShrinkwrapMeshConjGrad(mesh, pts)
    assert cg._vertex_mask is None                            # the valid mask is built lazily
    out = cg.search(pts, lams=[5.0], num_iters=3, sigma_inv=0.1)
    names = [c[0] for c in calls]
    assert names == ['set_points', 'nw_set_topology_records', 'nw_set_regulariser', 'nw_search', 'nw_get_positions_strided']
    rec = dict(calls)['nw_set_topology_records']
    assert rec[3] == mesh._halfedges.strides[0] == 28         # half-edge 'vertex' field gathered in place
    assert rec[4] == len(mesh._halfedges) and rec[5] == len(mesh._vertices) and rec[6] == len(mesh.faces)
    wb = dict(calls)['nw_get_positions_strided']
    assert wb[1] == mesh._vertices.strides[0] == 120 and wb[2] == 1
    assert out.shape == (len(mesh._vertices), 3) and cg.loopcount == 3 and len(cg.tests) == 3
    assert inited == [1]                                       # :290
    # a second search() on the same object re-sends the (possibly edited) positions, not the topology
    calls.clear()
    cg.search(pts, lams=[5.0], num_iters=2, sigma_inv=0.1)
    assert [c[0] for c in calls] == ['set_points', 'nw_set_positions', 'nw_set_regulariser', 'nw_search', 'nw_get_positions_strided']


# ---- sessions and the recipe's mesh factory (no GPU: the library handle is stubbed) -------------------------------------
class _FakeHandle:
    created = 0

    def __init__(self, device=0):
        type(self).created += 1
        self.h = object()
        self.calls = []

    def call(self, name, *args):
        self.calls.append(name)

    def close(self):
        self.h = None


def test_session_is_reused_for_a_mesh_without_dict(monkeypatch):
    """a cdef-style mesh (no __dict__) must get the SAME session for every remesh block: one handle, one
    nw_set_points."""
    from ch_shrinkwrap_b200 import _lib, mesh_conj_grad as mcg
    monkeypatch.setattr(_lib, 'Handle', _FakeHandle)
    monkeypatch.setattr(mcg, '_FALLBACK_SESSIONS', {})
    _FakeHandle.created = 0

    class Slotted:
        __slots__ = ('x',)

    m = Slotted()
    s1 = mcg._session_for(m)
    s2 = mcg._session_for(m)
    assert s1 is s2 and _FakeHandle.created == 1
    pts = np.zeros((5, 3), np.float32)
    for _ in range(3):                      # three "blocks"
        mcg._session_for(m).set_points(pts, 1.0, None)
    assert s1.handle.calls.count('nw_set_points') == 1
    # the table is bounded and closes what it evicts
    others = [Slotted() for _ in range(mcg._FALLBACK_LIMIT)]
    for o in others:
        mcg._session_for(o)
    assert len(mcg._FALLBACK_SESSIONS) == mcg._FALLBACK_LIMIT and s1.handle.h is None
    mcg.release_session(others[-1])
    assert id(others[-1]) not in mcg._FALLBACK_SESSIONS


def test_mesh_factory_mixes_the_gpu_path_into_the_host_mesh_class(monkeypatch):
    """when the reference's ``ch_shrinkwrap._membrane_mesh`` imports, the recipe must still run the GPU
    solver -- not MembraneMesh.opt_conjugate_gradient's own CPU ShrinkwrapMeshConjGrad (_membrane_mesh.pyx:1428)."""
    import sys
    import types
    from ch_shrinkwrap_b200 import membrane_mesh as mm
    from ch_shrinkwrap_b200 import minimesh
    from ch_shrinkwrap_b200.recipe_modules import surface_fitting as sf

    ran = []

    class HostMembraneMesh(minimesh.MiniMesh):          # stands in for the PYME-backed class
        def __init__(self, mesh=None, **kw):
            minimesh.MiniMesh.__init__(self, mesh.vertices, mesh.faces)
            self.kc, self.step_size, self.max_iter, self.remesh_frequency, self.delaunay_remesh_frequency = 1.0, 1.0, 4, 2, 0
            self.shrink_weight, self.search_k, self.search_rad, self.smooth_curvature = 0, 200, 100, False
            for k, v in kw.items():
                setattr(self, k, v)

        def opt_conjugate_gradient(self, *a, **k):      # the CPU path that must NOT run
            ran.append('cpu')

        def remesh(self, n, target_length, l, n_relax=0):
            ran.append('host remesh')

    pkg = types.ModuleType('ch_shrinkwrap')
    mod = types.ModuleType('ch_shrinkwrap._membrane_mesh')
    mod.MembraneMesh = HostMembraneMesh
    pkg._membrane_mesh = mod
    monkeypatch.setitem(sys.modules, 'ch_shrinkwrap', pkg)
    monkeypatch.setitem(sys.modules, 'ch_shrinkwrap._membrane_mesh', mod)
    monkeypatch.setattr(sf, '_GPU_MESH_CLASS', {})

    class SolverSpy:
        def __init__(self, mesh, points, **kw):
            ran.append('gpu solver built')

        def search(self, points, lams, num_iters=10, sigma_inv=1.0, weights=None):
            ran.append('gpu search %d' % num_iters)

    monkeypatch.setattr(mm, 'ShrinkwrapMeshConjGrad', SolverSpy)
    base = minimesh.sphere_mesh(100.0, 3)
    mesh = sf._mesh_factory(base, kc=1.0, max_iter=4, step_size=10.0, remesh_frequency=2, delaunay_remesh_frequency=0)
    assert isinstance(mesh, HostMembraneMesh) and isinstance(mesh, mm.ShrinkwrapMeshMixin)
    assert type(mesh).opt_conjugate_gradient is mm.ShrinkwrapMeshMixin.opt_conjugate_gradient
    pts = np.zeros((10, 3), np.float32)
    mesh.shrink_wrap(pts, np.ones((10, 3), np.float32), method='conjugate_gradient')
    assert 'cpu' not in ran
    assert ran.count('gpu solver built') == 2 and ran.count('gpu search 2') == 2      # 4 iterations in blocks of 2
    assert 'host remesh' in ran                                                            # topology stays with the host class


def test_smoothing_skip_is_reported(caplog):
    from ch_shrinkwrap_b200 import membrane_mesh as mm
    m = mm.MembraneMesh.__new__(mm.MembraneMesh)
    m.smooth_curvature = True
    m.curvature_grad_c = lambda *a, **k: None
    import logging
    with caplog.at_level(logging.WARNING):
        m._populate_curvature_grad()
    assert 'unsmoothed' in caplog.text and m._curvature_smoothing.startswith('skipped')
