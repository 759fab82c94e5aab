"""CPU ORACLE (test infrastructure, not product code).

A numpy/scipy restatement of the reference's NanoWrap conjugate-gradient hot path.
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may
import this module; the product path (``ch_shrinkwrap_b200``) never does.

Pinning: this restatement is checked against outputs of the *unmodified* reference
(``mesh_conj_grad.py`` / ``conj_grad.py`` imported from ``/root/reference`` with its compiled
``conj_grad_utils.c``) by ``oracle/make_golden.py``, which wrote the fixtures in
``tests/golden/`` that ``tests/test_oracle_golden.py`` replays.  The reference has no
tests or golden vectors of its own for this path.

Third-party arithmetic on the path that is not under /root/reference:
``scipy.spatial.cKDTree`` (scipy is unpinned by the reference; 1.18.1 here) supplies
the nearest-centroid search exactly as at ``mesh_conj_grad.py:451-454``.

Every function cites the reference lines it follows.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np
import scipy.spatial

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def _clib():
    """liboracle.so: C restatement of the reference's C helpers (oracle/oracle_c.c)."""
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, 'liboracle.so')
        if not os.path.exists(path):
            from . import build as _b
            _b.build_oracle_c()
        _LIB = ctypes.CDLL(path)
    return _LIB


def _p(a, t):
    return a.ctypes.data_as(ctypes.POINTER(t))


def ah_scatter(v_idx, w, fv, out):
    """Sequential fp32 scatter-add, conj_grad_utils.c:153-162 (point order, then corner, then axis)."""
    v_idx = np.ascontiguousarray(v_idx, dtype=np.int32)
    w = np.ascontiguousarray(w, dtype=np.float32)
    fv = np.ascontiguousarray(fv, dtype=np.float32)
    assert out.dtype == np.float32 and out.flags.c_contiguous
    _clib().orc_ah_scatter(_p(v_idx, ctypes.c_int32), _p(w, ctypes.c_float), _p(fv, ctypes.c_float),
                           _p(out, ctypes.c_float), ctypes.c_int64(v_idx.shape[0]))


def nearest_face(points, face_centers):
    """mesh_conj_grad.py:451-454: exact nearest centroid, fp64, all host threads."""
    tree = scipy.spatial.cKDTree(face_centers)
    return tree.query(points, k=1, workers=-1)


class OracleConjGrad:
    """Restates ``ShrinkwrapMeshConjGrad`` (mesh_conj_grad.py:19-292, 433-588, 770-820,
    1002-1023) and ``TikhonovConjugateGradient.subsearch`` (conj_grad.py:183-229)."""

    def __init__(self, mesh, points, build_point_tree=False):
        self.mesh = mesh
        self.points = points
        # mesh_conj_grad.py:127-130 builds a kd-tree over the points that the active path never
        # queries; reproduced only on request so CPU-baseline timings can include or exclude it.
        self._tree = scipy.spatial.cKDTree(points) if build_point_tree else None
        self._valid = mesh._vertices['halfedge'] != -1            # :44
        self.vertices = mesh._vertices['position']                 # :46 (strided view)
        self.faces = mesh.faces                                     # :47
        nb = mesh._vertices['neighbors']
        n = mesh._halfedges['vertex'][nb]                           # :50
        n[nb == -1] = -1                                            # :51-52
        self.vertex_neighbors = n
        self.M = self.vertices.shape[0]
        self.tests, self.ress, self.prefs = [], [], []              # conj_grad.py:37-39
        self.Lfuncs, self.Lhfuncs = ["I"], ["I"]                    # mesh_conj_grad.py:38 (:39 offers ["wfunc"], ["wfunc"])
        self._prev_loopcount = -1
        self.loopcount = 0
        self.w = None
        self.d = None

    # -- weights: mesh_conj_grad.py:433-516 ------------------------------------
    def compute_weights(self, f):
        fv = f.reshape(-1, 3)
        face_centers = fv[self.faces].mean(1)                      # :443 (fp32)
        dmean, nearest = nearest_face(self.points, face_centers)   # :451-454
        self.nearest = nearest
        self.d = np.vstack([dmean, dmean, dmean]).T                # :483
        v_idx = self.faces[nearest, :]                             # :488
        d = np.zeros(v_idx.shape, 'f4')                            # :491
        for j in range(3):
            dv = fv[v_idx[:, j]] - self.points                     # :494
            d[:, j] = np.sqrt(np.sum(dv * dv, 1))                  # :495
        w = 1.0 / np.maximum(d, 1e-6)                              # :503
        w = w / w.sum(1)[:, None]                                  # :510
        assert not np.any(np.isnan(w))
        return v_idx, w

    def _calc_w(self):                                              # :1018-1023
        if self._prev_loopcount < self.loopcount:
            self._prev_loopcount = self.loopcount
            return True
        return False

    def Afunc(self, f):                                             # :518-551
        if self._calc_w():
            self.w = self.compute_weights(self.f)
        fv = f.reshape(-1, 3)
        out = np.zeros_like(self.points)
        v_idx, w = self.w
        for i in range(3):
            out += fv[v_idx[:, i]] * w[:, i][:, None]              # :544-545
        assert not np.any(np.isnan(out))
        return out.ravel()

    def Ahfunc(self, f):                                            # :553-588
        d = np.zeros([self.M, 3], dtype='f')
        fv = f.reshape(-1, 3)
        v_idx, w = self.w
        ah_scatter(v_idx, w, fv.astype('f'), d)
        assert not np.any(np.isnan(d))
        return d.ravel()

    def point_influence(self):                                      # _membrane_mesh.pyx:1625-1634
        s = self.Ahfunc(np.ones_like(self.res)).reshape(self.vertices.shape)
        return np.sqrt((s * s).sum(1))

    # -- curvature prior: mesh_conj_grad.py:770-820 ----------------------------
    def ncc(self):
        mesh = self.mesh
        hv = mesh._halfedges['vertex']
        vnb = mesh.vertex_neighbors
        vnn = hv[vnb]                                               # :777
        mask = vnb > -1                                             # :778
        ms = mask.sum(1)                                            # :779
        verts, normals = mesh.vertices, mesh.vertex_normals
        with np.errstate(invalid='ignore', divide='ignore'):
            vc = (verts[vnn, :] * mask[:, :, None]).sum(1) / ms[:, None]      # :782
            c_n = verts[vnn, :] - vc[:, None, :]                              # :785
            n_n = normals[vnn, :]                                             # :788
            n_dot_n = (n_n * normals[:, None, :]).sum(2)                      # :796
            alpha = ((c_n * n_n).sum(2)) / np.sqrt(2 * (np.maximum(n_dot_n, 0) + 1))   # :797
            alpha = (alpha * mask).sum(1) / ms                                # :800
            pi = self.point_influence()                                       # :807
            alpha = alpha * np.minimum(pi ** 2, 1)                            # :814
            vc = vc + alpha[:, None] * normals                                # :816
        vc[ms == 0, :] = verts[ms == 0, :]                                    # :818
        return vc

    def I(self, f):
        return f

    def wfunc(self, f):                                             # :725-736
        w = vertex_area_weights(np.ascontiguousarray(self.f), self.vertex_neighbors)
        return f * w

    def _stop_cond(self):                                           # :1009-1016
        if len(self.tests) < 3:
            return False
        a, b, c = self.tests[-3:]
        return (c < b) and (b < a) and (a < 1e-6)

    # -- conj_grad.py:183-229 ---------------------------------------------------
    def subsearch(self, f0, res, fdefs, lams, S):
        n_search = S.shape[1]
        c0 = (res * res).sum()
        prefs = [getattr(self, self.Lfuncs[0])(f0 - fdefs[0])]      # conj_grad.py:191
        wpreds = [(p * p).sum() for p in prefs]
        AS = np.zeros((np.size(res), n_search), 'f')
        LS = np.zeros((len(prefs[0]), n_search, 1), 'f')
        for k in range(n_search):
            AS[:, k] = self.Afunc(S[:, k])[self.mask]
            LS[:, k, 0] = getattr(self, self.Lfuncs[0])(S[:, k])    # conj_grad.py:200
        Hc = np.dot(AS.T, AS)
        Gc = np.dot(AS.T, res)
        Hw = np.zeros((n_search, n_search, 1))
        Gw = np.zeros((n_search, 1))
        H, G = Hc, Gc                                               # aliases (conj_grad.py:208)
        ls = LS[:, :, 0]
        Hw[:, :, 0] = np.dot(ls.T, ls)
        Gw[:, 0] = np.dot(-ls.T, prefs[0])
        l2 = lams[0] * lams[0]
        H += l2 * Hw[:, :, 0]
        G += l2 * Gw[:, 0]
        c = np.linalg.solve(H, G)
        cpred = c0 + np.dot(np.dot(c.T, Hc), c) - np.dot(c.T, Gc)
        wpreds[0] += np.dot(np.dot(c.T, Hw[:, :, 0]), c) - np.dot(c.T, Gw[:, 0])
        fnew = f0 + np.dot(S, c)
        self.c = c
        return fnew, cpred, wpreds

    # -- mesh_conj_grad.py:150-292 ---------------------------------------------
    def search(self, data, lams, num_iters=10, weights=None, sigma_inv=1.0, last_step=True):
        self._prev_loopcount = -1
        if weights is None:
            weights = sigma_inv
        if not np.isscalar(weights):
            self.mask = weights > 0
            weights = weights / weights.mean()
        else:
            self.mask = np.isfinite(data.ravel())
        self.fs = self.vertices.copy()                              # start_guess :1002-1007
        self.f = self.fs.ravel()
        data = data.ravel()
        self.res = 0 * data
        n_smooth = min(1, len(lams))
        n_search = n_smooth + 1
        s_size = n_search + 1
        prefs = np.zeros((np.size(self.f), n_smooth), 'f')
        S = np.zeros((np.size(self.f), s_size), 'f')
        self.loopcount = 0
        while (self.loopcount < num_iters) and (not self._stop_cond()):
            self.loopcount += 1
            self.res[:] = weights * (data - self.Afunc(self.f))     # :222
            defaults = [self.ncc().ravel()]                         # :224
            w = 1.0 / (self.d.ravel() * sigma_inv / 2.0 + 1)        # :231
            self.res *= w                                           # :248
            S[:, 0] = self.Ahfunc(self.res)                         # :253
            prefs[:, 0] = getattr(self, self.Lfuncs[0])(self.f - defaults[0])       # :257
            S[:, 1] = -1.0 * getattr(self, self.Lhfuncs[0])(prefs[:, 0])           # :258
            test = 1.0 - abs((S[:, 0] * S[:, 1]).sum()
                             / (np.linalg.norm(S[:, 0]) * np.linalg.norm(S[:, 1])))   # :262-265
            self.tests.append(test)
            self.ress.append(np.linalg.norm(self.res))
            self.prefs.append(np.linalg.norm(prefs, axis=0))
            fnew, self.cpred, self.wpreds = self.subsearch(self.f, self.res[self.mask], defaults,
                                                           lams, S[:, 0:n_search])    # :274
            if last_step:
                S[:, s_size - 1] = fnew - self.f                    # :282
                n_search = s_size
            self.S = S
            self.f[:] = fnew                                        # :288
            self.mesh._vertices['position'][self._valid] = fnew.reshape(self.vertices.shape)[self._valid]  # :289
            self.mesh._initialize_curvature_vectors()               # :290
        return np.real(self.fs)


class OracleConjGrad64(OracleConjGrad):
    """Same algorithm with the subspace matrices accumulated in float64.  The reference forms ``Hc = AS^T AS`` with a
    float32 sgemm and keeps ``H`` in float32 (conj_grad.py:194,202,208-215), so whenever the search
    directions are nearly dependent (cond(H) up to 1e8 once the last step is added) its own step carries visible
    float32 noise.  This variant separates that noise from implementation error in the parity tests."""

    def subsearch(self, f0, res, fdefs, lams, S):
        n_search = S.shape[1]
        L = getattr(self, self.Lfuncs[0])
        prefs = L(f0 - fdefs[0])
        AS = np.zeros((np.size(res), n_search), 'f')
        LS = np.zeros((len(prefs), n_search), 'f')
        for k in range(n_search):
            AS[:, k] = self.Afunc(S[:, k])[self.mask]
            LS[:, k] = L(S[:, k])
        AS = AS.astype(np.float64)
        S64 = LS.astype(np.float64)
        l2 = float(lams[0]) ** 2
        H = AS.T @ AS + l2 * (S64.T @ S64)
        G = AS.T @ np.asarray(res, np.float64) - l2 * (S64.T @ prefs)
        c = np.linalg.solve(H, G)
        self.c, self.cond = c, float(np.linalg.cond(H))
        return f0 + np.dot(S, c), 0.0, [0.0]


# ---- secondary 1-ring regularisers (conj_grad_utils.c:249-710) -----------------------
def _ring_call(name, f, nbrs, ref, out):
    f = np.ascontiguousarray(f, dtype=np.float32)
    nbrs = np.ascontiguousarray(nbrs, dtype=np.int32)
    ref = np.ascontiguousarray(ref, dtype=np.float32)
    getattr(_clib(), name)(_p(f, ctypes.c_float), _p(nbrs, ctypes.c_int32), _p(ref, ctypes.c_float),
                           _p(out, ctypes.c_float), ctypes.c_int(nbrs.shape[0]), ctypes.c_int(nbrs.shape[1]))
    return out


def l_func(f, nbrs):
    return _ring_call('orc_l_func', f, nbrs, f, np.zeros(f.size, np.float32))


def lh_func(f, nbrs):
    return _ring_call('orc_lh_func', f, nbrs, f, np.zeros(f.size, np.float32))


def lw_func(f, nbrs, ref):
    return _ring_call('orc_lw_func', f, nbrs, ref, np.zeros(f.size, np.float32))


def lhw_func(f, nbrs, ref):
    return _ring_call('orc_lhw_func', f, nbrs, ref, np.zeros(f.size, np.float32))


def vertex_area_weights(ref, nbrs):
    return _ring_call('orc_vertex_area_weights', ref, nbrs, ref, np.zeros(ref.size, np.float32))


# ---- curvature (membrane_mesh_utils.c:915-1250) ------------------------------------
CURV_SCALARS = ('k0', 'k1', 'H', 'K', 'dH', 'dK', 'E', 'pE', 'dE_neighbors')
CURV_VECTORS = ('e0', 'e1', 'dEdN')


def curvature_grad(mesh, dN=0.1, skip_prob=0.0, kc=1.0, kg=-20.0 * 0.0257, c0=0.0, jitter_u=None):
    """C restatement of ``c_curvature_grad``.  ``jitter_u``: the uniform [0,1) numbers the
    reference draws with ``rand()`` (3 per valid vertex, membrane_mesh_utils.c:1017); ``None`` = 0.5
    (no jitter)."""
    M = len(mesh._vertices)
    out = {k: np.zeros(M, np.float32) for k in CURV_SCALARS}
    out.update({k: np.zeros((M, 3), np.float32) for k in CURV_VECTORS})
    verts = np.ascontiguousarray(mesh._vertices)
    faces = np.ascontiguousarray(mesh._faces)
    hes = np.ascontiguousarray(mesh._halfedges)
    ju = None
    if jitter_u is not None:
        ju = np.ascontiguousarray(jitter_u, dtype=np.float64)
        assert ju.size >= 3 * int((verts['halfedge'] != -1).sum())
    fp = ctypes.POINTER(ctypes.c_float)
    _clib().orc_curvature_grad(
        ctypes.c_void_p(verts.ctypes.data), ctypes.c_void_p(faces.ctypes.data), ctypes.c_void_p(hes.ctypes.data),
        ctypes.c_float(dN), ctypes.c_float(skip_prob), ctypes.c_int(M),
        *[out[k].ctypes.data_as(fp) for k in ('k0', 'k1', 'e0', 'e1', 'H', 'K', 'dH', 'dK', 'E', 'pE', 'dE_neighbors')],
        ctypes.c_float(kc), ctypes.c_float(kg), ctypes.c_float(c0), out['dEdN'].ctypes.data_as(fp),
        None if ju is None else ju.ctypes.data_as(ctypes.POINTER(ctypes.c_double)))
    return out


# ---- quality metrics and hole-punch searches -----------------------------------------------
def points_from_mesh_samples(mesh, dx_min=5):
    """The ``d`` array of ``points_from_mesh`` (evaluation_utils.py:59-135) BEFORE the random permutation of :137,
    restated statement by statement (same dtypes: float32 frame, float64 grid)."""
    tris = mesh._vertices['position'][mesh.faces]                                   # :59
    norms = np.cross((tris[:, 2, :] - tris[:, 1, :]), (tris[:, 0, :] - tris[:, 1, :]))   # :61
    nn = np.linalg.norm(norms, axis=1)
    nan_mask = nn != 0                                                              # :67
    norms = norms[nan_mask] / nn[nan_mask, None]
    tris = tris[nan_mask, :, :]
    v0 = tris[:, 1, :] - tris[:, 0, :]                                              # :78-81
    e0n = np.linalg.norm(v0, axis=1)
    e0 = v0 / e0n[:, None]
    e1 = np.cross(norms, e0, axis=1)
    x0 = (tris[:, 0, :] * e0).sum(1); y0 = (tris[:, 0, :] * e1).sum(1)              # :84-89
    x1 = (tris[:, 1, :] * e0).sum(1); y1 = (tris[:, 1, :] * e1).sum(1)
    x2 = (tris[:, 2, :] * e0).sum(1); y2 = (tris[:, 2, :] * e1).sum(1)
    x0x1x2 = np.vstack([x0, x1, x2]).T; y0y1y2 = np.vstack([y0, y1, y2]).T          # :92-97
    xl, xu = np.min(x0x1x2, axis=1), np.max(x0x1x2, axis=1)
    yl, yu = np.min(y0y1y2, axis=1), np.max(y0y1y2, axis=1)
    x1x0, x2x1, x0x2 = x1 - x0, x2 - x1, x0 - x2                                    # :100-110
    with np.errstate(divide='ignore', invalid='ignore'):
        m0 = (y1 - y0) / x1x0; m0[x1x0 == 0] = 0
        m1 = (y2 - y1) / x2x1; m1[x2x1 == 0] = 0
        m2 = (y0 - y2) / x0x2; m2[x0x2 == 0] = 0
    s1, s2 = np.sign(m1), np.sign(m2)
    d = []
    for i in range(tris.shape[0]):                                                  # :117-133
        x = np.arange(xl[i] - x0[i] - dx_min / 2, xu[i] - x0[i], dx_min)
        y = np.arange(yl[i] - y0[i] - dx_min / 2, yu[i] - y0[i], dx_min)
        X, Y = np.meshgrid(x, y)
        X_mask = (Y > X * m0[i]) & (s1[i] * Y > s1[i] * (y1[i] - y0[i] + (X - x1[i] + x0[i]) * m1[i])) & \
                 (s2[i] * Y < s2[i] * (y2[i] - y0[i] + (X - x2[i] + x0[i]) * m2[i]))
        pos = X[X_mask].ravel()[:, None] * e0[i, None, :] + Y[X_mask].ravel()[:, None] * e1[i, None, :] + tris[i, 0, :]
        d.append(pos)
    return np.vstack(d) if d else np.zeros((0, 3))


def average_squared_distance(points0, points1):
    """evaluation_utils.py:147-180."""
    t0, t1 = scipy.spatial.cKDTree(points0), scipy.spatial.cKDTree(points1)
    e0, _ = t0.query(points1, k=1)
    e1, _ = t1.query(points0, k=1)
    return np.nansum(e0 ** 2) / len(e0), np.nansum(e1 ** 2) / len(e1)


def holepunch_find_candidate_faces(mesh, points, eps=10.0):
    """_membrane_mesh.pyx:877-887."""
    tree = scipy.spatial.cKDTree(points)
    dist, _ = tree.query(mesh._vertices['position'][mesh.faces].mean(1))
    inds = np.flatnonzero(mesh._faces['halfedge'] != -1).astype('i4')
    return inds[dist > eps]


def holepunch_pairs(mesh, candidates):
    """Raw ``pairs`` of c_holepunch_pair_candidate_faces (membrane_mesh_utils.c:1301-1379, restated in oracle_c.c)."""
    verts, faces, hes = (np.ascontiguousarray(a) for a in (mesh._vertices, mesh._faces, mesh._halfedges))
    cand = np.ascontiguousarray(candidates, dtype=np.int32)
    pairs = -1 * np.ones(len(cand), np.int32)
    _clib().orc_holepunch_pair(ctypes.c_void_p(verts.ctypes.data), ctypes.c_void_p(faces.ctypes.data), ctypes.c_void_p(hes.ctypes.data),
                               _p(cand, ctypes.c_int32), ctypes.c_int(len(cand)), _p(pairs, ctypes.c_int32))
    return pairs


def holepunch_pair_candidate_faces(mesh, candidates):
    """_membrane_mesh.pyx:898-907 (the USE_C branch)."""
    candidates = np.asarray(candidates)
    pairs = holepunch_pairs(mesh, candidates)
    pair_inds = pairs != -1
    new_inds = np.cumsum(pair_inds) - 1
    return candidates[pair_inds], new_inds[pairs[pair_inds]]


# ---- block driver (_membrane_mesh.pyx:1427-1560, 1201-1219) -----------------------------------------------------------
def opt_conjugate_gradient(mesh, points, sigma, max_iter=10, step_size=1.0, weights=None, remesh=None, log=None, **kwargs):
    """Restates ``MembraneMesh.opt_conjugate_gradient`` for a harness mesh (mini-mesh: structured arrays, ``update_geometry``,
    ``area``, ``_mean_edge_length``) with the oracle solver.  Topology edits are PYME's (absent): ``remesh(mesh, target_length)``
    is the caller's stand-in, neck removal stops at the candidate list (:1212-1213).  ``log`` collects what a test compares:
    per block ('block', j, n_it), ('necks', j, candidates), ('remesh', j, target_length)."""
    import math
    log = [] if log is None else log
    rfq, drf = mesh.remesh_frequency, mesh.delaunay_remesh_frequency
    r = (rfq != 0) and (rfq <= max_iter)                                           # :1430
    dr = (drf != 0) and (drf <= max_iter)
    if r and dr:
        rf = math.gcd(rfq, drf)                                                    # :1435
    elif r:
        rf = rfq
    elif dr:
        rf = drf
    else:
        rf = max_iter
    if r:
        initial_length = mesh._mean_edge_length                                    # :1444
        if kwargs.get('minimum_edge_length', -1) < 0:
            final_length = np.clip(np.min(sigma) / 2.5, 1.0, 50.0)
        else:
            final_length = kwargs.get('minimum_edge_length')
        m = (final_length - initial_length) / (rf * np.ceil(max_iter / rf))        # :1455
    neck_first_iter = getattr(mesh, 'neck_first_iter', -1)
    if np.isscalar(sigma):                                                         # :1460-1473
        s = float(sigma)
    elif (len(sigma.shape) == 1) and (sigma.shape[0] == points.shape[0]):
        s = 1.0 / np.repeat(sigma, points.shape[1])
    elif (len(sigma.shape) == 2) and (sigma.shape[0] == points.shape[0]) and (sigma.shape[1] == points.shape[1]):
        s = (1.0 / sigma.ravel())
    else:
        raise ValueError('Sigma must be of shape (P,) or (P,3).')
    mesh.cg = None
    j = 0
    lams = [step_size * mesh.kc / 2.0, mesh.shrink_weight] if mesh.shrink_weight > 0 else [step_size * mesh.kc / 2.0]   # :1483-1486
    n_iter = min(max_iter, getattr(mesh, 'truncate_at', max_iter))                 # :1490
    while j < n_iter:
        mesh.cg = OracleConjGrad(mesh, points)                                     # :1510
        n_it = min(n_iter - j, rf)
        mesh.cg.search(points, lams=lams, num_iters=n_it, sigma_inv=s, weights=weights)   # :1516
        j += n_it
        log.append(('block', j, n_it))
        mesh.update_geometry()                                                     # :1524-1527 (normals, neighbours refreshed)
        if r and ((j % rfq) == 0):                                                 # :1537
            if (neck_first_iter > 0) and (j > neck_first_iter):
                K = curvature_grad(mesh, kc=mesh.kc, kg=mesh.kg, c0=mesh.c0)['K']  # :1201-1213 (remove_necks up to the candidate list)
                lo, hi = getattr(mesh, 'neck_threshold_low', -1e-4), getattr(mesh, 'neck_threshold_high', 1e-2)
                log.append(('necks', j, np.flatnonzero((K < lo) | (K > hi))))
            target_length = (initial_length + m * (j + 1))                         # :1545
            log.append(('remesh', j, float(target_length)))
            if remesh is not None:
                remesh(mesh, target_length)
            mesh.cg = None
    return j
