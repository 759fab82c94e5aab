"""Host half-edge mini-mesh with PYME's record layouts.

The reference solver hangs off ``PYME.experimental._triangle_mesh.TriangleMesh``
(not in the reference repository, not installed here).  This
module is *harness input*, not reference behaviour: it builds the handful of
attributes the hot path reads (``mesh_conj_grad.py:44-52,289-290,777-816``) from a
plain ``(vertices, faces)`` pair so tests and ``bench.py`` have something to fit.

Record layouts mirror the reference's own header (``membrane_mesh_utils.h:31-65``):
``halfedge_t`` 28 B, ``face_t`` 24 B, ``vertex_t`` 120 B, all 4-byte fields, packed.

Conventions used (the ones the reference C code relies on):

* ``halfedge['vertex']`` is the vertex the half-edge points **to**
  (``mesh_conj_grad.py:50``).
* ``vertex['neighbors']`` holds **outgoing half-edge indices** in ring order,
  ``-1`` terminated (``membrane_mesh_utils.c:981-1008``).
* face corner order is ``prev.vertex, h.vertex, next.vertex``
  (``membrane_mesh_utils.c:1271-1274``).
* a deleted vertex keeps its row and has ``halfedge == -1`` (``mesh_conj_grad.py:44``).
"""
from __future__ import annotations

import numpy as np

NEIGHBORSIZE = 20  # membrane_mesh_utils.h:29

HALFEDGE_DTYPE = np.dtype([
    ('vertex', 'i4'), ('face', 'i4'), ('twin', 'i4'), ('next', 'i4'), ('prev', 'i4'),
    ('length', 'f4'), ('component', 'i4')])
FACE_DTYPE = np.dtype([
    ('halfedge', 'i4'), ('normal', 'f4', (3,)), ('area', 'f4'), ('component', 'i4')])
VERTEX_DTYPE = np.dtype([
    ('position', 'f4', (3,)), ('normal', 'f4', (3,)), ('halfedge', 'i4'), ('valence', 'i4'),
    ('neighbors', 'i4', (NEIGHBORSIZE,)), ('component', 'i4'), ('locally_manifold', 'i4')])

assert HALFEDGE_DTYPE.itemsize == 28 and FACE_DTYPE.itemsize == 24 and VERTEX_DTYPE.itemsize == 120


class MiniMesh:
    """Minimal half-edge triangle mesh exposing what the shrinkwrap solver reads."""

    def __init__(self, vertices, faces):
        vertices = np.ascontiguousarray(vertices, dtype=np.float32)
        faces = np.ascontiguousarray(faces, dtype=np.int32)
        M, F = len(vertices), len(faces)
        self._vertices = np.zeros(M, VERTEX_DTYPE)
        self._faces = np.zeros(F, FACE_DTYPE)
        self._halfedges = np.zeros(3 * F, HALFEDGE_DTYPE)
        self._vertices['position'] = vertices
        self._vertices['halfedge'] = -1
        self._vertices['neighbors'] = -1
        self._vertices['locally_manifold'] = 1
        self._build_topology(faces)
        self.update_geometry()
        self.cg = None
        self._curv_cache = None

    # ---- topology -------------------------------------------------------------
    def _build_topology(self, faces):
        F = len(faces)
        M = len(self._vertices)
        he = self._halfedges
        h = np.arange(3 * F, dtype=np.int32)
        f = h // 3
        k = h % 3
        # half-edge 3f+k runs faces[f,k] -> faces[f,(k+1)%3]
        src = faces[f, k]
        dst = faces[f, (k + 1) % 3]
        he['vertex'] = dst
        he['face'] = f
        he['next'] = 3 * f + (k + 1) % 3
        he['prev'] = 3 * f + (k + 2) % 3
        he['component'] = 0
        self._faces['halfedge'] = 3 * np.arange(F, dtype=np.int32)
        # corner order prev.vertex, h.vertex, next.vertex of half-edge 3f == faces[f] rolled
        # so that mesh.faces (below) reproduces the input rows exactly:
        # prev(3f)=3f+2 -> vertex faces[f,0]; 3f -> faces[f,1]; next=3f+1 -> faces[f,2]

        # twins: match directed edge (src,dst) with (dst,src)
        key = src.astype(np.int64) * M + dst
        rkey = dst.astype(np.int64) * M + src
        order = np.argsort(key, kind='stable')
        pos = np.searchsorted(key[order], rkey)
        pos_c = np.minimum(pos, len(order) - 1)
        found = key[order][pos_c] == rkey
        twin = np.where(found, order[pos_c], -1).astype(np.int32)
        he['twin'] = twin

        # ring-ordered outgoing half-edges per vertex
        start = np.full(M, -1, np.int32)
        # prefer a boundary start (outgoing half-edge whose prev has no twin) so open fans are complete
        start[src[::-1]] = h[::-1]
        # "clockwise-most" start for boundary vertices: outgoing h whose twin == -1 when walking backwards
        # walking rule used below: h -> twin[prev[h]] ; its inverse is h -> next[twin[h]]
        bnd = twin == -1
        # outgoing half-edges from which the backward walk (next[twin[h]]) cannot proceed
        start[src[bnd]] = h[bnd]
        nbrs = np.full((M, NEIGHBORSIZE), -1, np.int32)
        cur = start.copy()
        alive = cur != -1
        valence = np.zeros(M, np.int32)
        for j in range(NEIGHBORSIZE):
            if not alive.any():
                break
            idx = np.flatnonzero(alive)
            nbrs[idx, j] = cur[idx]
            valence[idx] += 1
            nxt = twin[he['prev'][cur[idx]]]
            done = (nxt == -1) | (nxt == start[idx])
            cur[idx] = np.where(done, -1, nxt)
            alive[idx] = ~done
        self._vertices['halfedge'] = start
        self._vertices['neighbors'] = nbrs
        self._vertices['valence'] = valence
        self._faces_cache = faces.copy()

    # ---- geometry -------------------------------------------------------------
    def update_geometry(self):
        """Recompute face normals/areas, vertex normals and edge lengths from positions."""
        pos = self._vertices['position']
        faces = self.faces
        v0, v1, v2 = pos[faces[:, 0]], pos[faces[:, 1]], pos[faces[:, 2]]
        n = np.cross((v1 - v0).astype(np.float64), (v2 - v0).astype(np.float64))
        nn = np.sqrt((n * n).sum(1))
        area = 0.5 * nn
        with np.errstate(invalid='ignore', divide='ignore'):
            fn = np.where(nn[:, None] > 0, n / nn[:, None], 0.0)
        self._faces['normal'] = fn.astype(np.float32)
        self._faces['area'] = area.astype(np.float32)
        vn = np.zeros((len(pos), 3), np.float64)
        for k in range(3):            # area-weighted sum of face normals (bincount: add.at is 20x slower)
            for a in range(3):
                vn[:, a] += np.bincount(faces[:, k], weights=n[:, a], minlength=len(pos))
        vnn = np.sqrt((vn * vn).sum(1))
        with np.errstate(invalid='ignore', divide='ignore'):
            vn = np.where(vnn[:, None] > 0, vn / vnn[:, None], 0.0)
        self._vertices['normal'] = vn.astype(np.float32)
        he = self._halfedges
        src = he['vertex'][he['prev']]
        d = pos[he['vertex']] - pos[src]
        he['length'] = np.sqrt((d * d).sum(1))

    # ---- attributes the solver reads -------------------------------------------
    @property
    def faces(self):
        return self._faces_cache

    @property
    def vertices(self):
        return self._vertices['position']

    @property
    def vertex_neighbors(self):
        return self._vertices['neighbors']

    @property
    def vertex_normals(self):
        return self._vertices['normal']

    @property
    def face_normals(self):
        return self._faces['normal']

    @property
    def _mean_edge_length(self):
        return float(self._halfedges['length'].mean())

    def area(self):
        return float(self._faces['area'].sum())

    @property
    def point_influence(self):
        # _membrane_mesh.pyx:1625-1634
        s = self.cg.Ahfunc(np.ones_like(self.cg.res)).reshape(self.vertices.shape)
        return np.sqrt((s * s).sum(1))

    def _initialize_curvature_vectors(self):
        self._curv_cache = None

    def neighbor_vertices(self):
        """(M,20) neighbour *vertex* ids, -1 padded (mesh_conj_grad.py:50-54)."""
        nb = self._vertices['neighbors']
        n = self._halfedges['vertex'][nb]
        n[nb == -1] = -1
        return np.ascontiguousarray(n, dtype=np.int32)


# ---- generators -----------------------------------------------------------------
def _icosahedron():
    t = (1.0 + 5.0 ** 0.5) / 2.0
    v = np.array([[-1, t, 0], [1, t, 0], [-1, -t, 0], [1, -t, 0],
                  [0, -1, t], [0, 1, t], [0, -1, -t], [0, 1, -t],
                  [t, 0, -1], [t, 0, 1], [-t, 0, -1], [-t, 0, 1]], dtype=np.float64)
    v /= np.linalg.norm(v, axis=1)[:, None]
    f = np.array([[0, 11, 5], [0, 5, 1], [0, 1, 7], [0, 7, 10], [0, 10, 11],
                  [1, 5, 9], [5, 11, 4], [11, 10, 2], [10, 7, 6], [7, 1, 8],
                  [3, 9, 4], [3, 4, 2], [3, 2, 6], [3, 6, 8], [3, 8, 9],
                  [4, 9, 5], [2, 4, 11], [6, 2, 10], [8, 6, 7], [9, 8, 1]], dtype=np.int64)
    return v, f


def geodesic_sphere(n):
    """Unit geodesic sphere: every icosahedron face split into n*n triangles.

    Returns (vertices float64 (10 n^2 + 2, 3), faces int32 (20 n^2, 3)), outward orientation.
    """
    v0, f0 = _icosahedron()
    if n == 1:
        return v0, f0.astype(np.int32)
    # barycentric lattice per face, vertices welded by a quantised key
    ii, jj = np.meshgrid(np.arange(n + 1), np.arange(n + 1), indexing='ij')
    keep = (ii + jj) <= n
    ii, jj = ii[keep], jj[keep]
    kk = n - ii - jj
    lut = -np.ones((n + 1, n + 1), np.int64)
    lut[ii, jj] = np.arange(len(ii))
    # local triangles
    a, b = np.meshgrid(np.arange(n), np.arange(n), indexing='ij')
    up = (a + b) <= n - 1
    au, bu = a[up], b[up]
    tri_up = np.stack([lut[au, bu], lut[au + 1, bu], lut[au, bu + 1]], 1)
    dn = (a + b) <= n - 2
    ad, bd = a[dn], b[dn]
    tri_dn = np.stack([lut[ad + 1, bd], lut[ad + 1, bd + 1], lut[ad, bd + 1]], 1)
    tri_local = np.concatenate([tri_up, tri_dn], 0)
    pts = []
    tris = []
    for fi in range(20):
        A, B, C = v0[f0[fi, 0]], v0[f0[fi, 1]], v0[f0[fi, 2]]
        p = (kk[:, None] * A + ii[:, None] * B + jj[:, None] * C) / n
        pts.append(p)
        tris.append(tri_local + fi * len(ii))
    pts = np.concatenate(pts, 0)
    tris = np.concatenate(tris, 0)
    # weld duplicates on shared edges / corners
    q = np.round(pts * (n * 64)).astype(np.int64)
    _, first, inv = np.unique(q, axis=0, return_index=True, return_inverse=True)
    inv = inv.reshape(-1)
    verts = pts[first]
    verts /= np.linalg.norm(verts, axis=1)[:, None]
    faces = inv[tris].astype(np.int32)
    return verts, faces


def spatially_sorted(vertices, faces, bits=10):
    """Renumber vertices (and sort faces) along a Morton curve.  The order of a mesh's vertex array is arbitrary; a
    spatially coherent one (what marching-cubes / remeshing codes produce) keeps 1-ring gathers cache-local."""
    v = np.asarray(vertices, dtype=np.float64)
    lo, hi = v.min(0), v.max(0)
    q = np.clip(((v - lo) / max(float((hi - lo).max()), 1e-30) * (2 ** bits - 1)).astype(np.uint64), 0, 2 ** bits - 1)

    def spread(x):
        r = np.zeros_like(x)
        for b in range(bits):
            r |= ((x >> np.uint64(b)) & np.uint64(1)) << np.uint64(3 * b)
        return r
    key = spread(q[:, 0]) | (spread(q[:, 1]) << np.uint64(1)) | (spread(q[:, 2]) << np.uint64(2))
    order = np.argsort(key, kind='stable')
    rank = np.empty(len(v), np.int64)
    rank[order] = np.arange(len(v))
    f = rank[np.asarray(faces)]
    f = f[np.argsort(f.min(1), kind='stable')]
    return v[order], f.astype(np.int32)


def icosphere(level):
    """Icosahedron subdivided ``level`` times (10*4^level + 2 vertices)."""
    return geodesic_sphere(2 ** level)


def sphere_mesh(radius=1.0, n=8, centre=(0.0, 0.0, 0.0)):
    v, f = geodesic_sphere(n)
    return MiniMesh(v * radius + np.asarray(centre), f)


def planar_mesh(a=1.0, n_subdivision=1):
    """Square [0,a]^2 split into 2*n^2 triangles (reference tests/test_membrane_mesh.py:23-41 builds the same shape)."""
    g = np.linspace(0.0, a, n_subdivision + 1)
    x, y = np.meshgrid(g, g, indexing='ij')
    v = np.stack([x.ravel(), y.ravel(), np.zeros(x.size)], 1)
    idx = np.arange((n_subdivision + 1) ** 2).reshape(n_subdivision + 1, n_subdivision + 1)
    ll, lr, ul, ur = idx[:-1, :-1].ravel(), idx[1:, :-1].ravel(), idx[:-1, 1:].ravel(), idx[1:, 1:].ravel()
    f = np.concatenate([np.stack([ll, lr, ur], 1), np.stack([ll, ur, ul], 1)], 0)
    return MiniMesh(v, f)
