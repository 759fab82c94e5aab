"""GPU versions of the reference's mesh-quality helpers (``ch_shrinkwrap/evaluation_utils.py:35-180``), same names,
arguments and return values:

* ``points_from_mesh(mesh, dx_min=5, p=1.0, return_normals=False)`` -- the per-triangle Python loop becomes one kernel
  (``nw_points_from_mesh``); the samples are bit-identical to the reference's ``d`` array.  The reference returns them in a
  random order (``np.random.choice`` without replacement, :137); here the order is random too, drawn from
  ``np.random`` exactly like the reference, so seeding ``np.random`` gives the same permutation for the same count.
* ``average_squared_distance(points0, points1)`` -- the two ``cKDTree`` builds + queries become two nearest-point
  searches in the solver's hierarchy (``nw_set_point_targets``); distances are the same float64 numbers.
* ``nearest_point_distance(targets, queries)`` -- the building block, also used by the hole-punch candidate search.

No CPU fallback: these raise if libnanowrap.so or the GPU is missing.
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib


def _xyz(a):
    a = np.asarray(a)
    if a.ndim != 2 or a.shape[1] != 3:
        raise ValueError('expected an (N,3) array')
    if a.dtype == np.float32:
        return np.ascontiguousarray(a), 0
    return np.ascontiguousarray(a, dtype=np.float64), 1


def nearest_point_distance(targets, queries, handle=None, return_index=False):
    """``scipy.spatial.cKDTree(targets).query(queries, k=1)``: float64 distances (and indices) of the nearest target."""
    t, t64 = _xyz(targets)
    q, q64 = _xyz(queries)
    h = handle if handle is not None else _lib.Handle(0)
    try:
        dist = np.empty(len(q), np.float64)
        idx = np.empty(len(q), np.int32)
        if len(q) == 0:
            return (dist, idx) if return_index else dist
        if len(t) == 0:
            raise ValueError('no target points')
        h.call('nw_set_point_targets', ctypes.c_void_p(t.ctypes.data), t64, len(t))
        h.call('nw_set_points', ctypes.c_void_p(q.ctypes.data), q64, len(q), None, 1.0, None)
        h.call('nw_compute_weights')
        h.call('nw_get_weights', None, None, _lib.dptr(dist), _lib.iptr(idx))
    finally:
        if handle is None:
            h.close()
    return (dist, idx) if return_index else dist


def average_squared_distance(points0, points1, handle=None):
    """evaluation_utils.py:147-180: (mean squared distance of points1 to their nearest points0, and the converse)."""
    points0_err = nearest_point_distance(points0, points1, handle)          # :175
    points1_err = nearest_point_distance(points1, points0, handle)          # :176
    points0_mse = np.nansum(points0_err ** 2) / len(points0_err)            # :178
    points1_mse = np.nansum(points1_err ** 2) / len(points1_err)            # :179
    return points0_mse, points1_mse


def mesh_samples(mesh, dx_min=5, handle=None):
    """The reference's ``d`` array (evaluation_utils.py:59-135) before its random permutation: (n,3) float64."""
    pos = _lib.as_f32(mesh._vertices['position'])
    faces = np.ascontiguousarray(mesh.faces, dtype=np.int32)
    h = handle if handle is not None else _lib.Handle(0)
    try:
        n = ctypes.c_int64(0)
        h.call('nw_points_from_mesh', _lib.fptr(pos), _lib.iptr(faces), len(pos), len(faces), float(dx_min), None, 0, ctypes.byref(n))
        d = np.empty((n.value, 3), np.float64)
        if n.value:
            h.call('nw_points_from_mesh', _lib.fptr(pos), _lib.iptr(faces), len(pos), len(faces), float(dx_min), _lib.dptr(d), n.value,
                   ctypes.byref(n))
    finally:
        if handle is None:
            h.close()
    return d


def points_from_mesh(mesh, dx_min=5, p=1.0, return_normals=False):
    """evaluation_utils.py:35-145."""
    h = _lib.Handle(0)
    try:
        d = mesh_samples(mesh, dx_min, h)
        subsamp = np.random.choice(np.arange(d.shape[0]), size=int(p * d.shape[0]), replace=False)      # :137
        if return_normals:
            # nearest face centroid of every sample (:143-149): the solver's own nearest-face search with the samples as points
            face_centers = mesh._vertices['position'][mesh.faces].mean(1)
            _, faces = nearest_point_distance(face_centers, d, h, return_index=True)
            normals = mesh._faces['normal'][mesh._faces['halfedge'] != -1][faces]
            return d[subsamp], normals[subsamp]
        return d[subsamp]
    finally:
        h.close()
