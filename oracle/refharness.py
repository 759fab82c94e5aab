"""Run the UNMODIFIED reference (test infrastructure; this container only).

Imports ``mesh_conj_grad.py`` / ``conj_grad.py`` straight from /root/reference (nothing is
copied) and the reference's C compiled by ``oracle/build.py`` into oracle/_ref/.  Used by
``oracle/make_golden.py`` to write tests/golden/ and by tests marked ``needs_reference``.
/root/reference does not exist on the GPU box, so nothing under ``-m gpu`` may import this.

Import recipe: empty ``sys.modules`` stubs for
``numpy.compat.py3k`` (removed in numpy 2; imported but unused, mesh_conj_grad.py:7),
``PYME.experimental.isosurface`` (delaunay_utils.py:5) and ``PYME.experimental.octree``
(mesh_conj_grad.py:440).
"""
from __future__ import annotations

import ctypes
import importlib.machinery
import importlib.util
import os
import sys
import types

import numpy as np

from . import build as _build

REF_PKG_DIR = '/root/reference/ch_shrinkwrap'


def available():
    return os.path.isdir(REF_PKG_DIR)


_mods = {}


def load_reference():
    """Returns (mesh_conj_grad module, conj_grad_utils module)."""
    if 'mcg' in _mods:
        return _mods['mcg'], _mods['cgu']
    if not available():
        raise RuntimeError('/root/reference is not present on this machine')
    cg_so, _ = _build.build_ref()

    def stub(name, **attrs):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__dict__.update(attrs)
            sys.modules[name] = m
        return sys.modules[name]

    stub('numpy.compat', py3k=None)
    stub('numpy.compat.py3k', npy_load_module=None)
    stub('PYME'); stub('PYME.experimental')
    stub('PYME.experimental.isosurface'); stub('PYME.experimental.octree')
    sys.modules['PYME.experimental'].octree = sys.modules['PYME.experimental.octree']
    sys.modules['PYME.experimental'].isosurface = sys.modules['PYME.experimental.isosurface']

    pkg = types.ModuleType('ch_shrinkwrap')
    pkg.__path__ = [REF_PKG_DIR]
    sys.modules['ch_shrinkwrap'] = pkg
    loader = importlib.machinery.ExtensionFileLoader('ch_shrinkwrap.conj_grad_utils', cg_so)
    spec = importlib.util.spec_from_file_location('ch_shrinkwrap.conj_grad_utils', cg_so, loader=loader)
    cgu = importlib.util.module_from_spec(spec)
    loader.exec_module(cgu)
    sys.modules['ch_shrinkwrap.conj_grad_utils'] = cgu
    pkg.conj_grad_utils = cgu
    import ch_shrinkwrap.mesh_conj_grad as mcg  # noqa: E402  (the reference file, unmodified)
    _mods['mcg'], _mods['cgu'] = mcg, cgu
    return mcg, cgu


def reference_solver(mesh, points):
    """ShrinkwrapMeshConjGrad(mesh, points) from the reference; sets mesh.cg like
    MembraneMesh.opt_conjugate_gradient does (_membrane_mesh.pyx:1510)."""
    mcg, _ = load_reference()
    cg = mcg.ShrinkwrapMeshConjGrad(mesh, points)
    mesh.cg = cg
    return cg


_libc = ctypes.CDLL(None)
_libc.rand.restype = ctypes.c_int


def libc_uniforms(seed, n):
    """The doubles ``dr2()`` (membrane_mesh_utils.c:428-431) yields after ``srand(seed)``."""
    _libc.srand(ctypes.c_uint(seed))
    return np.array([_libc.rand() for _ in range(n)], dtype=np.float64) / 2147483648.0


def reference_curvature(mesh, dN=0.1, skip_prob=0.0, kc=1.0, kg=-20.0 * 0.0257, c0=0.0, seed=None):
    """Reference ``c_curvature_grad`` (membrane_mesh_utils.c:915) on the mesh's structured arrays.
    ``seed``: call libc ``srand(seed)`` first so the jitter draws are reproducible."""
    _, so = _build.build_ref()
    lib = ctypes.PyDLL(so)
    from .nanowrap_oracle import CURV_SCALARS, CURV_VECTORS
    M = len(mesh._vertices)
    out = {k: np.zeros(M, np.float32) for k in CURV_SCALARS}
    out.update({k: np.zeros((M, 3), np.float32) for k in CURV_VECTORS})
    verts = np.ascontiguousarray(mesh._vertices)
    faces = np.ascontiguousarray(mesh._faces)
    hes = np.ascontiguousarray(mesh._halfedges)
    if seed is not None:
        _libc.srand(ctypes.c_uint(seed))
    fp = ctypes.POINTER(ctypes.c_float)
    lib.ref_curvature_grad(
        ctypes.c_void_p(verts.ctypes.data), ctypes.c_void_p(faces.ctypes.data), ctypes.c_void_p(hes.ctypes.data),
        ctypes.c_float(dN), ctypes.c_float(skip_prob), ctypes.c_int(M),
        *[out[k].ctypes.data_as(fp) for k in ('k0', 'k1', 'e0', 'e1', 'H', 'K', 'dH', 'dK', 'E', 'pE', 'dE_neighbors')],
        ctypes.c_float(kc), ctypes.c_float(kg), ctypes.c_float(c0), out['dEdN'].ctypes.data_as(fp))
    return out
