"""Run the UNMODIFIED reference (test infrastructure).

Imports ``mesh_conj_grad.py`` / ``conj_grad.py`` straight from /root/reference when it exists (this
container), else from the byte-for-byte build product ``oracle/build.py`` staged under
oracle/_ref/ch_shrinkwrap/ (the GPU box), together with the reference's C compiled into oracle/_ref/.
Used by ``oracle/make_golden.py`` to write tests/golden/, by tests marked ``needs_reference`` and by
``bench.py``'s reference arm / CPU-baseline leg -- never by the product path.

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

def ref_pkg_dir():
    return _build.ref_python_dir()


def available():
    """True if the unmodified reference can be imported here (live tree or staged build product + its compiled C)."""
    if ref_pkg_dir() is None:
        return False
    return os.path.isdir(_build.REF_SRC) or all(os.path.exists(p) for p in _build.ref_paths())


def kind():
    return 'live tree /root/reference' if os.path.isdir(_build.REF_SRC) else 'staged copy oracle/_ref/ch_shrinkwrap'


_mods = {}


def load_reference():
    """Returns (mesh_conj_grad module, conj_grad_utils module)."""
    if 'mcg' in _mods:
        return _mods['mcg'], _mods['cgu']
    if not available():
        raise RuntimeError('the reference is neither at /root/reference nor staged under oracle/_ref/ (run oracle/build.py where it is)')
    cg_so = (_build.build_ref() or _build.ref_paths())[0]

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
    pkg.__path__ = [ref_pkg_dir()]
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
    so = (_build.build_ref() or _build.ref_paths())[1]
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


def reference_holepunch_pairs(mesh, candidates):
    """Reference ``c_holepunch_pair_candidate_faces`` (membrane_mesh_utils.c:1301-1379) on the mesh's structured arrays:
    returns the raw ``pairs`` array (index into candidates, -1 = none), initialised to -1 like _membrane_mesh.pyx:899."""
    so = (_build.build_ref() or _build.ref_paths())[1]
    lib = ctypes.PyDLL(so)
    verts = np.ascontiguousarray(mesh._vertices)
    faces = np.ascontiguousarray(mesh._faces)
    hes = np.ascontiguousarray(mesh._halfedges)
    cand = np.ascontiguousarray(candidates, dtype=np.int32)
    pairs = -1 * np.ones(len(cand), np.int32)
    ip = ctypes.POINTER(ctypes.c_int)
    lib.ref_holepunch_pair_candidate_faces(ctypes.c_void_p(verts.ctypes.data), ctypes.c_void_p(faces.ctypes.data),
                                           ctypes.c_void_p(hes.ctypes.data), cand.ctypes.data_as(ip), ctypes.c_int(len(cand)),
                                           pairs.ctypes.data_as(ip))
    return pairs


def load_evaluation_utils():
    """The reference's ``evaluation_utils`` module, unmodified (points_from_mesh, average_squared_distance are pure
    numpy/scipy; the module-level imports of the compiled mesh class and of PYME's simulator are stubbed)."""
    if 'evu' in _mods:
        return _mods['evu']
    load_reference()

    def stub(name):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
        return sys.modules[name]

    stub('ch_shrinkwrap._membrane_mesh')
    sys.modules['ch_shrinkwrap']._membrane_mesh = sys.modules['ch_shrinkwrap._membrane_mesh']
    stub('PYME.simulation'); stub('PYME.simulation.locify').points_from_sdf = None
    sys.modules['PYME'].simulation = sys.modules['PYME.simulation']
    sys.modules['PYME.simulation'].locify = sys.modules['PYME.simulation.locify']
    import ch_shrinkwrap.evaluation_utils as evu  # noqa: E402  (the reference file, unmodified)
    _mods['evu'] = evu
    return evu
