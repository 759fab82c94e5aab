"""The C-ABI library loads and exports every symbol include/nanowrap.h declares (no compute calls: no GPU here)."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    txt = open(os.path.join(ROOT, 'include', 'nanowrap.h')).read()
    txt = re.sub(r'/\*.*?\*/', '', txt, flags=re.S)
    return sorted(set(re.findall(r'\b(nw_[a-z_0-9A-Z]+)\s*\(', txt)))


def test_library_builds_and_exports_every_declared_symbol():
    from ch_shrinkwrap_b200 import build
    lib = build.build()
    out = subprocess.run(['nm', '-D', '--defined-only', lib], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r' T (nw_\w+)', out))
    decl = declared_symbols()
    assert len(decl) >= 30
    missing = [s for s in decl if s not in exported]
    assert not missing, 'declared in nanowrap.h but not exported: %s' % missing


def test_ctypes_signatures_cover_the_header():
    from ch_shrinkwrap_b200 import _lib
    lib = _lib.load()
    assert lib.nw_version() >= 100
    decl = set(declared_symbols())
    assert decl == set(_lib.SIGNATURES), decl ^ set(_lib.SIGNATURES)


def test_product_path_fails_loudly_without_gpu():
    import ctypes
    from ch_shrinkwrap_b200 import _lib
    n = ctypes.c_int(0)
    try:
        cudart_has_gpu = subprocess.run(['nvidia-smi', '-L'], capture_output=True).returncode == 0
    except FileNotFoundError:
        cudart_has_gpu = False
    if cudart_has_gpu:
        pytest.skip('a GPU is present')
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        _lib.Handle(0)


def test_sass_is_sm100a_and_has_no_float_atomics_in_the_adjoint():
    """Cheap static evidence: the cubin targets sm_100a, the adjoint kernels use 64-bit integer REDs only and form their
    warp-level sums on the tensor cores."""
    from ch_shrinkwrap_b200 import build
    lib = build.build()
    r = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True)
    if r.returncode != 0:
        pytest.skip('cuobjdump unavailable')
    assert 'sm_100a' in r.stdout
    blocks = r.stdout.split('Function : ')
    adj = [b for b in blocks if 'k_adjoint' in b.split('\n')[0]]
    assert adj
    for b in adj:
        lines = b.split('\n')
        assert '.F32' not in ''.join(l for l in lines if 'RED' in l or 'ATOM' in l), 'float atomic in adjoint kernel'
        assert any(('RED' in l or 'ATOM' in l) and '.64' in l for l in lines), lines[0]
        # the per-face sums of a warp are a u8 integer matrix product fed by the byte-transposing ldmatrix
        assert any('IMMA.16832.U8.U8' in l for l in lines), lines[0]
        assert any('LDSM' in l and 'MT1616' in l for l in lines), lines[0]
        assert not any('REDUX' in l for l in lines), lines[0]
