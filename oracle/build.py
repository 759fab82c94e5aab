"""Build recipes for the oracle (test infrastructure).

* ``build_oracle_c()``  -> oracle/liboracle.so from oracle/oracle_c.c (our C restatement).
* ``build_ref()``       -> oracle/_ref/ : the reference's own C compiled from the sources where
  they lie under /root/reference (never copied): ``conj_grad_utils`` as the CPython extension it
  is, and ``c_curvature_grad`` through oracle/ref_curvature_wrapper.c.  Skipped when
  /root/reference is absent (the GPU box uses the prebuilt files that travel with the snapshot).
"""
from __future__ import annotations

import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = '/root/reference/ch_shrinkwrap'
REF_OUT = os.path.join(HERE, '_ref')


def _run(cmd):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError('oracle build failed: %s\n%s' % (' '.join(cmd), r.stderr[-4000:]))


def _stale(out, srcs):
    return (not os.path.exists(out)) or any(os.path.getmtime(s) > os.path.getmtime(out) for s in srcs)


def build_oracle_c(force=False):
    src = os.path.join(HERE, 'oracle_c.c')
    out = os.path.join(HERE, 'liboracle.so')
    if force or _stale(out, [src]):
        # no -march: like the reference build, no FMA contraction
        _run(['gcc', '-O2', '-fPIC', '-shared', '-ffp-contract=off', '-std=c99', src, '-o', out, '-lm'])
    return out


def ref_paths():
    ext = sysconfig.get_config_var('EXT_SUFFIX')
    return (os.path.join(REF_OUT, 'conj_grad_utils' + ext), os.path.join(REF_OUT, 'libref_curvature.so'))


def build_ref(force=False):
    """Compile the reference's C from /root/reference into oracle/_ref/. Returns paths or None."""
    if not os.path.isdir(REF_SRC):
        return None
    import numpy
    os.makedirs(REF_OUT, exist_ok=True)
    inc = ['-I' + sysconfig.get_paths()['include'], '-I' + numpy.get_include(), '-I' + REF_SRC]
    cg, curv = ref_paths()
    src_cg = os.path.join(REF_SRC, 'conj_grad_utils.c')
    if force or _stale(cg, [src_cg]):
        _run(['gcc', '-O2', '-fPIC', '-shared', '-w'] + inc + [src_cg, '-o', cg, '-lm'])
    wrap = os.path.join(HERE, 'ref_curvature_wrapper.c')
    if force or _stale(curv, [wrap, os.path.join(REF_SRC, 'membrane_mesh_utils.c')]):
        _run(['gcc', '-O2', '-fPIC', '-shared', '-w'] + inc + [wrap, '-o', curv, '-lm'])
    return cg, curv


if __name__ == '__main__':
    print(build_oracle_c(force='-f' in sys.argv))
    print(build_ref(force='-f' in sys.argv))
