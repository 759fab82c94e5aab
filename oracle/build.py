"""Build recipes for the oracle (test infrastructure).

* ``build_oracle_c()``  -> oracle/liboracle.so from oracle/oracle_c.c (our C restatement).
* ``build_ref()``       -> oracle/_ref/ : the reference's own C compiled from the sources where
  they lie under /root/reference: ``conj_grad_utils`` as the CPython extension it is, and
  ``c_curvature_grad`` / ``c_holepunch_pair_candidate_faces`` through oracle/ref_curvature_wrapper.c;
  plus the reference's Python solver path (the modules ``mesh_conj_grad`` imports), staged byte for
  byte as a BUILD PRODUCT into oracle/_ref/ch_shrinkwrap/ so that ``bench.py --impl reference``
  and the CPU-baseline leg can run the unmodified reference on the GPU box, where /root/reference
  does not exist.  oracle/_ref/ is git-ignored (never part of the repository's history) and travels
  with gpurun like any other built file.  Skipped when /root/reference is absent (the GPU box uses
  the prebuilt files).
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.environ.get('NW_REFERENCE_SRC', '/root/reference/ch_shrinkwrap')   # the override exists to test the staged-copy path
REF_OUT = os.path.join(HERE, '_ref')
REF_PY_OUT = os.path.join(REF_OUT, 'ch_shrinkwrap')
# what `import ch_shrinkwrap.mesh_conj_grad` / `evaluation_utils` pull in from the package (mesh_conj_grad.py:8,18;
# delaunay_utils.py:7; evaluation_utils.py:14-15,30)
REF_PY_FILES = ('mesh_conj_grad.py', 'conj_grad.py', 'delaunay_utils.py', 'sdf.py', 'util.py', 'evaluation_utils.py', 'shape.py')


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
    os.makedirs(REF_PY_OUT, exist_ok=True)
    for name in REF_PY_FILES:
        src, dst = os.path.join(REF_SRC, name), os.path.join(REF_PY_OUT, name)
        if force or _stale(dst, [src]):
            shutil.copyfile(src, dst)
    return cg, curv


def ref_python_dir():
    """Directory holding the reference's unmodified solver modules: the live tree when present, else the staged copy."""
    if os.path.isdir(REF_SRC):
        return REF_SRC
    if all(os.path.exists(os.path.join(REF_PY_OUT, n)) for n in REF_PY_FILES):
        return REF_PY_OUT
    return None


if __name__ == '__main__':
    print(build_oracle_c(force='-f' in sys.argv))
    print(build_ref(force='-f' in sys.argv))
