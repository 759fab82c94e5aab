"""Cost of one CG iteration late in a fit (mesh close to the localisations): sweep1 ms and node tests per point after
N iterations in blocks of 5 (argv[1], default 60)."""
import sys, os, ctypes, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from ch_shrinkwrap_b200.mesh_conj_grad import ShrinkwrapMeshConjGrad
n = int(sys.argv[1]) if len(sys.argv) > 1 else 60
mesh, pts, sig, cfg = bench.build_workload('c3', 1234)
s_inv = (1.0 / sig.ravel()).astype(np.float32)
st = (ctypes.c_uint64 * 4)(); sg = (ctypes.c_double * 16)(); sm = ctypes.c_double()
for blk in range(n // 5):
    cg = ShrinkwrapMeshConjGrad(mesh, pts); mesh.cg = cg
    cg._h.call('nw_set_profile', int(os.environ.get('NW_PROFILE', '1')))
    cg.search(pts, lams=[5.0], num_iters=5, sigma_inv=s_inv)
    cg._h.call('nw_get_traversal_stats', st)
    cg._h.call('nw_get_profile', sg, None, ctypes.byref(sm))
    if blk % 2 == 1 or blk == n // 5 - 1:
        print('iters %3d..%3d: search %.2f ms/iter, sweep1 %.2f ms/iter, tests/pt (last iter) %.1f, mean |res| %.3f' % (
            5 * blk, 5 * blk + 4, sm.value / 5, sg[2] / 5, st[0] / len(pts), float(np.mean(cg.ress[-1:]))))
