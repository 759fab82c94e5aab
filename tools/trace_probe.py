"""Bench window (first N iterations of a fit, blocks as bench.py runs them) with the per-iteration cost of every stage
(nw_get_stage_trace).  argv: workload blocks"""
import sys, os, ctypes, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from ch_shrinkwrap_b200.mesh_conj_grad import ShrinkwrapMeshConjGrad
wl = sys.argv[1] if len(sys.argv) > 1 else 'c3'
nblk = int(sys.argv[2]) if len(sys.argv) > 2 else 4
mesh, pts, sig, cfg = bench.build_workload(wl, 1234)
s_inv = (1.0 / sig.ravel()).astype(np.float32)
lam = cfg['curvature_weight'] / 2.0
P = len(pts)
names = bench.STAGES
tot = 0.0
for blk in range(nblk):
    cg = ShrinkwrapMeshConjGrad(mesh, pts); mesh.cg = cg
    cg._h.call('nw_set_profile', int(os.environ.get('NW_PROFILE', '1')))
    cg.search(pts, lams=[lam], num_iters=cfg['block'], sigma_inv=s_inv)
    n = ctypes.c_int(0)
    st = (ctypes.c_int32 * 4096)(); ms = (ctypes.c_float * 4096)()
    cg._h.call('nw_get_stage_trace', st, ms, 4096, ctypes.byref(n))
    sm = ctypes.c_double(); cg._h.call('nw_get_profile', None, None, ctypes.byref(sm))
    tot += sm.value
    rows, cur = [], {}
    for k in range(n.value):
        if st[k] == 1 and cur:          # 'shift' opens an iteration
            rows.append(cur); cur = {}
        cur[st[k]] = cur.get(st[k], 0.0) + ms[k]
    rows.append(cur)
    print('block %d: search %.2f ms' % (blk, sm.value))
    for it, r in enumerate(rows):
        print('   it %d: sweep1 %.3f refit %.3f sweep2 %.3f prior %.3f solve %.3f seeds %.3f shift %.3f' % (
            it, r.get(2, 0), r.get(0, 0), r.get(5, 0), r.get(4, 0), r.get(7, 0), r.get(8, 0), r.get(1, 0)))
print('total search ms %.2f over %d iterations: %.3f ms/iter' % (tot, nblk * cfg['block'], tot / (nblk * cfg['block'])))
import zlib
print('crc of final vertices %08x' % zlib.crc32(np.ascontiguousarray(mesh._vertices['position']).tobytes()))
