"""Per-iteration sweep1 time over the bench window (iterations 0..24 from the start mesh, blocks of 5)."""
import sys, os, ctypes, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from ch_shrinkwrap_b200.mesh_conj_grad import ShrinkwrapMeshConjGrad
mesh, pts, sig, cfg = bench.build_workload('c3', 1234)
s_inv = (1.0 / sig.ravel()).astype(np.float32)
st = (ctypes.c_uint64 * 4)(); sg = (ctypes.c_double * 16)(); sm = ctypes.c_double()
for blk in range(5):
    cg = ShrinkwrapMeshConjGrad(mesh, pts); mesh.cg = cg
    row = []
    for it in range(5):
        cg._h.call('nw_set_profile', int(os.environ.get('NW_PROFILE', '1')))
        cg.search(pts, lams=[5.0], num_iters=1, sigma_inv=s_inv)
        cg._h.call('nw_get_traversal_stats', st)
        cg._h.call('nw_get_profile', sg, None, ctypes.byref(sm))
        row.append('%.2f (%3.0f)' % (sg[2], st[0] / len(pts)))
    print('block %d  sweep1 ms (tests/pt): ' % blk + '  '.join(row))
# same window, 5 iterations per call (what bench.py times)
mesh, pts, sig, cfg = bench.build_workload('c3', 1234)
for blk in range(5):
    cg = ShrinkwrapMeshConjGrad(mesh, pts); mesh.cg = cg
    cg._h.call('nw_set_profile', int(os.environ.get('NW_PROFILE', '1')))
    cg.search(pts, lams=[5.0], num_iters=5, sigma_inv=s_inv)
    cg._h.call('nw_get_profile', sg, None, ctypes.byref(sm))
    print('block %d (5 iterations per call): sweep1 %.2f ms/iter, search %.2f ms/iter, stages %s' % (blk, sg[2] / 5, sm.value / 5, [round(x / 5, 3) for x in sg]))
