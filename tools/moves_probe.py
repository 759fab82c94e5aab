"""How far the face centroids move per CG iteration over a long fit at C3 (1-iteration search() calls, re-upload every 5), and
how many points the bound check of k_sweep1_fast cannot settle.  argv: workload n_iterations"""
import sys, os, ctypes, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from ch_shrinkwrap_b200.mesh_conj_grad import ShrinkwrapMeshConjGrad
wl = sys.argv[1] if len(sys.argv) > 1 else 'c3'
n_it = int(sys.argv[2]) if len(sys.argv) > 2 else 60
mesh, pts, sig, cfg = bench.build_workload(wl, 1234)
s_inv = (1.0 / sig.ravel()).astype(np.float32)
lam = cfg['curvature_weight'] / 2.0
P = len(pts)
NC = 299593
raw = np.zeros(NC, np.float32); dil = np.zeros(NC, np.float32)
fp = lambda a: a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))
sg = (ctypes.c_double * 16)()
listed = np.zeros(1)
cg = None
for it in range(n_it):
    if it % cfg['block'] == 0:
        cg = ShrinkwrapMeshConjGrad(mesh, pts); mesh.cg = cg
    cg._h.call('nw_set_profile', 1)
    cg.search(pts, lams=[lam], num_iters=1, sigma_inv=s_inv)
    cg._h.call('nw_get_profile', sg, None, None)
    cg._h.call('nw_get_search_counts', listed.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), 1)
    cg._h.call('nw_debug_dgrid', fp(raw), fp(dil), None)
    l6 = raw[37449:]; l6 = l6[l6 > 0]
    q = np.percentile(l6, [50, 90, 99]) if len(l6) else [0, 0, 0]
    d6 = dil[37449:]; d6 = d6[d6 > 0]
    qd = np.percentile(d6, [50, 90]) if len(d6) else [0, 0]
    print('it %3d: searched %6s  sweep1 full %.2f fast %.2f list %.2f refit %.2f | moved nm: global max %.3f, level-6 cells p50 %.4f p90 %.4f p99 %.4f, dilated p50 %.4f p90 %.4f' % (
        it, 'all' if listed[0] < 0 else '%.1f%%' % (100 * listed[0] / P), sg[2], sg[10], sg[12], sg[0], raw[0], q[0], q[1], q[2], qd[0], qd[1]), flush=True)
