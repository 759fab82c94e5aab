"""Does the floor of a small sweep come from the few deepest queries?  Slab 1/16 of the C3 cloud with and without the
points whose nearest-face distance (after 10 iterations) exceeds a threshold."""
import sys, os, ctypes, copy, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from ch_shrinkwrap_b200.mesh_conj_grad import ShrinkwrapMeshConjGrad
mesh0, pts, sig, cfg = bench.build_workload('c3', 1234)
lam = cfg['curvature_weight'] / 2.0
def run(p, s, label, want_d=False):
    mesh = copy.deepcopy(mesh0)
    s_inv = (1.0 / s.ravel()).astype(np.float32)
    for blk in range(2):
        cg = ShrinkwrapMeshConjGrad(mesh, p); mesh.cg = cg
        cg._h.call('nw_set_profile', 1)
        cg.search(p, lams=[lam], num_iters=cfg['block'], sigma_inv=s_inv)
        n = ctypes.c_int(0)
        st = (ctypes.c_int32 * 4096)(); ms = (ctypes.c_float * 4096)()
        cg._h.call('nw_get_stage_trace', st, ms, 4096, ctypes.byref(n))
        if blk == 1:
            per = [ms[k] for k in range(n.value) if st[k] == 2]
    print('%-34s P %9d  sweep1: %s' % (label, len(p), ' '.join('%.3f' % v for v in per)))
    if want_d:
        return np.asarray(cg.d).reshape(len(p), -1)[:, 0] / 3.0
x = pts[:, 0]
m = x <= np.quantile(x, 0.0625)
p, s = pts[m], sig[m]
d = run(p, s, 'slab 1/16', True)
print('distance quantiles 50/90/99/99.9/max: %s' % np.round(np.quantile(d, [0.5, 0.9, 0.99, 0.999, 1.0]), 1))
for thr in (200.0, 100.0, 50.0):
    k = d <= thr
    run(p[k], s[k], 'without d > %.0f nm (%d dropped)' % (thr, np.count_nonzero(~k)))
