"""sweep1 time against the number of points at fixed density: slabs x <= quantile(f) of the C3 cloud."""
import sys, os, ctypes, copy, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from ch_shrinkwrap_b200.mesh_conj_grad import ShrinkwrapMeshConjGrad
mesh0, pts, sig, cfg = bench.build_workload('c3', 1234)
lam = cfg['curvature_weight'] / 2.0
def run(p, s, label):
    mesh = copy.deepcopy(mesh0)
    s_inv = (1.0 / s.ravel()).astype(np.float32)
    tot = {}
    for blk in range(2):
        cg = ShrinkwrapMeshConjGrad(mesh, p); mesh.cg = cg
        cg._h.call('nw_set_profile', 1)
        cg.search(p, lams=[lam], num_iters=cfg['block'], sigma_inv=s_inv)
        n = ctypes.c_int(0)
        st = (ctypes.c_int32 * 4096)(); ms = (ctypes.c_float * 4096)()
        cg._h.call('nw_get_stage_trace', st, ms, 4096, ctypes.byref(n))
        if blk == 1:
            per = [ms[k] for k in range(n.value) if st[k] == 2]
    print('%-12s P %9d  sweep1 per iteration: %s  -> warm %.4f ms per Mpt' % (label, len(p), ' '.join('%.3f' % v for v in per), per[-1] / (len(p) / 1e6)))
x = pts[:, 0]
for f in (1.0, 0.5, 0.25, 0.125, 0.0625):
    m = x <= np.quantile(x, f)
    run(pts[m], sig[m], 'slab %.4f' % f)
