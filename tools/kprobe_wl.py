"""kernel_probe for any workload: argv = workload kernel..."""
import sys, os, ctypes, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from ch_shrinkwrap_b200.mesh_conj_grad import ShrinkwrapMeshConjGrad
mesh, pts, sig, cfg = bench.build_workload(sys.argv[1], 1234)
cg = ShrinkwrapMeshConjGrad(mesh, pts); mesh.cg = cg
cg.search(pts, lams=[cfg['curvature_weight'] / 2.0], num_iters=6, sigma_inv=(1.0 / sig.ravel()).astype(np.float32))
ms = ctypes.c_float()
for name in sys.argv[2:]:
    cg._h.call('nw_bench_kernel', name.encode(), 10, ctypes.byref(ms)); print('%s %-12s %.4f ms' % (sys.argv[1], name, ms.value))
