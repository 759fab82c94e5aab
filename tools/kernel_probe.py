"""Isolated kernel timings (nw_bench_kernel) after a few iterations at C3."""
import sys, os, ctypes, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from ch_shrinkwrap_b200.mesh_conj_grad import ShrinkwrapMeshConjGrad
mesh, pts, sig, cfg = bench.build_workload('c3', 1234)
s_inv = (1.0 / sig.ravel()).astype(np.float32)
cg = ShrinkwrapMeshConjGrad(mesh, pts); mesh.cg = cg
cg.search(pts, lams=[5.0], num_iters=6, sigma_inv=s_inv)
ms = ctypes.c_float()
if 'curvature' in sys.argv[1:]:
    from ch_shrinkwrap_b200.membrane_mesh import curvature_grad
    mesh.update_geometry(); curvature_grad(mesh, kc=1.0)
for name in sys.argv[1:] or ['sweep2', 'mesh_prior', 'refit', 'apply_A', 'apply_AH', 'sweep1']:
    cg._h.call('nw_bench_kernel', name.encode(), 10, ctypes.byref(ms)); print('%-12s %.4f ms' % (name, ms.value))
