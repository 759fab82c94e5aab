"""NVTX stage ranges (nw_set_profile bit 4): `ncu --nvtx --nvtx-include "adjoint/"` must select the kernels of that stage."""
import sys, os, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from ch_shrinkwrap_b200.mesh_conj_grad import ShrinkwrapMeshConjGrad
mesh, pts, sig, cfg = bench.build_workload('c2', 1234)
cg = ShrinkwrapMeshConjGrad(mesh, pts); mesh.cg = cg
cg._h.call('nw_set_profile', 4)
cg.search(pts, lams=[5.0], num_iters=1, sigma_inv=(1.0 / sig.ravel()).astype(np.float32))
print('done')
