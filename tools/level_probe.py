import sys, os, ctypes, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from ch_shrinkwrap_b200.mesh_conj_grad import ShrinkwrapMeshConjGrad
mesh, pts, sig, cfg = bench.build_workload('c3', 1234)
s_inv=(1.0/sig.ravel()).astype(np.float32)
cg=ShrinkwrapMeshConjGrad(mesh, pts); mesh.cg=cg
cg.search(pts, lams=[5.0], num_iters=6, sigma_inv=s_inv)
st=(ctypes.c_uint64*4)()
cg.search(pts, lams=[5.0], num_iters=1, sigma_inv=s_inv)
cg._h.call('nw_get_traversal_stats', st)
print('tests/pt', st[0]/len(pts))
