import sys, os, ctypes, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from ch_shrinkwrap_b200.mesh_conj_grad import ShrinkwrapMeshConjGrad
mesh, pts, sig, cfg = bench.build_workload('c3', 1234)
s_inv=(1.0/sig.ravel()).astype(np.float32)
cg=ShrinkwrapMeshConjGrad(mesh, pts); mesh.cg=cg
cg.search(pts, lams=[5.0], num_iters=6, sigma_inv=s_inv)
counts=(ctypes.c_int*32)(); nl=ctypes.c_int()
cg._h.call('nw_debug_tree', -1, None, counts, ctypes.byref(nl))
print('levels', nl.value, list(counts)[:nl.value])
for lvl in range(1, min(nl.value,5)):
    n=counts[lvl]; b=np.zeros((n,16),np.float32)
    cg._h.call('nw_debug_tree', lvl, b.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), None, None)
    print('level', lvl)
    for k in range(min(n,8)):
        r=b[k]
        # Box layout (csrc/common.cuh): a = n.x t1.x n.y t1.y | b = n.z t1.z n-min t1-min | c = t2.xyz t2-max | d = n-max t1-max link t2-min
        print('  n=(%.2f %.2f %.2f) ext n[%.0f,%.0f] t1[%.0f,%.0f] t2[%.0f,%.0f]' % (r[0], r[2], r[4], r[6], r[12], r[7], r[13], r[15], r[11]))
