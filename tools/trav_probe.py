import sys, ctypes, time, numpy as np
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from ch_shrinkwrap_b200.mesh_conj_grad import ShrinkwrapMeshConjGrad
wl = sys.argv[1] if len(sys.argv)>1 else 'c3'
mesh, pts, sig, cfg = bench.build_workload(wl, 1234)
s_inv=(1.0/sig.ravel()).astype(np.float32)
cg=ShrinkwrapMeshConjGrad(mesh, pts); mesh.cg=cg
P=len(pts)
st=(ctypes.c_uint64*4)(); sg=(ctypes.c_double*16)(); cg._h.call('nw_set_profile', int(os.environ.get('NW_PROFILE', '1')))
sm=ctypes.c_double()
for it in range(8):
    if it == 4: cg._upload_topology()
    cg.search(pts, lams=[5.0], num_iters=1, sigma_inv=s_inv)
    cg._h.call('nw_get_traversal_stats', st)
    cg._h.call('nw_get_profile', sg, None, ctypes.byref(sm)); print('   stages', [round(x,2) for x in sg]); cg._h.call('nw_set_profile', int(os.environ.get('NW_PROFILE', '1')))
    print('iter',it,'ms %.2f'%sm.value,'tests/pt %.1f leaves/pt %.2f exact/pt %.2f max tests %d'%(st[0]/P, st[1]/P, st[2]/P, st[3]))
ms=ctypes.c_float()
for name in (b'nn_weights', b'sweep1', b'apply_AH'):
    cg._h.call('nw_bench_kernel', name, 5, ctypes.byref(ms)); print(name, ms.value)
