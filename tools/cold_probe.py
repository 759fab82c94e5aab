"""First block of bench.py's timed run: 5 warm-up iterations, positions reset to the start mesh, then per-iteration cost."""
import sys, os, ctypes, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from ch_shrinkwrap_b200.mesh_conj_grad import ShrinkwrapMeshConjGrad
mesh, pts, sig, cfg = bench.build_workload('c3', 1234)
s_inv = (1.0 / sig.ravel()).astype(np.float32)
start = mesh._vertices['position'].copy()
sg = (ctypes.c_double * 16)(); sm = ctypes.c_double()
for rep in range(2):
    mesh._vertices['position'][:] = start; mesh.update_geometry()
    cg = ShrinkwrapMeshConjGrad(mesh, pts); mesh.cg = cg
    row = []
    for it in range(5):
        cg._h.call('nw_set_profile', 1)
        cg.search(pts, lams=[5.0], num_iters=1, sigma_inv=s_inv)
        cg._h.call('nw_get_profile', sg, None, ctypes.byref(sm))
        row.append('%.2f+%.2f' % (sg[2], sg[8]))
    print(('cold (leaders)   ' if rep == 0 else 'feet of iteration 5') + ': sweep1+seed ms per iteration: ' + '  '.join(row))
