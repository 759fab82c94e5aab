"""Device time of a 5-iteration nw_search (first to last kernel, CUDA events) with and without the CUDA-graph replay
(NW_NO_GRAPH=1), small workloads."""
import sys, os, ctypes, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from ch_shrinkwrap_b200.mesh_conj_grad import ShrinkwrapMeshConjGrad
for wl in sys.argv[1:] or ['c1', 'c2']:
    mesh, pts, sig, cfg = bench.build_workload(wl, 1234)
    s_inv = (1.0 / sig.ravel()).astype(np.float32)
    sm = ctypes.c_double(); ts = []
    for blk in range(6):
        cg = ShrinkwrapMeshConjGrad(mesh, pts); mesh.cg = cg
        cg.search(pts, lams=[5.0], num_iters=5, sigma_inv=s_inv)
        cg._h.call('nw_get_profile', None, None, ctypes.byref(sm)); ts.append(sm.value)
    print('%s graph=%s: nw_search device ms per block of 5 (blocks 2..5): %s' % (wl, 'off' if os.environ.get('NW_NO_GRAPH') else 'on', ' '.join('%.3f' % t for t in ts[2:])))
