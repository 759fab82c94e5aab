"""Replays bench.py's e2e leg with per-block wall times (constructor / search) to see where the end-to-end time goes."""
import sys, os, time, ctypes, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from ch_shrinkwrap_b200 import mesh_conj_grad as mcg
mesh, pts, sig, cfg = bench.build_workload('c3', 1234)
s_inv = (1.0 / sig.ravel()).astype(np.float32)
start_pos = mesh._vertices['position'].copy()
mcg._session_for(mesh)
bench.run_blocks(mesh, pts, s_inv, 5.0, 5, 5)
for rep in range(2):
    mesh._vertices['position'][:] = start_pos
    mesh.update_geometry()
    mesh._nw_session.points_key = None
    T0 = time.perf_counter()
    for blk in range(4):
        t0 = time.perf_counter(); cg = mcg.ShrinkwrapMeshConjGrad(mesh, pts); mesh.cg = cg; t1 = time.perf_counter()
        cg._sigma_inv = s_inv
        cg._session.set_points(pts, s_inv, None); t2 = time.perf_counter()
        cg._upload_topology(); t3 = time.perf_counter()
        cg.search(pts, lams=[5.0], num_iters=5, sigma_inv=s_inv); t4 = time.perf_counter()
        sm = ctypes.c_double(); cg._h.call('nw_get_profile', None, None, ctypes.byref(sm))
        print('rep %d block %d: ctor %.1f set_points %.1f topo %.1f search %.1f (device %.1f) ms' % (rep, blk, 1e3*(t1-t0), 1e3*(t2-t1), 1e3*(t3-t2), 1e3*(t4-t3), sm.value))
    print('total %.1f ms' % (1e3 * (time.perf_counter() - T0)))
