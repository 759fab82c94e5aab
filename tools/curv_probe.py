"""Curvature through the C ABI with host buffers (what remove_necks / the recipe's last step call): wall time of repeated
calls and the kernel alone.  argv: workload"""
import sys, os, ctypes, time, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from ch_shrinkwrap_b200.membrane_mesh import curvature_grad
wl = sys.argv[1] if len(sys.argv) > 1 else 'c3'
mesh, pts, sig, cfg = bench.build_workload(wl, 1234, points=1000)
M = len(mesh._vertices)
out = None
for rep in range(4):
    t0 = time.perf_counter()
    out = curvature_grad(mesh, kc=1.0, out=out)
    print('%s: nw_curvature_grad call %d: %.2f ms (M = %d, %.1f MB up, %.1f MB down)' % (
        wl, rep, 1e3 * (time.perf_counter() - t0), M, (120 * M + 24 * len(mesh._faces) + 28 * len(mesh._halfedges)) / 1e6, 72 * M / 1e6))
ms = ctypes.c_float()
h = mesh._nw_session.handle
h.call('nw_bench_kernel', b'curvature', 10, ctypes.byref(ms))
print('%s: k_curvature %.4f ms  (%.0f M vertices/s)' % (wl, ms.value, M / ms.value / 1e3))
