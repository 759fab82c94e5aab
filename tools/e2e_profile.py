"""cProfile of bench.run_blocks (the e2e leg): where the host time of a block goes."""
import sys, os, time, ctypes, cProfile, pstats, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from ch_shrinkwrap_b200 import mesh_conj_grad as mcg
mesh, pts, sig, cfg = bench.build_workload('c3', 1234)
s_inv = (1.0 / sig.ravel()).astype(np.float32)
mcg._session_for(mesh)
bench.run_blocks(mesh, pts, s_inv, 5.0, 10, 5)
t0 = time.perf_counter()
pr = cProfile.Profile(); pr.enable()
dev_ms, wall, cg = bench.run_blocks(mesh, pts, s_inv, 5.0, 20, 5)
pr.disable()
print('20 iterations: wall %.1f ms, device (nw_search) %.1f ms' % (1e3 * (time.perf_counter() - t0), dev_ms))
pstats.Stats(pr).sort_stats('tottime').print_stats(12)
# wall time per C entry point
import collections
from ch_shrinkwrap_b200 import _lib
acc = collections.defaultdict(lambda: [0, 0.0])
orig = _lib.Handle.call
def timed(self, name, *a):
    t = time.perf_counter()
    try:
        return orig(self, name, *a)
    finally:
        acc[name][0] += 1; acc[name][1] += time.perf_counter() - t
_lib.Handle.call = timed
dev_ms, wall, cg = bench.run_blocks(mesh, pts, s_inv, 5.0, 20, 5)
for k, (n, t) in sorted(acc.items(), key=lambda kv: -kv[1][1]):
    print('%-28s %2d calls %7.1f ms' % (k, n, 1e3 * t))
print('device inside nw_search %.1f ms' % dev_ms)
