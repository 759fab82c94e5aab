"""CPU only: how far apart are the reference algorithm with float32 Gram sums (as shipped) and with float64 Gram sums
after 10 iterations at config-1 scale?  This is the reference's own numerical noise floor for the final mesh."""
import copy, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from oracle import nanowrap_oracle as orc
mesh, pts, sig, cfg = bench.build_workload('c2', 1234)
s_inv = (1.0 / sig.ravel()).astype(np.float32)
out = []
for cls in (orc.OracleConjGrad, orc.OracleConjGrad64):
    m = copy.deepcopy(mesh)
    for blk in range(2):
        oc = cls(m, pts); m.cg = oc
        v = oc.search(pts, lams=[5.0], num_iters=5, sigma_inv=s_inv)
        m.update_geometry()
    out.append(v.astype(np.float64))
d = np.sqrt(((out[0] - out[1]) ** 2).sum(1))
print('float32-Gram reference path vs float64-Gram: max %.3g nm, mean %.3g nm, 99.9th pct %.3g nm' % (d.max(), d.mean(), np.percentile(d, 99.9)))
