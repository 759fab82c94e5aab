"""Final-mesh parity at BASELINE config 1 scale (1 M localisations, 50 412 vertices): GPU vs CPU oracle after 10 CG
iterations in two blocks of 5, plus a physical sanity check (distance of the fitted vertices to the true surface)."""
import copy, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from ch_shrinkwrap_b200 import synth
from ch_shrinkwrap_b200.mesh_conj_grad import ShrinkwrapMeshConjGrad
from oracle import nanowrap_oracle as orc

mesh, pts, sig, cfg = bench.build_workload('c2', 1234)
s_inv = (1.0 / sig.ravel()).astype(np.float32)
lam = 5.0
mo = copy.deepcopy(mesh)
shape = synth.two_lobed()
print('initial rms |sdf(v)| %.2f nm' % np.sqrt((shape.sdf(mesh.vertices.astype(np.float64)) ** 2).mean()))
t0 = time.perf_counter()
for blk in range(2):
    cg = ShrinkwrapMeshConjGrad(mesh, pts); mesh.cg = cg
    vg = cg.search(pts, lams=[lam], num_iters=5, sigma_inv=s_inv)
    mesh.update_geometry()
tg = time.perf_counter() - t0
t0 = time.perf_counter()
for blk in range(2):
    oc = orc.OracleConjGrad(mo, pts); mo.cg = oc
    vo = oc.search(pts, lams=[lam], num_iters=5, sigma_inv=s_inv)
    mo.update_geometry()
tc = time.perf_counter() - t0
disp = np.sqrt(((vg.astype(np.float64) - vo) ** 2).sum(1))
print('GPU %.2f s, CPU oracle %.1f s' % (tg, tc))
print('max vertex displacement GPU vs oracle after 10 iterations: %.3g nm (mean %.3g nm)' % (disp.max(), disp.mean()))
print('final rms |sdf(v)| GPU %.2f nm, oracle %.2f nm' % (np.sqrt((shape.sdf(vg.astype(np.float64)) ** 2).mean()),
                                                      np.sqrt((shape.sdf(vo.astype(np.float64)) ** 2).mean())))
same = np.all(cg.w[0] == oc.w[0], axis=1)
print('nearest faces of the last iteration identical for %.4f %% of the points' % (100 * same.mean()))
