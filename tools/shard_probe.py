"""Per-point sweep cost of ONE rank's share of an N-way split on one GPU: full cloud vs rank 0 of an interleaved split
(several cube sizes) vs a random 1/N subset.  argv: workload world"""
import sys, os, ctypes, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from ch_shrinkwrap_b200.mesh_conj_grad import ShrinkwrapMeshConjGrad
from ch_shrinkwrap_b200 import sharding
wl = sys.argv[1] if len(sys.argv) > 1 else 'c3'
world = int(sys.argv[2]) if len(sys.argv) > 2 else 8
mesh0, pts, sig, cfg = bench.build_workload(wl, 1234)
lam = cfg['curvature_weight'] / 2.0
import copy


def run(p, s, label):
    mesh = copy.deepcopy(mesh0)
    s_inv = (1.0 / s.ravel()).astype(np.float32)
    tot = {}
    for blk in range(2):
        cg = ShrinkwrapMeshConjGrad(mesh, p); mesh.cg = cg
        cg._h.call('nw_set_profile', 1)
        cg.search(p, lams=[lam], num_iters=cfg['block'], sigma_inv=s_inv)
        n = ctypes.c_int(0)
        st = (ctypes.c_int32 * 4096)(); ms = (ctypes.c_float * 4096)()
        cg._h.call('nw_get_stage_trace', st, ms, 4096, ctypes.byref(n))
        if blk == 1:
            for k in range(n.value):
                tot[st[k]] = tot.get(st[k], 0.0) + ms[k]
    it = cfg['block']
    print('%-28s P %9d  sweep1 %.3f ms/iter = %.4f ms per Mpt   adjoint %.3f  sweep2 %.3f' % (
        label, len(p), tot.get(2, 0) / it, tot.get(2, 0) / it / (len(p) / 1e6), tot.get(10, 0) / it, tot.get(5, 0) / it))


run(pts, sig, 'full cloud')
for tb in (4096,):
    m = sharding.interleaved_shard(pts, world, 0, target_block=tb)
    run(pts[m], sig[m], 'interleaved 1/%d, block %d' % (world, tb))
rng = np.random.default_rng(0)
m = rng.random(len(pts)) < 1.0 / world
run(pts[m], sig[m], 'random 1/%d' % world)
