"""torchrun probe: per-rank stage times of the hot loop (is the accumulator allreduce slow, or is it waiting for the
slowest rank?) and the collective alone (all ranks in lockstep), plus the same payload through torch.distributed."""
import sys, os, ctypes, time, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import bench
from ch_shrinkwrap_b200 import _lib
rank = int(os.environ['RANK']); local = int(os.environ['LOCAL_RANK']); world = int(os.environ['WORLD_SIZE'])
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
uid = ctypes.create_string_buffer(128)
if rank == 0:
    assert _lib.load().nw_comm_unique_id(uid) == 0
t = torch.frombuffer(bytearray(uid.raw), dtype=torch.uint8).cuda()
dist.broadcast(t, 0)
comm = (rank, world, bytes(t.cpu().numpy().tobytes()))
mesh, pts, sig, cfg = bench.build_workload('c3', 1234 + 1000 * rank)
mesh._nw_device = local; mesh._nw_comm = comm
s_inv = (1.0 / sig.ravel()).astype(np.float32)
bench.run_blocks(mesh, pts, s_inv, 5.0, 5, 5)
h = mesh._nw_session.handle
h.call('nw_set_profile', 1)
dist.barrier()
bench.run_blocks(mesh, pts, s_inv, 5.0, 10, 5)
sg = (ctypes.c_double * 16)()
h.call('nw_get_profile', sg, None, None)
ms = ctypes.c_float()
h.call('nw_bench_kernel', b'allreduce_acc', 20, ctypes.byref(ms))
x = torch.zeros(4 * len(mesh._vertices), dtype=torch.int64, device='cuda')
for _ in range(3): dist.all_reduce(x)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(20): dist.all_reduce(x)
torch.cuda.synchronize(); tt = (time.perf_counter() - t0) / 20 * 1e3
for r in range(world):
    dist.barrier()
    if r == rank:
        print('rank %d: sweep1 %.2f ms/iter  allreduce_acc stage %.2f ms/iter | collective alone: library %.3f ms, torch %.3f ms (%.1f MB)' % (
            rank, sg[2] / 10, sg[3] / 10, ms.value, tt, x.numel() * 8 / 1e6), flush=True)
dist.destroy_process_group()
