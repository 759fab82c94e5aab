"""How to split one localisation cloud over several GPUs (points sharded, mesh replicated).

The solver accepts any partition -- results do not depend on it (the adjoint accumulates in integers) -- but speed does:
the nearest-face search walks the tree once per WARP of 32 neighbouring points, so a rank should hold points that are
dense where it holds any.  A random 1/N subset of the cloud is N times sparser everywhere (measured at BASELINE config 3,
100 M points on 8 GPUs: 0.30 ms per million points and sweep against 0.21 on one GPU); a contiguous range of a
space-filling curve is dense but gives one rank the deep, expensive queries and another the cheap ones.  Hence
**interleaved spatial blocks**: the bounding box is cut into cubes holding a few thousand points each, the cubes are numbered
along a Morton curve, and cube k belongs to rank k mod N -- every rank sees full density inside its cubes and the same mix
of regions.
"""
from __future__ import annotations

import numpy as np


def block_ids(points, lo, hi, n_points_total, target_block=512):
    """Index (x + c (y + c z)) of the cube each point lies in, and c = cubes per axis.  The cube size is chosen so that a
    cube of the bounding box [lo, hi] would hold about `target_block` points if the cloud filled the box uniformly (a
    surface cloud fills far fewer cubes, so the occupied ones hold more -- only the order of magnitude matters: a cube
    must be much larger than the 32 points of a warp and much smaller than a rank's share)."""
    lo = np.asarray(lo, np.float64)
    ext = float(np.max(np.asarray(hi, np.float64) - lo))
    c = int(np.clip(round((max(n_points_total, 1) / float(target_block)) ** (1.0 / 3.0)), 1, 256))
    if not ext > 0:
        return np.zeros(len(points), np.int64), c
    q = np.clip(((np.asarray(points, np.float64) - lo) * (c / ext)).astype(np.int64), 0, c - 1)
    return q[:, 0] + c * (q[:, 1] + c * q[:, 2]), c


def balanced_owner_table(cube_counts, world):
    """Owner rank of every cube from the GLOBAL number of points per cube: cubes in descending order of population, each to
    the rank with the fewest points so far (longest-processing-time rule; ties -> lowest rank, so every rank computes the same
    table).  Neighbouring cubes end up on different ranks, every rank holds the same number of points to within one cube."""
    cube_counts = np.asarray(cube_counts, np.int64)
    owner = np.zeros(len(cube_counts), np.int32)
    load = np.zeros(world, np.int64)
    order = np.argsort(-cube_counts, kind='stable')
    import heapq
    heap = [(0, r) for r in range(world)]
    for k in order:
        n = int(cube_counts[k])
        if n == 0:
            break
        l, r = heapq.heappop(heap)
        owner[k] = r
        load[r] = l + n
        heapq.heappush(heap, (l + n, r))
    return owner, load


def interleaved_shard(points, world, rank, target_block=512):
    """Boolean mask of the rows of `points` (the WHOLE cloud) that rank `rank` of `world` should hold."""
    ids, c = block_ids(points, points.min(0), points.max(0), len(points), target_block)
    owner, _ = balanced_owner_table(np.bincount(ids, minlength=c ** 3), world)
    return owner[ids] == rank


def exchange_to_owners(arrays, owner, dist, device):
    """All-to-all over torch.distributed (NCCL): every rank sends each row of `arrays` (same length, float32, 2-D) to the
    rank named in `owner` and returns the rows it receives.  One-off ingest step, not part of the solver."""
    import torch
    world = dist.get_world_size()
    order = np.argsort(owner, kind='stable')
    counts = np.bincount(owner, minlength=world).astype(np.int64)
    send_counts = torch.from_numpy(counts).to(device)
    recv_counts = torch.empty_like(send_counts)
    dist.all_to_all_single(recv_counts, send_counts)
    rc = recv_counts.cpu().numpy()
    out = []
    for a in arrays:
        a = np.ascontiguousarray(a[order], dtype=np.float32)
        w = a.shape[1]
        src = torch.from_numpy(a).to(device).reshape(-1)
        dst = torch.empty(int(rc.sum()) * w, dtype=torch.float32, device=device)
        dist.all_to_all_single(dst, src, output_split_sizes=[int(c) * w for c in rc], input_split_sizes=[int(c) * w for c in counts])
        out.append(dst.reshape(-1, w).cpu().numpy())
        del src, dst
    return out
