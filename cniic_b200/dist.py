"""Multi-GPU plumbing: one process per GPU (torchrun), torch.distributed for rendezvous, NCCL inside the library.

The Lloyd loop shards ROWS of the image: rank r owns rows [r*h//W, (r+1)*h//W).  Per iteration the
only exchange is an all-reduce of the k x (D+1) u64 partial sums (+1 moved counter) issued by the library on its own
stream (ncclAllReduce, ncclUint64/ncclSum) -- integer sums, so the result is bit-identical for any world size.
Initial centroids (kmeans.rs:101-108) are k points at fixed global indices; each rank contributes the ones it owns
and a sum all-reduce (torch.distributed, gloo or nccl) assembles them on every rank.
"""
from __future__ import annotations

import numpy as np


def row_shard(h: int, world: int, rank: int) -> tuple[int, int]:
    """(first row, number of rows) owned by `rank`."""
    y0 = rank * h // world
    y1 = (rank + 1) * h // world
    return y0, y1 - y0


def curve_shard(n: int, world: int, rank: int, align: int = 4096) -> tuple[int, int]:
    """[i_begin, i_end) of the Hilbert-curve indices owned by `rank` (the integer stages shard the CURVE, not
    the rows).  Boundaries are multiples of `align` (the tile kernels work on 4096-index blocks), the last rank takes the rest."""
    blocks = (n + align - 1) // align
    b0 = rank * blocks // world
    b1 = (rank + 1) * blocks // world
    return min(n, b0 * align), min(n, b1 * align)


def merge_histograms(parts) -> tuple[np.ndarray, np.ndarray]:
    """Sum of per-rank partial histograms [(keys ascending, counts), ...] -> (keys ascending, counts) (utils.rs:4-16 over the
    whole stream: count_freqs is additive over any partition of its input)."""
    keys = np.concatenate([np.asarray(k, dtype=np.uint32) for k, _ in parts]) if parts else np.zeros(0, np.uint32)
    cnts = np.concatenate([np.asarray(c, dtype=np.uint64) for _, c in parts]) if parts else np.zeros(0, np.uint64)
    if len(keys) == 0:
        return keys, cnts
    uk, inv = np.unique(keys, return_inverse=True)
    out = np.zeros(len(uk), np.uint64)
    np.add.at(out, inv, cnts)
    return uk.astype(np.uint32), out


def init_point_indices(n_total: int, k: int) -> np.ndarray:
    """Global point index of each initial centroid: N-(i+1)*ppc for i < k-1, and 0 for i = k-1 (kmeans.rs:61-108)."""
    ppc = n_total // k
    if ppc == 0:
        raise ValueError("fewer points than clusters (kmeans.rs:67-68)")
    idx = n_total - (np.arange(k, dtype=np.int64) + 1) * ppc
    idx[k - 1] = 0
    return idx


def local_init_contribution(kind_dims: int, local_rgb: np.ndarray, w: int, y0: int, n_total: int, k: int) -> np.ndarray:
    """(k, D) int64 array holding the initial centroids this rank owns (zeros elsewhere)."""
    flat = np.ascontiguousarray(local_rgb, dtype=np.uint8).reshape(-1, 3)
    first = y0 * w if kind_dims == 5 else y0
    idx = init_point_indices(n_total, k)
    mine = (idx >= first) & (idx < first + len(flat))
    out = np.zeros((k, kind_dims), np.int64)
    li = idx[mine] - first
    if kind_dims == 5:
        out[mine, 0] = idx[mine] % w
        out[mine, 1] = idx[mine] // w
        out[mine, 2:] = flat[li]
    else:
        out[mine] = flat[li]
    return out


def gather_init_centroids(kind_dims, local_rgb, w, y0, n_total, k, device=None) -> np.ndarray:
    """All ranks call this; returns the same (k, D) int32 initial centroids everywhere."""
    import torch
    import torch.distributed as dist
    contrib = local_init_contribution(kind_dims, local_rgb, w, y0, n_total, k)
    t = torch.from_numpy(contrib)
    if device is not None:
        t = t.to(device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy().astype(np.int32)


def make_context(local_rank: int | None = None):
    """Create the library context of this rank; broadcasts the ncclUniqueId through torch.distributed."""
    import os
    import torch
    import torch.distributed as dist
    from .api import Context
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    dev = int(os.environ.get("LOCAL_RANK", 0)) if local_rank is None else local_rank
    if world == 1:
        return Context(dev)
    uid = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        uid = torch.frombuffer(bytearray(Context.nccl_unique_id()), dtype=torch.uint8).clone()
    if dist.get_backend() == "nccl":
        uid = uid.cuda(dev)
    dist.broadcast(uid, src=0)
    ctx = Context(dev, rank=rank, world=world, nccl_unique_id=bytes(uid.cpu().numpy().tobytes()))
    if os.environ.get("CNIIC_P2P", "1") == "1":
        # peer-memory all-reduce fused into the finalize kernel (default; CNIIC_P2P=0 selects ncclAllReduce): exchange the
        # CUDA IPC handles of the per-rank exchange regions
        mine = torch.frombuffer(bytearray(ctx.p2p_export()), dtype=torch.uint8).clone()
        if dist.get_backend() == "nccl":
            mine = mine.cuda(dev)
        allh = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allh, mine)
        ctx.p2p_connect(b"".join(bytes(h.cpu().numpy().tobytes()) for h in allh))
        dist.barrier()
    return ctx
