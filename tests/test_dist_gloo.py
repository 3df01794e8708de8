"""world_size-2 gloo tests (CPU): the host-side logic of the row-sharded Lloyd loop -- shard plan, initial-centroid
assembly through an all-reduce, and the per-iteration exchange (sum all-reduce of k x (D+1) partial sums + moved).
The per-shard arithmetic is numpy here; on the GPU the same partial sums come from the fused kernel."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle as O
import cniic_b200 as cb
from cniic_b200 import dist as cdist


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _points(img, w, y0):
    h = img.shape[0]
    flat = img.reshape(-1, 3).astype(np.int64)
    idx = np.arange(h * w, dtype=np.int64)
    return np.concatenate([(idx % w)[:, None], (y0 + idx // w)[:, None], flat], axis=1)


def _worker(rank, world, port, w, h, k, iters, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    full = cb.synth_image_host(w, h, 77, 6)
    y0, hl = cdist.row_shard(h, world, rank)
    local = full[y0:y0 + hl]
    n_total = w * h
    cen = cdist.gather_init_centroids(5, local, w, y0, n_total, k).astype(np.int64)
    pts = _points(local, w, y0)
    gidx = y0 * w + np.arange(len(pts), dtype=np.int64)
    ppc = n_total // k
    cur = np.where(gidx >= n_total - (k - 1) * ppc, (n_total - 1 - gidx) // ppc, k - 1)
    for _ in range(iters):
        d2 = ((pts[:, None, :] - cen[None]) ** 2).sum(-1)
        best = d2.argmin(1)
        keep = d2[np.arange(len(pts)), cur] == d2.min(1)
        new = np.where(keep, cur, best)
        sums = np.zeros(k * 6 + 1, np.int64)
        for j in range(5):
            sums[j:k * 6:6] = np.bincount(new, weights=pts[:, j], minlength=k).astype(np.int64)
        sums[5:k * 6:6] = np.bincount(new, minlength=k)
        sums[k * 6] = int((new != cur).sum())
        t = torch.from_numpy(sums)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)  # the one exchange per iteration
        s = t.numpy().reshape(-1)[:k * 6].reshape(k, 6)
        cnt = s[:, 5]
        assert (cnt > 0).all()
        cen = s[:, :5] // cnt[:, None]
        cur = new
    np.save(os.path.join(out_dir, f"cen{rank}.npy"), cen)
    np.save(os.path.join(out_dir, f"asg{rank}.npy"), cur)
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_row_sharded_lloyd_matches_single_process(tmp_path):
    w, h, k, iters, world = 48, 37, 12, 4, 2
    mp.spawn(_worker, args=(world, _free_port(), w, h, k, iters, str(tmp_path)), nprocs=world, join=True)
    o = O.kmeans_xyrgb(cb.synth_image_host(w, h, 77, 6), k, mode=O.MODE_EXACT, tie=O.TIE_KEEP_CURRENT, max_iters=iters)
    c0, c1 = np.load(tmp_path / "cen0.npy"), np.load(tmp_path / "cen1.npy")
    assert np.array_equal(c0, c1)  # every rank ends with identical centroids, no broadcast needed
    assert np.array_equal(c0, o.centroids)
    asg = np.concatenate([np.load(tmp_path / "asg0.npy"), np.load(tmp_path / "asg1.npy")])
    assert np.array_equal(asg, o.assign)
