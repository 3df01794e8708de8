"""Multi-GPU parity (needs >= 2 GPUs; skipped otherwise): the row-sharded Lloyd loop with the in-library NCCL all-reduce
must reproduce the single-GPU result bit for bit."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
sys.path.insert(0, os.environ["CNIIC_ROOT"])
import numpy as np, torch, torch.distributed as dist
import cniic_b200 as cb
from cniic_b200 import dist as cdist
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
ctx = cdist.make_context(lr)
out = {}
for kind, w, h, k in ((cb.POINTS_XYRGB, 640, 363, 128), (cb.POINTS_RGB, 512, 301, 64)):
    full = cb.synth_image_host(w, h, 123, 16)
    D = 5 if kind == cb.POINTS_XYRGB else 3
    y0, hl = cdist.row_shard(h, world, rank)
    local = np.ascontiguousarray(full[y0:y0 + hl])
    init = cdist.gather_init_centroids(D, local, w, y0 if D == 5 else y0 * w, w * h, k, device=torch.device("cuda", lr))
    s = cb.KMeansSession(ctx, kind, k, local, w * hl, n_total=w * h, first_index=y0 * w, w=w, h_local=hl, y0=y0)
    s.reset(init)
    st = s.run(6)
    cen, wts, asg = s.get()
    s.close()
    out[f"cen{D}"], out[f"wts{D}"], out[f"asg{D}"], out[f"it{D}"] = cen, wts, asg, np.array([st.iterations, st.moved_last])
# empty-cluster repair across shards (kmeans.rs:117-134 stand-in): palette point lists that empty clusters
if True:
    for seed in (0, 4, 9):
        rng = np.random.default_rng(seed)
        npal, k, n = int(rng.integers(3, 10)), int(rng.integers(4, 16)), int(rng.integers(40, 400))
        palette = rng.integers(0, 256, size=(npal, 3), dtype=np.uint8)
        pts = palette[rng.integers(0, npal, size=n)]
        jit = rng.integers(-3, 4, size=pts.shape); mask = rng.random(n) < 0.3
        pts = np.clip(pts.astype(int) + jit * mask[:, None], 0, 255).astype(np.uint8)
        lo, cnt = cdist.row_shard(n, world, rank)
        local = np.ascontiguousarray(pts[lo:lo + cnt])
        init = cdist.gather_init_centroids(3, local, 1, lo, n, k, device=torch.device("cuda", lr))
        s = cb.KMeansSession(ctx, cb.POINTS_RGB, k, local, cnt, n_total=n, first_index=lo, tie=cb.TIE_LOWEST_INDEX)
        s.reset(init)
        st = s.run(8)
        cen, wts, asg = s.get()
        s.close()
        out[f"e_cen{seed}"], out[f"e_asg{seed}"], out[f"e_it{seed}"] = cen, asg, np.array([st.iterations, st.empty_events])
np.savez(os.path.join(os.environ["CNIIC_OUT"], f"rank{rank}.npz"), **out)
dist.destroy_process_group()
'''


@pytest.mark.parametrize("p2p", ["0", "1"])
def test_two_gpu_row_sharded_kmeans_matches_single_gpu(tmp_path, p2p):
    """p2p=0: ncclAllReduce between assign and finalize (default); 1: peer-memory all-reduce fused into km_finalize."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import cniic_b200 as cb
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, CNIIC_ROOT=ROOT, CNIIC_OUT=str(tmp_path), CNIIC_P2P=p2p)
    subprocess.check_call([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                           "127.0.0.1", "--master-port", str(port), str(script)], env=env, timeout=120)
    r0, r1 = np.load(tmp_path / "rank0.npz"), np.load(tmp_path / "rank1.npz")
    ctx = cb.Context(0)
    for kind, w, h, k, D in ((cb.POINTS_XYRGB, 640, 363, 128, 5), (cb.POINTS_RGB, 512, 301, 64, 3)):
        full = cb.synth_image_host(w, h, 123, 16)
        g = ctx.kmeans_xyrgb(full, k, max_iters=6) if D == 5 else ctx.kmeans_rgb(full, k, max_iters=6)
        assert np.array_equal(r0[f"cen{D}"], r1[f"cen{D}"])  # identical on every rank without a broadcast
        assert np.array_equal(r0[f"cen{D}"], g.centroids)
        assert np.array_equal(r0[f"wts{D}"], g.weights)
        assert np.array_equal(np.concatenate([r0[f"asg{D}"], r1[f"asg{D}"]]), g.assign)
        assert r0[f"it{D}"].tolist() == [g.iterations, g.moved_last]
    if True:
        events = 0
        for seed in (0, 4, 9):
            rng = np.random.default_rng(seed)
            npal, k, n = int(rng.integers(3, 10)), int(rng.integers(4, 16)), int(rng.integers(40, 400))
            palette = rng.integers(0, 256, size=(npal, 3), dtype=np.uint8)
            pts = palette[rng.integers(0, npal, size=n)]
            jit = rng.integers(-3, 4, size=pts.shape)
            mask = rng.random(n) < 0.3
            pts = np.clip(pts.astype(int) + jit * mask[:, None], 0, 255).astype(np.uint8)
            g = ctx.kmeans_rgb(pts, k, max_iters=8, tie=cb.TIE_LOWEST_INDEX, allow_inactive=True)
            assert np.array_equal(r0[f"e_cen{seed}"], g.centroids) and np.array_equal(r1[f"e_cen{seed}"], g.centroids)
            assert np.array_equal(np.concatenate([r0[f"e_asg{seed}"], r1[f"e_asg{seed}"]]), g.assign)
            assert r0[f"e_it{seed}"].tolist() == [g.iterations, g.empty_events]
            events += g.empty_events
        assert events > 0
