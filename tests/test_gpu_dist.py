"""Multi-GPU parity (needs >= 2 GPUs; the 4- and 8-rank variants skip below that many): one process per GPU; the row-sharded Lloyd
loop must reproduce the single-GPU result bit for bit with either exchange, and the stages that shard without a collective
(voronoi fill by rows, delta + histogram by curve range) must concatenate / merge to the single-GPU output."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# (kind is 5 = (x,y,r,g,b) or 3 = rgb, w, h, k, max_iters).  k = 600 spans three 256-cluster slices of the update kernel (one CTA and
# one flag each).  The small problem converges before max_iters, so the remaining launches of the batch exit early on every rank (the
# exchange parity must follow the exchanges really made, not the launches), and it runs twice in a row on purpose.
CASES = ((5, 640, 363, 128, 6), (3, 512, 301, 64, 6), (5, 384, 203, 600, 5), (5, 96, 64, 4, 40), (5, 96, 64, 4, 40))

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
CASES = eval(os.environ["CNIIC_CASES"])
for ci, (D, w, h, k, iters) in enumerate(CASES):
    kind = cb.POINTS_XYRGB if D == 5 else cb.POINTS_RGB
    full = cb.synth_image_host(w, h, 123 + ci, 16)
    y0, hl = cdist.row_shard(h, world, rank)
    local = np.ascontiguousarray(full[y0:y0 + hl])
    init = cdist.gather_init_centroids(D, local, w, y0 if D == 5 else y0 * w, w * h, k, device=torch.device("cuda", lr))
    s = cb.KMeansSession(ctx, kind, k, local, w * hl, n_total=w * h, first_index=y0 * w, w=w, h_local=hl, y0=y0)
    s.reset(init)
    st = s.run(iters)
    cen, wts, asg = s.get()
    s.close()
    out[f"cen{ci}"], out[f"wts{ci}"], out[f"asg{ci}"], out[f"it{ci}"] = cen, wts, asg, np.array([st.iterations, st.moved_last, st.converged])
# empty-cluster repair across shards (kmeans.rs:117-134 stand-in): palette point lists that empty clusters
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
# the stages that shard without a collective, one rank per GPU: voronoi fill by rows, delta + histogram by curve range
w, h, k = 640, 363, 128
rng = np.random.default_rng(5)
cxy = np.stack([rng.integers(0, w, k), rng.integers(0, h, k)], axis=1).astype(np.uint32)
crgb = rng.integers(0, 256, (k, 3)).astype(np.uint8)
y0, hl = cdist.row_shard(h, world, rank)
out["fill"] = ctx.voronoi_fill_rows(cxy, crgb, w, h, y0, hl)
for (sw, sh) in ((256, 256), (200, 117)):
    simg = cb.synth_image_host(sw, sh, 31, 8)
    d_img = ctx.device_alloc(sw * sh * 3)
    ctx.h2d(d_img, simg)
    i0, i1 = cdist.curve_shard(sw * sh, world, rank)
    d_out = ctx.device_alloc(max(16, (i1 - i0) * 6))
    ctx.delta_range_device(d_img, sw, sh, i0, i1, d_out)
    part = np.zeros((i1 - i0, 3), np.int16)
    if i1 > i0:
        ctx.d2h(part, d_out)
    keys, cnts = ctx.hist_delta_range_device(d_img, sw, sh, i0, i1)
    out[f"delta{sw}"], out[f"hk{sw}"], out[f"hc{sw}"] = part, keys, cnts
    ctx.device_free(d_img); ctx.device_free(d_out)
np.savez(os.path.join(os.environ["CNIIC_OUT"], f"rank{rank}.npz"), **out)
dist.barrier()
dist.destroy_process_group()
'''


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("p2p", ["0", "1"])
def test_row_sharded_kmeans_matches_single_gpu(tmp_path, p2p, world):
    """p2p=1 (default): partial sums pushed over peer memory inside the multi-CTA update kernel; p2p=0: ncclAllReduce between the
    assign and the update kernel."""
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    import cniic_b200 as cb
    from cniic_b200 import dist as cdist
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, CNIIC_ROOT=ROOT, CNIIC_OUT=str(tmp_path), CNIIC_P2P=p2p, CNIIC_CASES=repr(CASES))
    subprocess.check_call([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr",
                           "127.0.0.1", "--master-port", str(port), str(script)], env=env, timeout=240)
    R = [np.load(tmp_path / f"rank{r}.npz") for r in range(world)]
    ctx = cb.Context(0)
    for ci, (D, w, h, k, iters) in enumerate(CASES):
        full = cb.synth_image_host(w, h, 123 + ci, 16)
        g = ctx.kmeans_xyrgb(full, k, max_iters=iters) if D == 5 else ctx.kmeans_rgb(full, k, max_iters=iters)
        for r in range(world):
            assert np.array_equal(R[r][f"cen{ci}"], g.centroids), (ci, r)  # identical on every rank without a broadcast
            assert np.array_equal(R[r][f"wts{ci}"], g.weights), (ci, r)
            assert R[r][f"it{ci}"].tolist() == [g.iterations, g.moved_last, int(g.converged)], (ci, r)
        assert np.array_equal(np.concatenate([R[r][f"asg{ci}"] for r in range(world)]), g.assign), ci
        if ci >= 3:
            assert g.converged and g.iterations < iters  # the early-exit path was really taken
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
        for r in range(world):
            assert np.array_equal(R[r][f"e_cen{seed}"], g.centroids)
            assert R[r][f"e_it{seed}"].tolist() == [g.iterations, g.empty_events]
        assert np.array_equal(np.concatenate([R[r][f"e_asg{seed}"] for r in range(world)]), g.assign)
        events += g.empty_events
    assert events > 0
    # fill rows and curve ranges, one rank per GPU
    w, h, k = 640, 363, 128
    rng = np.random.default_rng(5)
    cxy = np.stack([rng.integers(0, w, k), rng.integers(0, h, k)], axis=1).astype(np.uint32)
    crgb = rng.integers(0, 256, (k, 3)).astype(np.uint8)
    assert np.array_equal(np.concatenate([R[r]["fill"] for r in range(world)]), ctx.voronoi_fill(cxy, crgb, w, h))
    for (sw, sh) in ((256, 256), (200, 117)):
        simg = cb.synth_image_host(sw, sh, 31, 8)
        assert np.array_equal(np.concatenate([R[r][f"delta{sw}"] for r in range(world)]), ctx.delta(simg))
        mk, mc = cdist.merge_histograms([(R[r][f"hk{sw}"], R[r][f"hc{sw}"]) for r in range(world)])
        gk, gc = ctx.hist_delta(simg)
        assert np.array_equal(mk, gk) and np.array_equal(mc, gc)
    ctx.close()
