#!/usr/bin/env python
"""bench.py -- headline benchmark of the cniic_b200 hot path (BASELINE.json metric: Mpixel*iterations/s of Lloyd K-means).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c3|c2|c1|c4|c5|fill] [--impl ours|reference]

Default (no --workload): the north-star sharded configuration C3 -- voronoi k=2048 on ONE 7680x4320 image, rows sharded over the
N GPUs, STRONG scaling -- is the headline line at every N, and C2 (cluster-colors k=256, 4096x4096 per GPU) rides in the same
JSON line as `secondary`.
A "step" = one pass of the hot path over one batch of synthetic input = kmeans::cluster with max_iters = ITERS
(chunked init + ITERS fused assign/accumulate/update iterations) on the workload's image.
  value  : whole-job throughput, points resident in HBM when the timed region starts (CUDA events, max over ranks)
  e2e    : same metric through the host-buffer C-ABI call (cniic_kmeans_rgb / cniic_kmeans_xyrgb): pinned host image
           -> H2D -> kernels -> centroids/weights D2H, all inside the timed region
  roofline / cpu_baseline : see DESIGN.md "Measurement"
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ITERS = 10          # Lloyd iterations per step (fixed max_iters = 10 for throughput)
C4_BATCH = 64       # images per step per GPU of the batch workload (64 x 3 MiB = 192 MiB of input, larger than L2)
SEED = 0xC0FFEE

WORKLOADS = {
    # name: (kind, w, h, k, blobs, description, scaling)
    "c1": ("rgb", 512, 512, 16, 24, "cluster-colors k=16 on one 512x512 synthetic RGB image", "weak"),
    "c2": ("rgb", 4096, 4096, 256, 192, "cluster-colors k=256 on a 4096x4096 synthetic RGB image", "weak"),
    "c3": ("xyrgb", 7680, 4320, 2048, 2048, "voronoi k=2048 (x,y,r,g,b) on a 7680x4320 synthetic image", "strong"),
    "c4": ("rgb", 1024, 1024, 64, 16, "cluster-colors k=64, batches of 1024x1024 synthetic images, no collective "
                                         "(one batch per step per GPU through the batch API: one launch per stage for the whole batch)", "weak"),
}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "sm_max_mhz": d.get("sm_max_mhz", 1965.0), "source": "measured"}
    return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe).  ONE poller per job: rank 0 samples
    every GPU of the run (`gpus` = their indices); eight pollers at 100 ms each measurably disturb 30 us kernels."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpus):
        self.rows, self.proc = [], None
        self.gpus = [gpus] if isinstance(gpus, int) else list(gpus)

    def start(self):
        if not self.gpus:
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", ",".join(str(g) for g in self.gpus)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.gpus:
            return None
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for nm, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "gpus_sampled": len(self.gpus)}


def workload_image(name, w, h, blobs, seed_off=0, y0=0, rows=None):
    """Rows [y0, y0 + rows) of the workload's synthetic image, built by the numpy restatement of the generator (oracle/synth.py)
    so that the CPU legs never load the product library."""
    from oracle import synth
    return synth.synth_image(w, rows if rows else h, SEED + int(name[1]) + seed_off, blobs, y0=y0, h_total=h)


def cpu_sample_plan(name):
    """Bounded CPU sample of a workload: (band rows, k of the sample, description).  C3: a horizontal band of the SAME image with k
    scaled so that points-per-cluster stays N/k = 16200; RGB workloads: a band of the image with the full k (colour space has no
    spatial scale)."""
    kind, w, h, k, blobs, _, _ = WORKLOADS[name]
    if kind == "xyrgb":
        rows = max(8, h // 32)
        ks = max(1, round(k * rows / h))
        return rows, ks, f"band of {rows} rows of the {w}x{h} workload image, k={ks} (same points per cluster as k={k} on the whole image)"
    rows = min(h, max(8, (1 << 21) // w))
    return rows, min(k, w * rows), f"band of {rows} rows of the {w}x{h} workload image, k={min(k, w * rows)}"


def cpu_reference_leg(name, threads, seed_off=0, whole_image=False):
    """Times the CPU restatement of the reference's path (oracle VERBATIM mode = kmeans.rs incl. neighbour-list pruning; for the
    cluster-colors workloads through count_freqs + weighted points, clusterc.rs:19-28) on a bounded sample: `threads` bands, one
    per host thread, like bench.rs:27 (rayon, one image per worker).  Returns (Mpx*iter/s aggregate, description, seconds)."""
    import oracle as O
    kind, w, h, k, blobs, _, _ = WORKLOADS[name]
    if whole_image:
        rows, ks, what = h, k, f"the whole {w}x{h} workload image, k={k}"
    else:
        rows, ks, what = cpu_sample_plan(name)
    nb = max(1, h // rows)
    bands = [workload_image(name, w, h, blobs, seed_off, y0=((i * 7 + seed_off) % nb) * rows, rows=rows) for i in range(threads)]
    iters_done = [0] * threads
    uniq = [0] * threads

    def work(i):
        if kind == "rgb":
            keys, cnts = O.count_freqs_rgb(bands[i])                       # utils.rs:4-16 at clusterc.rs:21
            cols = np.stack([(keys >> 16) & 255, (keys >> 8) & 255, keys & 255], axis=1).astype(np.uint8)
            uniq[i] = len(keys)
            r = O.kmeans_rgb(cols, min(ks, len(keys)), counts=cnts.astype(np.uint32), mode=O.MODE_VERBATIM, max_iters=ITERS, allow_inactive=True)
        else:
            r = O.kmeans_xyrgb(bands[i], ks, mode=O.MODE_VERBATIM, max_iters=ITERS, allow_inactive=True)
        iters_done[i] = r.iterations

    O.lib()
    t0 = time.perf_counter()
    ths = [threading.Thread(target=work, args=(i,)) for i in range(threads)]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    dt = time.perf_counter() - t0
    px_iter = sum(w * rows * it for it in iters_done)
    desc = (f"{threads} x ({what}), {ITERS} Lloyd iterations each, one band per host thread (bench.rs:27 image-level parallelism); "
            f"oracle VERBATIM mode = C restatement of kmeans.rs incl. neighbour-list pruning"
            + (f" over the unique colours with counts (clusterc.rs:19-28; {int(np.mean(uniq))} unique colours per band on average, "
               f"count_freqs inside the timed region)" if kind == "rgb" else "") + ", not the Rust binary")
    return px_iter / dt / 1e6, desc, dt


def build_roofline(D, n_local, k, a_ms, pairs_per_launch, pk, kernel, traffic, brute_ms, culled=True):
    """Roofline object of the dominant kernel.  The default kernels cull exactly, so what bounds them is memory: the primary
    roofline is HBM (algorithmic bytes of SURVEY 8d: 3 B/pixel per Lloyd iteration).  The CUDA-core view the north star names is
    kept beside it under "fp32": `achieved` there counts the pairs the kernel actually scored, `algorithmic_equiv_tflops`
    divides the brute-force work (N*k*(2D+1) flops) by the same time, `brute_force_kernel` is a live measurement of the
    non-culled kernel in the same run."""
    flops_alg = (2 * D + 1) * n_local * k
    flops_exec = (2 * D + 1) * pairs_per_launch
    fp32_peak = 148 * 128 * 2 * pk["sm_max_mhz"] * 1e6 / 1e12
    exec_tf = flops_exec / (a_ms * 1e-3) / 1e12
    bytes_per_launch = 3 * n_local
    hbm_ach = bytes_per_launch / (a_ms * 1e-3) / 1e9
    brute_tf = flops_alg / (brute_ms * 1e-3) / 1e12 if brute_ms else None
    return {
        "bound": "hbm", "kernel": kernel, "achieved": hbm_ach, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": hbm_ach / pk["hbm_gbs"],
        "traffic": traffic, "launch_ms": a_ms, "algorithmic_bytes_per_launch": bytes_per_launch, "peak_source": pk["source"],
        "note": ("exact culling makes the Lloyd kernels memory/latency bound; the compute view is under `fp32`" if culled else
                 "small problem: the brute-force kernel runs (every pixel scores all k centroids), which is bound by integer-dot issue, "
                 "not HBM -- `fp32.frac_executed` is the fraction that describes it; the HBM figures are kept for the contract"),
        "fp32": {
            "peak": fp32_peak, "unit": "TFLOP/s",
            "peak_source": f"148 SM x 128 FP32 lanes x 2 x {pk['sm_max_mhz']:.0f} MHz ({pk['source']} sm_max_mhz)",
            "achieved_executed": exec_tf, "frac_executed": exec_tf / fp32_peak,
            "pairs_scored_frac_of_N_k": pairs_per_launch / (n_local * k) if n_local and k else None,
            "algorithmic_flops_per_launch": flops_alg, "algorithmic_equiv_tflops": flops_alg / (a_ms * 1e-3) / 1e12,
            "brute_force_kernel": None if brute_tf is None else {
                "kernel": "km_assign_rgb" if D == 3 else "km_assign_xyrgb", "launch_ms": brute_ms, "achieved": brute_tf,
                "frac": brute_tf / fp32_peak, "Mpx_iter_per_s": n_local / (brute_ms * 1e-3) / 1e6,
                "note": "every pixel scores all k centroids with IDP.4A/IDP.2A; frac > FFMA ceilings is possible (DESIGN.md)"},
        },
    }


def gpu_arm(args, name, K, W, rank, world, local_rank, ctxs, do_cpu):
    """Our arm for one K-means workload; returns the JSON line as a dict on every rank (rank 0 prints).  `ctxs` caches the
    library contexts (plain / distributed) so that the primary and the secondary workload share one NCCL communicator."""
    import torch
    import torch.distributed as dist
    import cniic_b200 as cb
    from cniic_b200 import dist as cdist

    kind, w, h, k, blobs, desc, scaling = WORKLOADS[name]
    D = 5 if kind == "xyrgb" else 3
    independent = world == 1 or name in ("c4", "c2")   # one image (batch) per GPU, no data-path collective: plain per-GPU context
    unique = name == "c2"   # cluster-colors at full size: K-means over the unique colours with counts, as the reference does (clusterc.rs:19-28)
    key = "plain" if independent else "dist"
    if key not in ctxs:
        ctxs[key] = cb.Context(local_rank) if independent else cdist.make_context(local_rank)
    ctx = ctxs[key]
    sctx = ctx
    stream = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", local_rank))

    # shard plan: weak scaling = every rank owns a full w x h slab of a w x (h*world) image; strong = rows of one image
    if independent:
        h_total, y0, h_local = h, 0, h
    elif scaling == "weak":
        h_total, y0, h_local = h * world, h * rank, h
    else:
        h_total = h
        y0, h_local = cdist.row_shard(h, world, rank)
    n_local, n_total = w * h_local, w * h_total
    B = max(1, args.batch) if name == "c4" else 1   # images per step per GPU
    d_img = ctx.device_alloc(n_local * 3 * B)
    for b in range(B):  # c4: B different images back to back; every rank gets its own
        cb.synth_image_device(ctx, d_img + b * n_local * 3, w, h_local, SEED + int(name[1]) + (rank * B + b if independent else 0),
                              blobs, y0=y0, h_total=h_total)
    ctx.sync()
    d_img_s = d_img
    kind_id = cb.POINTS_XYRGB if kind == "xyrgb" else cb.POINTS_RGB
    if independent:
        init = None
        skw = dict(w=w, h_local=h_local, on_device=True)
    else:
        first = y0 * w
        skw = dict(n_total=n_total, first_index=first, w=w, h_local=h_local, y0=y0, on_device=True)
        host_local = np.zeros((h_local, w, 3), np.uint8)
        ctx.d2h(host_local, d_img)
        init = cdist.gather_init_centroids(D, host_local, w, y0 if D == 5 else first, n_total, k, device=torch.device("cuda", local_rank))

    flush = ctxs.get("flush")
    if flush is None:
        flush = ctxs["flush"] = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def aligned_start():
        """barrier + synchronize, then every rank leaves at the same instant of the node's monotonic clock: ranks drop out of an NCCL
        barrier tens of microseconds apart, which a 0.5 ms step whose first exchange waits for the last rank would count as work."""
        barrier()
        if world > 1:
            t = torch.tensor([time.monotonic_ns() + 400_000], dtype=torch.int64, device="cuda")
            dist.broadcast(t, src=0)
            t_go = int(t.item())
            while time.monotonic_ns() < t_go:
                pass

    n_unique = [0]

    def one_step(flags=0, iters=ITERS):
        # a step = the whole kmeans::cluster call on HBM-resident points: session set-up (incl. the one-time colour ordering of
        # the culled RGB path), chunked init, ITERS Lloyd iterations
        if unique and not flags:
            # C2: count_freqs (Morton-binned histogram + ordered compaction = deduplicated, colour-sorted weighted points) and
            # kmeans::cluster over them, the front half of ClusterColors::encode without the recolour pass
            _, n_unique[0], st = sctx.cluster_colors_device(d_img_s, n_local, k, max_iters=iters)
            return st
        if B > 1:  # batch of independent images: same work per image, one launch per stage for the whole batch
            ss = [cb.KMeansSession(sctx, kind_id, k, d_img_s + b * n_local * 3, n_local, flags=flags, **skw) for b in range(B)]
            cb.kmeans_reset_batch(ss)
            sts = cb.kmeans_run_batch(ss, iters)
            for s_ in ss:
                s_.close()
            st = sts[0]
            st.iterations = max(x.iterations for x in sts)
            st.pairs_scored = sum(x.pairs_scored for x in sts)
            return st
        # one C call = kmeans::cluster on this rank's points (open + reset + run + close; nothing is read back inside `value`)
        return cb.kmeans_cluster(sctx, kind_id, k, d_img_s, n_local, max_iters=iters, init_centroids=init, flags=flags,
                                 want_centroids=False, **skw)[3]

    for _ in range(W):
        one_step()
    sampler = ClockSampler(list(range(world)) if rank == 0 else [])  # rank 0 polls all GPUs of the job (one process per GPU, one node)
    sampler.start()        # spawns nvidia-smi: do it BEFORE the last warm-up step and the barrier, so no rank enters the timed
    one_step()             # region late (a late rank shows up as a wait inside every other rank's per-iteration exchange)
    barrier()
    launches0 = sctx.launches
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    assign_ms, loop_ms, iters_run, pairs_per_launch = [], [], 0, 0.0
    sess_culled = unique or D == 5 or n_total * k >= (1 << 27)  # mirrors cniic_kmeans_open: small RGB problems run the brute-force kernel
    t_wall0 = time.perf_counter()
    for i in range(K):
        flush.fill_(i & 0xff)           # evict the image from L2 between timed steps (untimed)
        if world > 1:
            aligned_start()             # every rank enters every timed step together (a late rank is waited for inside the exchange)
        else:
            torch.cuda.synchronize()
        ev[i][0].record(stream)
        st = one_step()
        ev[i][1].record(stream)
        assign_ms.append(st.assign_ms_avg)
        loop_ms.append(st.device_ms)
        iters_run += st.iterations
        if os.environ.get("CNIIC_BENCH_DEBUG"):
            torch.cuda.synchronize()
            print(f"[rank {rank}] step {i}: {ev[i][0].elapsed_time(ev[i][1]):.3f} ms total, loop {st.device_ms:.3f} ms, assign avg {st.assign_ms_avg:.4f} ms", file=sys.stderr, flush=True)
        pairs_per_launch = st.pairs_scored / max(1, st.iterations)
    barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = sctx.launches - launches0
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    t = torch.tensor([dev_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    # keep the same load running (untimed) for ~0.5 s so the clock sampler sees it; the count is identical on every rank
    for _ in range(min(300, int(500.0 / max(total_ms / K, 0.05)) + 1)):
        one_step()
    clocks = sampler.stop()
    px_total = n_total if not independent else n_local * world * B
    value = px_total * ITERS * K / (total_ms * 1e-3) / 1e6

    # ---- e2e: host buffers through the one-shot C-ABI call (H2D + kernels + D2H inside the timed region) ----
    pinned = torch.empty((B * h_local, w, 3), dtype=torch.uint8).pin_memory()
    host_img = pinned.numpy()
    ctx.d2h(host_img, d_img)
    host_imgs = [host_img[b * h_local:(b + 1) * h_local] for b in range(B)]
    e2e = None
    if independent:
        def e2e_step(c=sctx):
            if B > 1:
                return c.kmeans_rgb_batch(host_imgs, k, max_iters=ITERS, want_assign=False)
            if unique:
                return c.cluster_colors(host_img, k, max_iters=ITERS, want_image=False)
            if kind == "rgb":
                return c.kmeans_rgb(host_img, k, max_iters=ITERS, want_assign=False)
            return c.kmeans_xyrgb(host_img, k, max_iters=ITERS, want_assign=False)
        api_name = ("cniic_kmeans_rgb_batch" if B > 1 else "cniic_cluster_colors (out_rgb = NULL)" if unique else
                    "cniic_kmeans_rgb" if kind == "rgb" else "cniic_kmeans_xyrgb")
        d2h_bytes = int((k * 3 * 4 + k * 8 + 64) * B)
    else:
        # sharded session: per step the shard is re-uploaded from pinned host memory and centroids/weights read back
        def e2e_step(c=sctx):
            return cb.kmeans_cluster(c, kind_id, k, host_img, n_local, max_iters=ITERS, init_centroids=init, n_total=n_total,
                                     first_index=y0 * w, w=w, h_local=h_local, y0=y0, on_device=False, want_centroids=True)
        api_name = "cniic_kmeans_cluster (row-sharded, host points in, centroids + weights out)"
        d2h_bytes = int(k * D * 4 + k * 8)
    for _ in range(2):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(K):
        e2e_step()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    e2e = {"value": px_total * ITERS * K / float(tt.item()) / 1e6, "unit": "Mpx*iter/s",
           "h2d_bytes_per_step": int(n_local * 3 * B), "d2h_bytes_per_step": d2h_bytes, "api": api_name, "callers": 1}
    if world == 1:
        # The same calls from TWO host threads, one context each -- how the reference's harness drives a codec (bench.rs:27:
        # par_iter over the images): one caller's upload overlaps the other's kernels.  Every call still carries its own
        # host->device and device->host copies; `value` above stays the single-caller figure.
        try:
            import threading
            ctx2 = cb.Context(local_rank)
            e2e_step(ctx2)
            torch.cuda.synchronize()
            share = [K - K // 2, K // 2]
            errs = []

            def caller(c, n_calls):
                try:
                    for _ in range(n_calls):
                        e2e_step(c)
                except BaseException as e:  # noqa: BLE001 -- re-raised below
                    errs.append(e)
            ths = [threading.Thread(target=caller, args=(c, n_calls)) for c, n_calls in zip((sctx, ctx2), share) if n_calls]
            t0 = time.perf_counter()
            for th in ths:
                th.start()
            for th in ths:
                th.join()
            torch.cuda.synchronize()
            dt2 = time.perf_counter() - t0
            ctx2.close()
            if errs:
                raise errs[0]
            e2e["two_callers"] = {"value": px_total * ITERS * K / dt2 / 1e6, "callers": len(ths), "calls": K,
                                  "note": "one context per host thread, as bench.rs:27 runs codec calls; copies of one call overlap kernels of the other"}
        except Exception as e:  # the single-caller figure stands on its own
            e2e["two_callers"] = {"error": repr(e)[:200]}

    # the brute-force kernel (every pixel scores all k centroids) for reference: same results, no culling.
    # Collective in the sharded case, so every rank runs it.
    one_step(cb._lib.KMEANS_NO_CULL, 1)
    stb = one_step(cb._lib.KMEANS_NO_CULL, 3)
    ctx.device_free(d_img)
    if rank != 0:
        return None

    # ---- roofline of the dominant kernel (fused assign+accumulate), measured live with CUDA events ----
    pk = peaks()
    a_ms = float(np.mean(assign_ms))
    # second kernel versions are the default; CNIIC_*_CULL_V1=1 selects the first ones for A/B runs
    kernel = (("km_assign_rgb_cull" if os.environ.get("CNIIC_RGB_CULL_V1") else "km_assign_rgb_cull2") if D == 3 else
              ("km_assign_xyrgb_cull" if os.environ.get("CNIIC_XY_CULL_V1") else "km_assign_xyrgb_cull2"))
    if not sess_culled:
        kernel = "km_assign_rgb" if D == 3 else "km_assign_xyrgb"
    if unique:
        kernel = "km_assign_rgb_cull2<true>"  # weighted points
    if B > 1:
        kernel += "_batch"
    traffic = None
    for tp in (os.path.join(ROOT, "profiles", "r02_traffic.json"), os.path.join(ROOT, "profiles", "r01_traffic.json")):
        if os.path.exists(tp) and world == 1 and traffic is None:
            ent = json.load(open(tp)).get(kernel)
            if ent and ent.get("workload") == name:
                traffic = ent["bytes"]  # dram bytes per launch from the committed ncu --set full capture of this kernel/workload
    roofline = build_roofline(D, n_local * B, k, a_ms, pairs_per_launch, pk, kernel, traffic, stb.assign_ms_avg, culled=sess_culled)  # one launch = B images

    cpu = None
    if do_cpu:
        # bounded sample, one host thread (the reference's K-means is single-threaded): bands until >= 10 s of CPU work are timed
        tot_px_iter, tot_dt, nb, descr = 0.0, 0.0, 0, ""
        while tot_dt < 10.0 and nb < 8:
            v, descr, dt_ = cpu_reference_leg(name, 1, seed_off=nb)
            tot_px_iter += v * dt_
            tot_dt += dt_
            nb += 1
        cpu = {"value": tot_px_iter / tot_dt, "unit": "Mpx*iter/s", "cores": 1, "kind": "port",
               "sample": f"{nb} x ({descr})", "seconds": tot_dt}

    if unique:
        u = n_unique[0]
        roofline["unique_colour_view"] = {
            "unique_colours": u, "fraction_of_pixels": u / max(1, n_local),
            "Mcolour_iter_per_s": u * ITERS * K / (total_ms * 1e-3) / 1e6,
            "algorithmic_bytes_per_launch": 10 * u, "achieved": 10 * u / (a_ms * 1e-3) / 1e9, "unit": "GB/s",
            "note": "the reference clusters unique colours with counts (clusterc.rs:19-28), so does this step: one launch reads 4 B colour "
                    "+ 4 B count + 2 B current cluster per unique colour; `achieved`/`frac` above keep the contract's 3 B per PIXEL the "
                    "launch accounts for (SURVEY 8d)"}
    exchange = ("partial sums pushed over NVLink peer memory inside the multi-CTA update kernel" if os.environ.get("CNIIC_P2P", "1") == "1"
                else "ncclAllReduce")
    step_ms = total_ms / K
    return {"metric": "Mpixel*iter/s Lloyd K-means", "value": value, "unit": "Mpx*iter/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": step_ms, "higher_is_better": True, "scaling": scaling if world > 1 or name != "c3" else "strong",
            "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": {"workload": desc + ((f", {world} slabs of {w}x{h} (row-sharded {w}x{h_total}" if scaling == "weak" else
                                             f" (rows sharded over {world} GPUs") + ", u64 partial sums exchanged per iteration: " + exchange + ")"
                                            if world > 1 and not independent else ""),
                       "name": name, "k": k, "dims": D, "iters_per_step": ITERS, "pixels": int(px_total), "images_per_step_per_gpu": B,
                       "parallelism": "1 GPU" if world == 1 else (f"{world} GPUs, independent batches of images" if independent else f"row-sharded x{world}"),
                       "l2": "512 MiB buffer written between timed steps (L2 flush); the image stays L2/HBM resident across the "
                             "iterations of one step, as the algorithm iterates over it"},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches // K), "clocks": clocks,
            "breakdown": {"assign_kernel_ms_avg": a_ms, "iteration_loop_ms_per_step": float(np.mean(loop_ms)),
                          "session_setup_and_readback_ms_per_step": step_ms - float(np.mean(loop_ms)),
                          "per_iteration_overhead_ms": (float(np.mean(loop_ms)) - ITERS * a_ms) / ITERS,
                          "note": "rank 0's own CUDA events; overhead = update kernel + exchange + launch gaps per Lloyd iteration"},
            "iterations_run": iters_run, "wall_s": t_wall}


def reference_arm(args, name, K, W):
    """--impl reference: the reference's own CPU path for the workload (oracle port; the Rust crate cannot be built here), with all
    the host threads it can use, on bounded samples.  Loads nothing of the product: images come from oracle/synth.py."""
    kind, w, h, k, blobs, desc, scaling = WORKLOADS[name]
    cores = os.cpu_count() or 1
    threads = max(1, min(cores, 32))
    vals, descr = [], ""
    for i in range(W + K):
        v, descr, dt = cpu_reference_leg(name, threads, seed_off=i)
        if i >= W:
            vals.append((v, dt))
    value = float(np.mean([v for v, _ in vals]))
    ms = float(np.mean([dt for _, dt in vals]) * 1e3)
    single, sdesc, sdt = cpu_reference_leg(name, 1, seed_off=0)  # what kmeans.rs does for ONE image: one thread
    line = {"impl": "reference", "metric": "Mpixel*iter/s Lloyd K-means", "value": value, "unit": "Mpx*iter/s",
            "n_gpus": args.gpus, "steps": K, "warmup": W, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "strong" if name == "c3" else scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": desc, "name": name, "k": k, "iters_per_step": ITERS, "sample": descr},
            "cpu_baseline": {"value": value, "unit": "Mpx*iter/s", "cores": threads, "kind": "port", "sample": descr},
            "single_image_single_thread": {"value": single, "unit": "Mpx*iter/s", "cores": 1, "sample": sdesc, "seconds": sdt,
                                           "note": "the reference clusters one image on one thread (kmeans.rs has no parallelism); the "
                                                   "headline value above is one band per host thread (bench.rs:27)"},
            "e2e": {"value": value, "unit": "Mpx*iter/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None, choices=sorted(WORKLOADS) + ["c5", "fill"],
                    help="default: c3 (north-star sharded configuration, strong scaling) with c2 as `secondary` in the same line")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-secondary", action="store_true", help="default run: skip the secondary C2 object")
    ap.add_argument("--batch", type=int, default=C4_BATCH, help="images per step per GPU (workload c4)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    if args.workload in ("c5", "fill"):  # HBM-bound stage workloads (sharded without any collective): bench_stages.py
        if args.impl == "reference":
            if rank == 0:
                print(json.dumps({"impl": "reference", "unavailable": "stage workloads report their CPU leg as cpu_baseline of the default arm"}))
            return 0
        import bench_stages
        return bench_stages.run(args, args.workload, peaks, ClockSampler)
    primary = args.workload or "c3"
    with_secondary = args.workload is None and not args.no_secondary
    W = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    K = max(args.steps, 1)

    if args.impl == "reference":
        if rank != 0:
            return 0
        line = reference_arm(args, primary, K, W)
        if with_secondary:
            sec = reference_arm(args, "c2", 1, 0)
            # the reference's real C2 path at the SAME config: one 4096x4096 image, unique colours with counts, one thread
            v, d, dt = cpu_reference_leg("c2", 1, whole_image=True)
            sec["same_config_single_image"] = {"value": v, "unit": "Mpx*iter/s", "cores": 1, "sample": d, "seconds": dt}
            line["secondary"] = {kk: sec[kk] for kk in ("metric", "value", "unit", "ms_per_step", "scaling", "config", "cpu_baseline",
                                                        "single_image_single_thread", "same_config_single_image")}
        print(json.dumps(line), flush=True)
        return 0

    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ctxs = {}
    line = gpu_arm(args, primary, K, W, rank, world, local_rank, ctxs, do_cpu=not args.no_cpu)
    if with_secondary:
        sec = gpu_arm(args, "c2", max(3, min(K, 5)), 3, rank, world, local_rank, ctxs, do_cpu=False)
        if rank == 0:
            line["secondary"] = {kk: sec[kk] for kk in ("metric", "value", "unit", "n_gpus", "steps", "ms_per_step", "scaling", "config",
                                                        "roofline", "e2e", "gpu_launches", "breakdown", "clocks")}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
