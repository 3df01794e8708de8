"""Stage workloads of bench.py (--workload c5 | fill): the HBM-bound integer stages and the voronoi decode fill.
Same JSON contract as the K-means workloads; metric = Mpix/s.  Multi-GPU (one process per GPU, no data-path collective,
SURVEY.md 8e): the integer stages shard the Hilbert-CURVE index range (every rank holds the image, partial histograms are
merged on the host), the fill shards rows; strong scaling, time = max over ranks."""
from __future__ import annotations

import ctypes as C
import json
import os
import time

import numpy as np


SIZES = {"c5": (8192, 8192, 4096), "fill": (7680, 4320, 2048)}  # (w, h, blobs | k); tests/emu shrinks them


def run(args, workload, peaks, ClockSampler):
    import torch
    import torch.distributed as dist
    import cniic_b200 as cb
    from cniic_b200 import dist as cdist
    import oracle as O

    W, K = max(args.warmup, 3), max(args.steps, 1)
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ctx = cb.Context(local_rank)
    lib = ctx._lib
    stream = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", local_rank))
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
    pk = peaks()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    if workload == "c5":
        w, h, blobs = SIZES["c5"]
        desc = f"hilbert-rle / delta pre-Huffman stages (Hilbert index map + delta + histograms) on a {w}x{h} synthetic image"
        d_img = ctx.device_alloc(w * h * 3)
        cb.synth_image_device(ctx, d_img, w, h, 0xC0FFEE + 5, blobs)   # every rank holds the whole image
        i0, i1 = cdist.curve_shard(w * h, world, rank)                  # ... and owns a range of the curve
        d_delta = ctx.device_alloc(max(16, (i1 - i0) * 6))
        nuniq = C.c_size_t(0)

        def dev_step():  # one pass of the path: delta stream (huf.rs:38 pass 2 input) + fused symbol histogram (pass 1)
            if world == 1:
                ctx.check(lib.cniic_delta_i16_device(ctx.h, C.c_void_p(d_img), C.c_uint32(w), C.c_uint32(h), C.c_void_p(d_delta)))
                ctx.check(lib.cniic_hist_delta_device(ctx.h, C.c_void_p(d_img), C.c_uint32(w), C.c_uint32(h), C.byref(nuniq)))
            else:  # this rank's share of the curve; the (key, count) lists come back to the host for the merge
                ctx.delta_range_device(d_img, w, h, i0, i1, d_delta)
                ctx.hist_delta_range_device(d_img, w, h, i0, i1)
        launches_per_step = None
        alg_bytes = (3 + 6 + 3) * w * h  # delta: 3 B/px read + 6 B/px written; fused histogram: 3 B/px read
        kernel = "hilbert_tile_tma_kernel<1> (delta)" if os.environ.get("CNIIC_TILE_V1") else "hilbert_tile_tma2_kernel<1> (delta)"
        kernel_bytes = 9 * (i1 - i0)
        pinned = torch.empty((h, w, 3), dtype=torch.uint8).pin_memory()
        host = pinned.numpy()
        ctx.d2h(host, d_img)
        out_host = torch.empty((w * h, 3), dtype=torch.int16).pin_memory().numpy()

        def e2e_step():
            if world == 1:
                ctx.check(lib.cniic_delta_i16(ctx.h, host.ctypes.data_as(C.c_void_p), C.c_uint32(w), C.c_uint32(h), out_host.ctypes.data_as(C.c_void_p)))
            else:  # whole image up (every rank needs it), this rank's share of the stream down
                ctx.h2d(d_img, host)
                ctx.delta_range_device(d_img, w, h, i0, i1, d_delta)
                ctx.d2h(out_host[:i1 - i0], d_delta)
        h2d, d2h = w * h * 3, (i1 - i0) * 6
        cs = min(2048, w)
        crop = cb.synth_image_host(cs, cs, 0xC0FFEE + 5, 256)

        def cpu_step():
            d = O.delta(crop)
            O.hist_delta(d)
            return crop.shape[0] * crop.shape[1]
        cpu_desc = f"oracle delta + hist_delta (hilbertc.rs:449-477, utils.rs:4-16) on a {cs}x{cs} crop, 1 thread"

        def kernel_only():
            ctx.delta_range_device(d_img, w, h, i0, i1, d_delta)

        def second_kernel():  # the fused histogram pass alone (its compaction included)
            if world == 1:
                ctx.check(lib.cniic_hist_delta_device(ctx.h, C.c_void_p(d_img), C.c_uint32(w), C.c_uint32(h), C.byref(nuniq)))
            else:
                ctx.hist_delta_range_device(d_img, w, h, i0, i1)
    else:
        w, h, k = SIZES["fill"]
        desc = f"voronoi decode fill (clusterc.rs:179-186) k={k} on a {w}x{h} image"
        d_img = ctx.device_alloc(w * h * 3)
        cb.synth_image_device(ctx, d_img, w, h, 0xC0FFEE + 3, k)
        y0, hl = cdist.row_shard(h, world, rank)  # the fill shards rows; the centroid table (19 B x k) goes to every rank
        s = cb.KMeansSession(ctx, cb.POINTS_XYRGB, k, d_img, w * h, w=w, h_local=h, on_device=True)
        s.reset()
        s.run(3)
        cen, _, _ = s.get(want_assign=False)
        s.close()
        cxy = np.ascontiguousarray(cen[:, :2].astype(np.uint32))
        crgb = np.ascontiguousarray(cen[:, 2:].astype(np.uint8))
        d_cxy, d_crgb, d_out = ctx.device_alloc(cxy.nbytes), ctx.device_alloc(crgb.nbytes), ctx.device_alloc(max(16, w * hl * 3))
        ctx.h2d(d_cxy, cxy)
        ctx.h2d(d_crgb, crgb)

        def dev_step():
            ctx.check(lib.cniic_voronoi_fill_device(ctx.h, C.c_void_p(d_cxy), C.c_void_p(d_crgb), C.c_uint32(k), C.c_uint32(w), C.c_uint32(h),
                                                    C.c_uint32(y0), C.c_uint32(hl), C.c_void_p(d_out)))
        kernel_only = dev_step
        second_kernel = None
        alg_bytes = 3 * w * h + 19 * k
        kernel, kernel_bytes = "fill_kernel", 3 * w * hl
        out_host = torch.empty((max(1, hl), w, 3), dtype=torch.uint8).pin_memory().numpy()

        def e2e_step():
            if world == 1:
                ctx.check(lib.cniic_voronoi_fill(ctx.h, cxy.ctypes.data_as(C.c_void_p), crgb.ctypes.data_as(C.c_void_p), C.c_uint32(k), C.c_uint32(w),
                                                 C.c_uint32(h), out_host.ctypes.data_as(C.c_void_p)))
            else:  # centroid table up, this rank's rows down
                ctx.h2d(d_cxy, cxy)
                ctx.h2d(d_crgb, crgb)
                dev_step()
                ctx.d2h(out_host, d_out)
        h2d, d2h = 19 * k, w * hl * 3

        def cpu_step():
            O.voronoi_fill(cxy, crgb, w, min(24, h))  # 24 rows of the full-width image against all k centroids
            return w * min(24, h)
        cpu_desc = "oracle voronoi_fill (clusterc.rs:179-186, brute force over k) on 24 full-width rows, 1 thread"

    for _ in range(W):
        dev_step()
    ctx.sync()
    sampler = ClockSampler(list(range(world)) if rank == 0 else [])  # one poller per job (rank 0, all GPUs)
    sampler.start()
    dev_step()
    barrier()  # all ranks enter the timed region together
    l0 = ctx.launches
    tot = 0.0
    for i in range(K):
        flush.fill_(i & 0xff)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        dev_step()
        b.record(stream)
        ctx.sync()
        tot += a.elapsed_time(b)
    launches = (ctx.launches - l0) // K
    barrier()
    tot = max_over_ranks(tot)
    t_s = time.perf_counter()
    while time.perf_counter() - t_s < 0.5:  # keep the same load running (untimed) until the clock sampler has its samples
        dev_step()
        ctx.sync()
    clocks = sampler.stop()
    value = w * h * K / (tot * 1e-3) / 1e6
    # dominant kernel alone (CUDA events on the launching stream)
    kt = 0.0
    for i in range(5):
        flush.fill_(i)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        kernel_only()
        b.record(stream)
        ctx.sync()
        kt += a.elapsed_time(b) / 5
    ach = kernel_bytes / (kt * 1e-3) / 1e9
    second_ms = None
    if second_kernel is not None:
        second_ms = 0.0
        for i in range(5):
            flush.fill_(i)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            second_kernel()
            b.record(stream)
            ctx.sync()
            second_ms += a.elapsed_time(b) / 5
    traffic = None  # dram bytes per launch of the committed ncu --set full capture of this kernel at this size (single GPU)
    tp = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r02_traffic.json")
    if world == 1 and os.path.exists(tp) and not os.environ.get("CNIIC_STAGES_NO_TMA"):
        ent = json.load(open(tp)).get(kernel)
        if ent and ent.get("workload") == workload:
            traffic = ent["bytes"]
    roofline = {"bound": "hbm", "kernel": kernel, "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": ach / pk["hbm_gbs"],
                "traffic": traffic, "launch_ms": kt, "algorithmic_bytes_per_launch": kernel_bytes, "peak_source": pk["source"],
                "step_algorithmic_bytes": alg_bytes, "step_hbm_frac": alg_bytes * K / (tot * 1e-3) / 1e9 / pk["hbm_gbs"]}
    if second_ms is not None:  # c5: the fused histogram call (tile kernel + page compaction) beside the delta kernel
        roofline["histogram_call_ms"] = second_ms
        roofline["histogram_call_hbm_frac"] = 3 * (i1 - i0) / (second_ms * 1e-3) / 1e9 / pk["hbm_gbs"]
        # calibration: the delta kernel writes two bytes for each byte it reads, and `peak` is a COPY bandwidth (one read per write).
        # A plain streaming pass with the kernel's mix (torch u8 -> i16 cast of as many elements, contiguous both sides) shows what
        # the memory system delivers for it; reported beside the roofline, never as its denominator.
        src = torch.empty(3 * (i1 - i0), dtype=torch.uint8, device="cuda")
        dst = torch.empty(3 * (i1 - i0), dtype=torch.int16, device="cuda")
        dst.copy_(src)
        mix_ms = 0.0
        for i in range(5):
            flush.fill_(i)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            dst.copy_(src)
            b.record()
            torch.cuda.synchronize()
            mix_ms += a.elapsed_time(b) / 5
        roofline["same_mix_streaming_pass"] = {"what": "torch u8->i16 cast, 1 B read : 2 B written, contiguous", "ms": mix_ms,
                                               "gbs": kernel_bytes / (mix_ms * 1e-3) / 1e9, "kernel_vs_this": mix_ms / kt}
        del src, dst
    for _ in range(2):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(K):
        e2e_step()
    torch.cuda.synchronize()
    dt = max_over_ranks(time.perf_counter() - t0)
    e2e = {"value": w * h * K / dt / 1e6, "unit": "Mpix/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)}
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0
    cpu = None
    if not args.no_cpu:
        t0, px = time.perf_counter(), 0
        while time.perf_counter() - t0 < 10.0:
            px += cpu_step()
        dtc = time.perf_counter() - t0
        cpu = {"value": px / dtc / 1e6, "unit": "Mpix/s", "cores": 1, "kind": "port", "sample": cpu_desc, "seconds": dtc}
    shard = "" if world == 1 else (f", curve index range sharded over {world} GPUs (no collective; histograms merged on the host)"
                                   if workload == "c5" else f", rows sharded over {world} GPUs (no collective)")
    line = {"metric": "Mpix/s", "value": value, "unit": "Mpix/s", "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": tot / K,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u8/i16/u32", "data": "synthetic",
            "config": {"workload": desc + shard, "pixels": w * h, "l2": "512 MiB buffer written between timed steps (L2 flush)"},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0
