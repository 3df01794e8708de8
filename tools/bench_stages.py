"""Secondary measurements (GPU box): voronoi fill and the integer pre-Huffman stages, device-resident, CUDA events.
Prints one JSON line per stage; results are copied into profiles/ by hand.  Not the headline bench (bench.py)."""
import ctypes as C, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import cniic_b200 as cb

HBM = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
    os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else 6650.0
ctx = cb.Context(0)
lib = ctx._lib
stream = torch.cuda.ExternalStream(ctx.stream)
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")


def timeit(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    ctx.sync()
    ts = []
    for i in range(reps):
        flush.fill_(i)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        fn()
        b.record(stream)
        ctx.sync()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


def line(name, px, ms, bytes_per_px, extra=None):
    gbs = px * bytes_per_px / (ms * 1e-3) / 1e9
    d = {"stage": name, "pixels": px, "ms": ms, "Mpix_per_s": px / (ms * 1e-3) / 1e6, "algorithmic_bytes_per_px": bytes_per_px,
         "achieved_GBps": gbs, "hbm_frac_of_measured": gbs / HBM}
    d.update(extra or {})
    print(json.dumps(d), flush=True)


# ---- voronoi fill at C3 size (7680x4320, k = 2048): centroids from 3 Lloyd iterations ----
w, h, k = 7680, 4320, 2048
d_img = ctx.device_alloc(w * h * 3)
cb.synth_image_device(ctx, d_img, w, h, 0xC0FFEE + 3, 2048)
s = cb.KMeansSession(ctx, cb.POINTS_XYRGB, k, d_img, w * h, w=w, h_local=h, on_device=True)
s.reset(); s.run(3)
cen, _, _ = s.get(want_assign=False)
s.close()
cxy = np.ascontiguousarray(cen[:, :2].astype(np.uint32)); crgb = np.ascontiguousarray(cen[:, 2:].astype(np.uint8))
d_cxy, d_crgb, d_out = ctx.device_alloc(cxy.nbytes), ctx.device_alloc(crgb.nbytes), ctx.device_alloc(w * h * 3)
ctx.h2d(d_cxy, cxy); ctx.h2d(d_crgb, crgb)
f = lambda: ctx.check(lib.cniic_voronoi_fill_device(ctx.h, C.c_void_p(d_cxy), C.c_void_p(d_crgb), C.c_uint32(k), C.c_uint32(w), C.c_uint32(h),
                                                     C.c_uint32(0), C.c_uint32(h), C.c_void_p(d_out)))
line("voronoi_fill 7680x4320 k=2048 (clusterc.rs:179-186)", w * h, timeit(f), 3, {"brute_force_pairs": w * h * k})
for p in (d_img, d_cxy, d_crgb, d_out):
    ctx.device_free(p)

# ---- integer stages at C5 size (8192x8192) ----
w = h = 8192
d_img = ctx.device_alloc(w * h * 3)
cb.synth_image_device(ctx, d_img, w, h, 0xC0FFEE + 5, 4096)
d_delta = ctx.device_alloc(w * h * 6)
f = lambda: ctx.check(lib.cniic_delta_i16_device(ctx.h, C.c_void_p(d_img), C.c_uint32(w), C.c_uint32(h), C.c_void_p(d_delta)))
line("hilbert gather + delta i16 8192x8192 (hilbert.rs:34-43, hilbertc.rs:449-477)", w * h, timeit(f), 9)
nuniq = C.c_size_t(0)
f = lambda: ctx.check(lib.cniic_hist_delta_device(ctx.h, C.c_void_p(d_img), C.c_uint32(w), C.c_uint32(h), C.byref(nuniq)))
ms = timeit(f, reps=3, warm=1)
line("fused hilbert + delta + joint-symbol histogram + compaction 8192x8192 (huf.rs:30)", w * h, ms, 3, {"distinct_symbols": nuniq.value})
