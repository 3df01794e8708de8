"""Whole-codec wall times at BASELINE sizes through the C ABI (host buffers in, bytes out).  GPU box helper."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import cniic_b200 as cb
from cniic_b200 import codecs

ctx = cb.Context(0)
cases = [("cluster-colors(256)", 4096, 4096, 192, 0), ("voronoi(2048)", 7680, 4320, 2048, 0), ("delta", 8192, 8192, 4096, 0),
         ("hufman", 4096, 4096, 192, 0), ("hilbert(rle)", 4096, 4096, 192, 0)]
for expr, w, h, blobs, max_iters in cases:
    d = ctx.device_alloc(w * h * 3)
    cb.synth_image_device(ctx, d, w, h, 0xC0FFEE + 7, blobs)
    img = np.zeros((h, w, 3), np.uint8)
    ctx.d2h(img, d)
    ctx.device_free(d)
    if expr.startswith("hilbert"):
        img = (img // 32) * 32
    c = codecs.Codec.from_str(ctx, expr, max_iters)
    c.encode(img[:64, :64].copy())  # warm the context
    t0 = time.perf_counter(); data = c.encode(img); t1 = time.perf_counter()
    dec = c.decode(data); t2 = time.perf_counter()
    sse = ctx.sse(img, dec)
    ok = (sse == 0) if c.is_lossless() else True
    print(json.dumps({"codec": c.name(), "image": f"{w}x{h}", "encode_s": round(t1 - t0, 4), "decode_s": round(t2 - t1, 4),
                      "bytes": len(data), "ratio_vs_raw": round(len(data) / (w * h * 3), 5), "mse": sse / (w * h),
                      "lossless_roundtrip_ok": ok, "encode_Mpix_s": round(w * h / (t1 - t0) / 1e6, 1)}), flush=True)
