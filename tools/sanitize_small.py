"""compute-sanitizer memcheck target: one small invocation of every kernel family (GPU box helper, not a test)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import cniic_b200 as cb
from cniic_b200 import codecs
ctx = cb.Context(0)
img = cb.synth_image_host(203, 77, 1, 6)
sq = cb.synth_image_host(128, 128, 2, 6)
for flag in (cb._lib.KMEANS_NO_CULL, cb._lib.KMEANS_FORCE_CULL):
    s = cb.KMeansSession(ctx, cb.POINTS_RGB, 37, img, 203 * 77, flags=flag); s.reset(); s.run(3); s.get(); s.close()
    s = cb.KMeansSession(ctx, cb.POINTS_XYRGB, 300, img, 203 * 77, w=203, h_local=77, flags=flag & 1); s.reset(); s.run(3); s.get(); s.close()
keys, cnts = ctx.hist_rgb(img)
s = cb.KMeansSession(ctx, cb.POINTS_RGB, 16, np.stack([keys >> 16, (keys >> 8) & 255, keys & 255], 1).astype(np.uint8), len(keys),
                     weights=cnts.astype(np.uint32), flags=cb._lib.KMEANS_FORCE_CULL); s.reset(); s.run(3); s.get(); s.close()
ctx.cluster_colors(img, 16, max_iters=3)
ctx.voronoi_fill(np.array([[3, 4], [100, 50], [202, 76]], np.uint32), np.array([[1, 2, 3], [4, 5, 6], [7, 8, 9]], np.uint8), 203, 77)
for im in (img, sq):
    ctx.hilbert_xy(im.shape[1], im.shape[0]); ctx.hilbert_gather(im); d = ctx.delta(im); ctx.undelta(d, im.shape[1], im.shape[0]); ctx.hist_delta(im)
    for e in ("hufman", "delta", "hilbert(rle)", "voronoi(8)", "cluster-colors(8)"):
        c = codecs.Codec.from_str(ctx, e, 3); c.decode(c.encode(im))
# batch API (every kernel family through its batch entry point) and the fibonacci-coded stream of the parallel Huffman decoder
for kind, flag in (("rgb", cb._lib.KMEANS_NO_CULL), ("rgb", cb._lib.KMEANS_FORCE_CULL), ("xy", cb._lib.KMEANS_NO_CULL), ("xy", 0)):
    ims = [cb.synth_image_host(w, h, 9 + w, 4) for w, h in ((96, 40), (64, 64), (130, 17))]
    ss = [cb.KMeansSession(ctx, cb.POINTS_RGB, 12, im, im.shape[0] * im.shape[1], flags=flag) if kind == "rgb" else
          cb.KMeansSession(ctx, cb.POINTS_XYRGB, 12, im, im.shape[0] * im.shape[1], w=im.shape[1], h_local=im.shape[0], flags=flag) for im in ims]
    cb.kmeans_reset_batch(ss); cb.kmeans_run_batch(ss, 3)
    for s in ss:
        s.get(); s.close()
ctx.kmeans_rgb_batch([img, sq, img[:20]], 9, max_iters=3)
fib = [1, 1]
while len(fib) < 24:
    fib.append(fib[-1] + fib[-2])
px = np.repeat(np.arange(24), fib); np.random.default_rng(0).shuffle(px)
deep = np.stack([px, px * 3 % 256, px * 7 % 256], 1).astype(np.uint8)[:256 * (len(px) // 256)].reshape(-1, 256, 3)
for e in ("hufman", "delta", "hilbert(rle)"):
    c = codecs.Codec.from_str(ctx, e); assert np.array_equal(c.decode(c.encode(deep)), deep)
# curve-sharded stages: three ranges of a 128x128 image (aligned -> tile kernels) and one unaligned range
d_img = ctx.device_alloc(128 * 128 * 3); ctx.h2d(d_img, sq); d_o = ctx.device_alloc(128 * 128 * 6)
for i0, i1 in ((0, 4096), (4096, 12288), (12288, 16384), (5, 4173)):
    ctx.delta_range_device(d_img, 128, 128, i0, i1, d_o); ctx.hist_delta_range_device(d_img, 128, 128, i0, i1)
ctx.device_free(d_img); ctx.device_free(d_o)
ctx.sse(img, img[::-1].copy())
print("sanitize target done")
