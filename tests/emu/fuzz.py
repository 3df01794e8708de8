#!/usr/bin/env python
"""fuzz.py -- TEST INFRASTRUCTURE: random shapes / k / tie rules / kernel variants through the emulated kernels vs the oracle.

    python tests/emu/fuzz.py [seconds] [seed]

Every case is bit-exact or the script stops and prints the failing parameters."""
import ctypes
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import numpy as np  # noqa: E402
import build_emu  # noqa: E402
from cniic_b200 import _lib as L  # noqa: E402

L._lib = L._declare(ctypes.CDLL(build_emu.build()))
import cniic_b200 as cb  # noqa: E402
from cniic_b200 import codecs  # noqa: E402
import oracle as O  # noqa: E402

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
seed = int(sys.argv[2]) if len(sys.argv) > 2 else int(time.time())
rng = np.random.default_rng(seed)
ctx = cb.Context()
t0, cases = time.time(), 0


def image(w, h):
    mode = rng.integers(0, 4)
    if mode == 0:
        return cb.synth_image_host(w, h, int(rng.integers(1 << 30)), int(rng.integers(1, 9)))
    if mode == 1:  # few colours: ties, empty clusters, long runs
        return (rng.integers(0, 3, size=(h, w, 3)) * int(rng.integers(1, 120))).astype(np.uint8)
    if mode == 2:
        return rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)
    return np.full((h, w, 3), int(rng.integers(0, 256)), np.uint8)


def same(g_cen, g_w, g_asg, g_it, o, what):
    ok = g_it == o.iterations and np.array_equal(g_cen, o.centroids) and np.array_equal(g_w, o.weights) and np.array_equal(g_asg, o.assign)
    if not ok:
        print("MISMATCH", what, "seed", seed, flush=True)
        sys.exit(1)


while time.time() - t0 < budget:
    cases += 1
    w, h = int(rng.integers(1, 140)), int(rng.integers(1, 90))
    n = w * h
    img = image(w, h)
    k = int(rng.integers(1, min(n, 300) + 1))
    tie = int(rng.integers(0, 2))
    iters = int(rng.integers(1, 5))
    what = dict(w=w, h=h, k=k, tie=tie, iters=iters)
    sel = rng.integers(0, 6)
    if sel == 0:  # D = 3 both kernels, optionally weighted
        wts = rng.integers(1, 5000, n).astype(np.uint32) if rng.random() < 0.4 else None
        o = O.kmeans_rgb(img.reshape(-1, 3), k, counts=wts, mode=O.MODE_EXACT, tie=tie, max_iters=iters, allow_inactive=True)
        for flag in (L.KMEANS_NO_CULL, L.KMEANS_FORCE_CULL):
            s = cb.KMeansSession(ctx, cb.POINTS_RGB, k, img, n, weights=wts, tie=tie, flags=flag)
            s.reset(); st = s.run(iters); cen, ws, asg = s.get(); s.close()
            same(cen, ws, asg, st.iterations, o, dict(what, kind="rgb", flag=flag, weighted=wts is not None))
    elif sel == 1:  # D = 5 both kernels
        o = O.kmeans_xyrgb(img, k, mode=O.MODE_EXACT, tie=tie, max_iters=iters, allow_inactive=True)
        for flag in (L.KMEANS_NO_CULL, 0):
            s = cb.KMeansSession(ctx, cb.POINTS_XYRGB, k, img, n, w=w, h_local=h, tie=tie, flags=flag)
            s.reset(); st = s.run(iters); cen, ws, asg = s.get(); s.close()
            same(cen, ws, asg, st.iterations, o, dict(what, kind="xyrgb", flag=flag))
    elif sel == 2:  # batch of mixed sizes
        imgs = [image(int(rng.integers(1, 80)), int(rng.integers(1, 50))) for _ in range(int(rng.integers(1, 5)))]
        kk = int(rng.integers(1, min(min(i.shape[0] * i.shape[1] for i in imgs), 40) + 1))
        res = ctx.kmeans_rgb_batch(imgs, kk, max_iters=iters, tie=tie, allow_inactive=True)
        for im, g in zip(imgs, res):
            o = O.kmeans_rgb(im.reshape(-1, 3), kk, mode=O.MODE_EXACT, tie=tie, max_iters=iters, allow_inactive=True)
            same(g.centroids, g.weights, g.assign, g.iterations, o, dict(what, kind="batch", k=kk, shapes=[i.shape for i in imgs]))
    elif sel == 3:  # voronoi fill
        cxy = np.stack([rng.integers(0, w, k), rng.integers(0, h, k)], axis=1).astype(np.uint32)
        crgb = rng.integers(0, 256, size=(k, 3), dtype=np.uint8)
        if not np.array_equal(ctx.voronoi_fill(cxy, crgb, w, h), O.voronoi_fill(cxy, crgb, w, h)):
            print("MISMATCH fill", what, "seed", seed); sys.exit(1)
    elif sel == 4:  # integer stages
        ok = np.array_equal(ctx.delta(img), O.delta(img)) and np.array_equal(ctx.hilbert_gather(img), O.hilbert_gather(img))
        gk, gc = ctx.hist_delta(img); ok_, oc_ = O.hist_delta(O.delta(img))
        ok = ok and np.array_equal(gk, ok_) and np.array_equal(gc, oc_) and np.array_equal(ctx.undelta(O.delta(img), w, h), img)
        if not ok:
            print("MISMATCH stages", what, "seed", seed); sys.exit(1)
    else:  # whole codecs incl. the parallel decoders
        for expr, oenc, odec in (("hufman", O.encode_hufman, O.decode_hufman), ("delta", O.encode_delta, O.decode_delta),
                                 ("hilbert(rle)", O.encode_hilbert_rle, O.decode_hilbert_rle)):
            c = codecs.Codec.from_str(ctx, expr)
            data = c.encode(img)
            if data != oenc(img) or not np.array_equal(c.decode(data), img):
                print("MISMATCH codec", expr, what, "seed", seed); sys.exit(1)
            cut = int(rng.integers(8, len(data) + 1))
            g, o2 = c.decode(data[:cut]), odec(data[:cut])
            if (g is None) != (o2 is None) or (o2 is not None and not np.array_equal(g, o2)):
                print("MISMATCH codec prefix", expr, cut, what, "seed", seed); sys.exit(1)
print(f"fuzz ok: {cases} cases in {time.time() - t0:.0f} s, seed {seed}")
