"""GPU parity tests: every CUDA stage, called through the C ABI (ctypes), against the CPU oracle on the same seeded
inputs.  Bar: bit-exact (all arithmetic on this path is integer / byte work)."""
import numpy as np
import pytest

import oracle as O
import cniic_b200 as cb
from cniic_b200 import codecs

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = cb.Context()
    yield c
    c.close()


def same_kmeans(g, o):
    assert g.iterations == o.iterations
    assert np.array_equal(g.centroids, o.centroids)
    assert np.array_equal(g.weights, o.weights)
    assert np.array_equal(g.assign, o.assign)
    assert g.moved_last == o.moved_last and g.moved_total == o.moved_total
    assert g.empty_events == o.empty_events


# ---- K-means D = 3 (ColorCount, clusterc.rs:68-114) ----
@pytest.mark.parametrize("w,h,k", [(512, 512, 16), (97, 61, 5), (200, 150, 64), (333, 257, 256), (64, 33, 1), (50, 41, 300)])
@pytest.mark.parametrize("tie", [cb.TIE_KEEP_CURRENT, cb.TIE_LOWEST_INDEX])
@pytest.mark.parametrize("max_iters", [1, 4, 0])
def test_kmeans_rgb_per_pixel(ctx, w, h, k, tie, max_iters):
    img = cb.synth_image_host(w, h, 0xC0FFEE + k, 24)
    g = ctx.kmeans_rgb(img, k, max_iters=max_iters, tie=tie)
    o = O.kmeans_rgb(img, k, mode=O.MODE_EXACT, tie=tie, max_iters=max_iters)
    same_kmeans(g, o)
    if max_iters == 0:
        assert g.converged and g.moved_last == 0


@pytest.mark.parametrize("k", [16, 256])
def test_kmeans_rgb_weighted_unique_colours(ctx, k):  # the reference clusters (colour, count) pairs, clusterc.rs:19-28
    img = cb.synth_image_host(320, 200, 99, 24)
    keys, cnts = O.count_freqs_rgb(img)
    urgb = np.stack([keys >> 16, (keys >> 8) & 255, keys & 255], axis=1).astype(np.uint8)
    g = ctx.kmeans_rgb(urgb, k, counts=cnts.astype(np.uint32))
    o = O.kmeans_rgb(urgb, k, counts=cnts.astype(np.uint32), mode=O.MODE_EXACT)
    same_kmeans(g, o)
    assert int(g.weights.sum()) == 320 * 200


def test_kmeans_rgb_heavy_weights_need_u64(ctx):
    rgb = np.array([[255, 255, 255], [254, 255, 255], [0, 0, 0], [1, 0, 0], [128, 128, 128], [129, 128, 128]], np.uint8)
    cnt = np.array([4_000_000_000, 4_000_000_001, 3_999_999_999, 17, 4_294_967_295, 4_294_967_295], np.uint32)
    g = ctx.kmeans_rgb(rgb, 3, counts=cnt)
    o = O.kmeans_rgb(rgb, 3, counts=cnt, mode=O.MODE_EXACT)
    same_kmeans(g, o)


def _palette_points(seed):
    rng = np.random.default_rng(seed)
    npal, k, n = int(rng.integers(3, 10)), int(rng.integers(4, 16)), int(rng.integers(40, 400))
    palette = rng.integers(0, 256, size=(npal, 3), dtype=np.uint8)
    pts = palette[rng.integers(0, npal, size=n)]
    jit = rng.integers(-3, 4, size=pts.shape)
    mask = rng.random(n) < 0.3
    return np.clip(pts.astype(int) + jit * mask[:, None], 0, 255).astype(np.uint8), k


@pytest.mark.parametrize("seed", [0, 3, 4, 9, 12, 18, 34, 46, 52, 58])
def test_kmeans_rgb_empty_clusters_are_repaired_deterministically(ctx, seed):  # kmeans.rs:117-134 stand-in
    pts, k = _palette_points(seed)
    events = 0
    for tie in (cb.TIE_KEEP_CURRENT, cb.TIE_LOWEST_INDEX):
        g = ctx.kmeans_rgb(pts, k, max_iters=8, tie=tie, allow_inactive=True)
        o = O.kmeans_rgb(pts, k, mode=O.MODE_EXACT, tie=tie, max_iters=8, allow_inactive=True)
        events += o.empty_events
        same_kmeans(g, o)
        # the same through the colour-sorted culled kernel (its repair path maps through the permutation)
        s = cb.KMeansSession(ctx, cb.POINTS_RGB, k, pts, len(pts), tie=tie, flags=cb._lib.KMEANS_FORCE_CULL)
        s.reset()
        st = s.run(8)
        cen, wsum, asg = s.get()
        s.close()
        assert st.iterations == o.iterations and st.empty_events == o.empty_events
        assert np.array_equal(cen, o.centroids) and np.array_equal(asg, o.assign) and np.array_equal(wsum, o.weights)
        assert g.status == o.status  # kmeans.rs:41-57 too-few-active check (seed 9 trips it)
    assert events > 0


@pytest.mark.parametrize("seed", [0, 1, 7, 10, 14, 15, 16, 19])
def test_kmeans_xyrgb_empty_clusters(ctx, seed):  # kmeans.rs:117-134 stand-in on the (x,y,r,g,b) path
    rng = np.random.default_rng(seed)
    w, h = int(rng.integers(4, 24)), int(rng.integers(3, 16))
    k = int(rng.integers(2, max(3, w * h // 2)))
    img = (rng.integers(0, 3, size=(h, w, 3)) * 100).astype(np.uint8)
    events = 0
    for tie in (cb.TIE_KEEP_CURRENT, cb.TIE_LOWEST_INDEX):
        g = ctx.kmeans_xyrgb(img, k, max_iters=8, tie=tie, allow_inactive=True)
        o = O.kmeans_xyrgb(img, k, mode=O.MODE_EXACT, tie=tie, max_iters=8, allow_inactive=True)
        same_kmeans(g, o)
        events += o.empty_events
    assert events > 0


def test_kmeans_errors(ctx):
    img = cb.synth_image_host(8, 2, 1, 2)
    with pytest.raises(cb.CniicError) as e:  # kmeans.rs:67-68
        ctx.kmeans_rgb(img, 17)
    assert e.value.code == cb.ERR_TOO_FEW_POINTS
    with pytest.raises(cb.CniicError) as e:
        ctx.kmeans_rgb(img, 0)
    assert e.value.code == cb.ERR_BAD_ARG
    with pytest.raises(cb.CniicError) as e:
        ctx.kmeans_xyrgb(img, cb.MAX_K + 1)
    assert e.value.code == cb.ERR_BAD_ARG


# ---- K-means D = 5 (ColorPos, clusterc.rs:148-153,200-248) ----
@pytest.mark.parametrize("w,h,k", [(64, 48, 16), (257, 131, 64), (300, 200, 256), (1024, 96, 100), (31, 9, 7), (640, 360, 2048)])
@pytest.mark.parametrize("tie", [cb.TIE_KEEP_CURRENT, cb.TIE_LOWEST_INDEX])
def test_kmeans_xyrgb(ctx, w, h, k, tie):
    img = cb.synth_image_host(w, h, 0xC0FFEE + 3, max(4, k // 8))
    for max_iters in ((1, 3) if k >= 256 else (1, 3, 0)):
        g = ctx.kmeans_xyrgb(img, k, max_iters=max_iters, tie=tie)
        o = O.kmeans_xyrgb(img, k, mode=O.MODE_EXACT, tie=tie, max_iters=max_iters)
        same_kmeans(g, o)


@pytest.mark.parametrize("w,h,k", [(1, 97, 5), (97, 1, 5), (7, 3, 21), (513, 65, 4096), (1030, 40, 1000)])
def test_kmeans_xyrgb_degenerate_shapes_and_max_k(ctx, w, h, k):
    img = cb.synth_image_host(w, h, 3, 9)
    for tie in (cb.TIE_KEEP_CURRENT, cb.TIE_LOWEST_INDEX):
        g = ctx.kmeans_xyrgb(img, k, max_iters=3, tie=tie, allow_inactive=True)
        o = O.kmeans_xyrgb(img, k, mode=O.MODE_EXACT, tie=tie, max_iters=3, allow_inactive=True)
        same_kmeans(g, o)


@pytest.mark.parametrize("n,k", [(1, 1), (7, 7), (9, 2), (4096 * 3, 4096), (100003, 1000)])
def test_kmeans_rgb_degenerate_sizes_and_max_k(ctx, n, k):
    pts = cb.synth_image_host(251, (n + 250) // 251, 5, 7).reshape(-1, 3)[:n]
    for flag in ("NO_CULL", "FORCE_CULL"):
        s = cb.KMeansSession(ctx, cb.POINTS_RGB, k, pts, n, flags=getattr(cb._lib, "KMEANS_" + flag))
        s.reset()
        st = s.run(3)
        cen, wsum, asg = s.get()
        s.close()
        o = O.kmeans_rgb(pts, k, mode=O.MODE_EXACT, max_iters=3, allow_inactive=True)
        assert st.iterations == o.iterations and np.array_equal(cen, o.centroids) and np.array_equal(asg, o.assign)


def test_kmeans_xyrgb_flat_image_has_exact_ties(ctx):
    """A constant-colour image makes many pixels equidistant from two centroids: the documented near-tie cases."""
    img = np.full((40, 64, 3), 77, np.uint8)
    for tie in (cb.TIE_KEEP_CURRENT, cb.TIE_LOWEST_INDEX):
        g = ctx.kmeans_xyrgb(img, 8, tie=tie)
        o = O.kmeans_xyrgb(img, 8, mode=O.MODE_EXACT, tie=tie)
        same_kmeans(g, o)


@pytest.mark.parametrize("w,h,k", [(257, 131, 64), (640, 360, 2048), (100, 70, 300)])
def test_kmeans_xyrgb_brute_force_kernel(ctx, w, h, k):
    """The non-culled D = 5 kernel (CNIIC_KMEANS_NO_CULL) is kept for roofline measurements; it must agree too."""
    img = cb.synth_image_host(w, h, 5, max(4, k // 8))
    s = cb.KMeansSession(ctx, cb.POINTS_XYRGB, k, img, w * h, w=w, h_local=h, flags=cb._lib.KMEANS_NO_CULL)
    s.reset()
    st = s.run(3)
    cen, wts, asg = s.get()
    s.close()
    o = O.kmeans_xyrgb(img, k, mode=O.MODE_EXACT, max_iters=3)
    assert st.iterations == 3 and np.array_equal(cen, o.centroids) and np.array_equal(asg, o.assign) and np.array_equal(wts, o.weights)


@pytest.mark.parametrize("flag", ["NO_CULL", "FORCE_CULL"])
@pytest.mark.parametrize("n,k,weighted", [(5000, 16, False), (70000, 256, False), (33333, 300, True), (2048 * 3 + 5, 64, True)])
def test_kmeans_rgb_both_kernels(ctx, n, k, weighted, flag):
    """Small D = 3 problems default to the brute-force kernel; both it and the colour-sorted culled kernel must agree with the oracle."""
    rng = np.random.default_rng(n)
    pts = cb.synth_image_host(256, (n + 255) // 256, 77, 9).reshape(-1, 3)[:n]
    wts = rng.integers(1, 1000, n).astype(np.uint32) if weighted else None
    s = cb.KMeansSession(ctx, cb.POINTS_RGB, k, pts, n, weights=wts, flags=getattr(cb._lib, "KMEANS_" + flag))
    s.reset()
    st = s.run(4)
    cen, wsum, asg = s.get()
    s.close()
    o = O.kmeans_rgb(pts, k, counts=wts, mode=O.MODE_EXACT, max_iters=4)
    assert st.iterations == o.iterations and np.array_equal(cen, o.centroids) and np.array_equal(asg, o.assign)
    assert np.array_equal(wsum, o.weights)


def test_kmeans_session_reuse(ctx):
    img = cb.synth_image_host(128, 64, 8, 8)
    s = cb.KMeansSession(ctx, cb.POINTS_XYRGB, 32, img, 128 * 64, w=128, h_local=64)
    outs = []
    for _ in range(2):
        s.reset()
        st = s.run(5)
        cen, wts, asg = s.get()
        outs.append((cen, wts, asg, st.iterations))
    s.close()
    o = O.kmeans_xyrgb(img, 32, max_iters=5)
    for cen, wts, asg, it in outs:
        assert it == 5 and np.array_equal(cen, o.centroids) and np.array_equal(asg, o.assign)


# ---- cluster-colors pipeline (clusterc.rs:18-52) ----
@pytest.mark.parametrize("k", [16, 64])
def test_cluster_colors_pipeline(ctx, k):
    img = cb.synth_image_host(256, 192, 21, 12)
    out, cen, st = ctx.cluster_colors(img, k)
    oout, ocen, oit = O.cluster_colors(img, k)
    assert st.iterations == oit
    assert np.array_equal(cen, ocen) and np.array_equal(out, oout)


@pytest.mark.parametrize("seed", [0, 1, 2, 3, 5, 8, 13])
def test_cluster_colors_unique_colour_path_edge_cases(ctx, seed):
    """The unique colours reach K-means deduplicated and Morton-sorted straight out of the histogram (no sort); the canonical order
    (init chunks, empty-cluster rule: kmeans.rs:61-108, 117-134 stand-in) stays ascending r<<16|g<<8|b.  Palette images with few
    colours, heavy counts, k close to the number of colours, odd sizes (unaligned 4-pixel groups), flat runs."""
    rng = np.random.default_rng(seed)
    w, h = int(rng.integers(3, 90)), int(rng.integers(3, 70))
    npal = int(rng.integers(2, 40))
    palette = rng.integers(0, 256, size=(npal, 3), dtype=np.uint8)
    if seed % 2:  # clustered palette: many near-duplicates -> empty clusters
        palette = np.clip(palette[rng.integers(0, max(1, npal // 4), npal)].astype(int) + rng.integers(-2, 3, (npal, 3)), 0, 255).astype(np.uint8)
    img = palette[rng.integers(0, npal, size=(h, w))]
    img[: h // 3] = palette[0]  # a flat area: runs of equal neighbours
    u = len(np.unique(img.reshape(-1, 3), axis=0))
    for k in sorted({1, max(1, u // 2), u}):
        for max_iters in (1, 0):
            oout, ocen, oit = O.cluster_colors(img, k, max_iters=max_iters)
            out, cen, st = ctx.cluster_colors(img, k, max_iters=max_iters)
            assert st.iterations == oit, (k, max_iters)
            assert np.array_equal(cen, ocen) and np.array_equal(out, oout), (k, max_iters)
    with pytest.raises(cb.CniicError) as e:  # kmeans.rs:67-68: fewer points (unique colours) than clusters
        ctx.cluster_colors(img, u + 1)
    assert e.value.code == cb.ERR_TOO_FEW_POINTS
    # device-resident form + the centroid-only host form agree with the full one
    d = ctx.device_alloc(img.nbytes)
    ctx.h2d(d, img)
    cen_d, nu, st = ctx.cluster_colors_device(d, w * h, max(1, u // 2), max_iters=3)
    ctx.device_free(d)
    _, ocen3, _ = O.cluster_colors(img, max(1, u // 2), max_iters=3)
    none_img, cen_h, _ = ctx.cluster_colors(img, max(1, u // 2), max_iters=3, want_image=False)
    assert nu == u and none_img is None and np.array_equal(cen_d, ocen3) and np.array_equal(cen_h, ocen3)


def test_hist_rgb_and_recolor(ctx):
    img = cb.synth_image_host(300, 100, 4, 9)
    keys, cnts = ctx.hist_rgb(img)
    okeys, ocnts = O.count_freqs_rgb(img)
    assert np.array_equal(keys, okeys) and np.array_equal(cnts, ocnts)
    asg = (np.arange(len(keys)) % 5).astype(np.uint16)
    cen = np.arange(15, dtype=np.uint8).reshape(5, 3) * 10
    out = ctx.recolor_rgb(img, keys, asg, cen)
    lut = {int(k): cen[a] for k, a in zip(keys, asg)}
    flat = img.reshape(-1, 3).astype(np.uint32)
    exp = np.array([lut[int((p[0] << 16) | (p[1] << 8) | p[2])] for p in flat[:2000]], np.uint8)
    assert np.array_equal(out.reshape(-1, 3)[:2000], exp)
    # extremes: a single colour, and all-distinct colours
    k1, c1 = ctx.hist_rgb(np.full((10, 10, 3), 200, np.uint8))
    assert k1.tolist() == [(200 << 16) | (200 << 8) | 200] and c1.tolist() == [100]
    k0, c0 = ctx.hist_rgb(np.zeros((0, 3), np.uint8))
    assert len(k0) == 0


# ---- voronoi fill (clusterc.rs:179-186) ----
@pytest.mark.parametrize("w,h,k", [(64, 64, 1), (200, 130, 17), (513, 97, 300), (640, 360, 2048)])
def test_voronoi_fill(ctx, w, h, k):
    rng = np.random.default_rng(k)
    cxy = np.stack([rng.integers(0, w, k), rng.integers(0, h, k)], axis=1).astype(np.uint32)
    cxy[k // 2] = cxy[0]  # duplicate centroid: the first one must win (min_by_key returns the first minimum)
    crgb = rng.integers(0, 256, size=(k, 3), dtype=np.uint8)
    assert np.array_equal(ctx.voronoi_fill(cxy, crgb, w, h), O.voronoi_fill(cxy, crgb, w, h))


def test_voronoi_fill_ties_and_outside_centroids(ctx):
    cxy = np.array([[10, 10], [30, 10], [10, 30], [30, 30], [500, 500], [20, 20], [20, 20]], np.uint32)
    crgb = (np.arange(21, dtype=np.uint8).reshape(7, 3) + 1) * 9
    assert np.array_equal(ctx.voronoi_fill(cxy, crgb, 41, 41), O.voronoi_fill(cxy, crgb, 41, 41))
    with pytest.raises(cb.CniicError):
        ctx.voronoi_fill(np.zeros((0, 2), np.uint32), np.zeros((0, 3), np.uint8), 4, 4)  # clusterc.rs:182-184 unwrap


# ---- Hilbert / delta / histograms (hilbert.rs:34-43, hilbertc.rs:449-477) ----
@pytest.mark.parametrize("w,h", [(1, 1), (1, 9), (9, 1), (4, 4), (64, 64), (5, 3), (3, 5), (100, 7), (123, 77), (256, 256), (640, 360)])
def test_hilbert_delta_stages(ctx, w, h):
    img = cb.synth_image_host(w, h, 17, 4)
    assert np.array_equal(ctx.hilbert_xy(w, h), O.hilbert_xy(w, h))
    assert np.array_equal(ctx.hilbert_gather(img), O.hilbert_gather(img))
    d = ctx.delta(img)
    assert np.array_equal(d, O.delta(img))
    assert np.array_equal(ctx.undelta(d, w, h), img)
    keys, cnts = ctx.hist_delta(img)
    okeys, ocnts = O.hist_delta(O.delta(img))
    assert np.array_equal(keys, okeys) and np.array_equal(cnts, ocnts)


def test_hist_delta_extremes(ctx):
    flat = np.full((64, 64, 3), 9, np.uint8)  # all-equal image: two symbols (first colour, zero diff)
    keys, cnts = ctx.hist_delta(flat)
    assert sorted(cnts.tolist()) == [1, 64 * 64 - 1]
    rng = np.random.default_rng(0)
    noise = rng.integers(0, 256, size=(128, 128, 3), dtype=np.uint8)
    keys, cnts = ctx.hist_delta(noise)
    okeys, ocnts = O.hist_delta(O.delta(noise))
    assert np.array_equal(keys, okeys) and np.array_equal(cnts, ocnts)


def test_sse(ctx):
    a = cb.synth_image_host(333, 77, 1, 5)
    b = cb.synth_image_host(333, 77, 2, 5)
    assert ctx.sse(a, b) == O.sse(a, b)
    assert ctx.sse(a, a) == 0
    assert ctx.sse(a, b) / (333 * 77) == pytest.approx(O.mse(a, b), rel=1e-9)  # bench.rs:95-104, tolerance 1e-6 in north_star


# ---- whole codecs: byte streams equal the oracle's, decoders invert (codec.rs:14-19, bench.rs:45-59) ----
@pytest.mark.parametrize("expr,oenc,odec", [
    ("hufman", lambda im: O.encode_hufman(im), O.decode_hufman),
    ("delta", lambda im: O.encode_delta(im), O.decode_delta),
    ("hilbert(rle)", lambda im: O.encode_hilbert_rle(im), O.decode_hilbert_rle),
    ("voronoi(24)", lambda im: O.encode_voronoi(im, 24), O.decode_voronoi),
    ("cluster-colors(16)", lambda im: O.encode_cluster_colors(im, 16), O.decode_hufman),
])
@pytest.mark.parametrize("w,h", [(96, 64), (57, 31)])
def test_codec_streams(ctx, expr, oenc, odec, w, h):
    img = cb.synth_image_host(w, h, 31, 5)
    if expr == "hilbert(rle)":
        img = (img // 64) * 64  # give the run-length coder some runs
    c = codecs.Codec.from_str(ctx, expr)
    data = c.encode(img)
    assert data == oenc(img)
    dec = c.decode(data)
    assert np.array_equal(dec, odec(data))
    if c.is_lossless():
        assert np.array_equal(dec, img)  # bench.rs:57-59
    else:
        assert ctx.sse(img, dec) / (w * h) == pytest.approx(O.mse(img, odec(data)), rel=1e-6)


@pytest.mark.parametrize("expr", ["hufman", "delta"])
def test_codec_streams_many_symbols(ctx, expr):
    """Noise image: ~all pixels are distinct symbols (long codes, payload words shared between threads everywhere)."""
    rng = np.random.default_rng(11)
    img = rng.integers(0, 256, size=(160, 224, 3), dtype=np.uint8)
    img[:40] = (img[:40] // 128) * 128  # plus a region with very frequent symbols (short codes)
    c = codecs.Codec.from_str(ctx, expr)
    data = c.encode(img)
    assert data == (O.encode_hufman(img) if expr == "hufman" else O.encode_delta(img))
    assert np.array_equal(c.decode(data), img)


def test_codec_decode_rejects_malformed(ctx):
    img = cb.synth_image_host(32, 16, 5, 3)
    for expr in ("hufman", "delta", "voronoi(8)", "hilbert(rle)"):
        c = codecs.Codec.from_str(ctx, expr)
        data = c.encode(img)
        assert c.decode(data[:len(data) // 2]) is None  # Codec::decode -> None
        assert c.decode(data[:5]) is None
    with pytest.raises(cb.CniicError):
        ctx.codec_encode("zip(dict)", img)


def test_single_colour_image_codecs(ctx):  # huf.rs:139-142 zero-length code
    img = np.full((8, 8, 3), 50, np.uint8)
    c = codecs.Hufman(ctx)
    data = c.encode(img)
    assert data == O.encode_hufman(img) and len(data) == 8 + 12
    assert np.array_equal(c.decode(data), img)


# ---- batches of independent images: one launch per stage for the whole batch (bench.rs:27, BASELINE config 4) ----
@pytest.mark.parametrize("k,sizes", [(16, [(64, 48), (64, 48), (64, 48)]), (64, [(128, 96), (40, 30), (97, 61), (128, 96), (16, 4)]), (5, [(33, 9)])])
@pytest.mark.parametrize("max_iters", [3, 0])
def test_kmeans_rgb_batch_equals_separate_runs(ctx, k, sizes, max_iters):
    imgs = [cb.synth_image_host(w, h, 100 + i, 6) for i, (w, h) in enumerate(sizes)]
    res = ctx.kmeans_rgb_batch(imgs, k, max_iters=max_iters)
    assert len(res) == len(imgs)
    for im, g in zip(imgs, res):
        o = O.kmeans_rgb(im, k, mode=O.MODE_EXACT, max_iters=max_iters)  # images converge after different numbers of passes
        same_kmeans(g, o)


@pytest.mark.parametrize("kind,flag", [("rgb", "NO_CULL"), ("rgb", "FORCE_CULL"), ("xyrgb", "NO_CULL"), ("xyrgb", 0)])
def test_kmeans_session_batch_all_kernel_variants(ctx, kind, flag):
    """Batched launches of every assign kernel family (brute / culled, D = 3 / D = 5), mixed image sizes, with an
    empty-cluster repair inside the batch, against one oracle run per image."""
    rng = np.random.default_rng(5)
    sizes = [(96, 40), (64, 64), (130, 17), (24, 12)]
    imgs = [cb.synth_image_host(w, h, 7 + i, 5) for i, (w, h) in enumerate(sizes)]
    imgs[3] = (rng.integers(0, 3, size=(12, 24, 3)) * 100).astype(np.uint8)  # few colours: empty clusters appear
    k, fl = 12, (getattr(cb._lib, "KMEANS_" + flag) if flag else 0)
    ss = []
    for im in imgs:
        h, w = im.shape[:2]
        if kind == "rgb":
            ss.append(cb.KMeansSession(ctx, cb.POINTS_RGB, k, im, w * h, flags=fl))
        else:
            ss.append(cb.KMeansSession(ctx, cb.POINTS_XYRGB, k, im, w * h, w=w, h_local=h, flags=fl))
    for rounds in range(2):  # a second reset + run on the same sessions gives the same answer
        cb.kmeans_reset_batch(ss)
        sts = cb.kmeans_run_batch(ss, 3)
        sts2 = cb.kmeans_run_batch(ss, 2)  # continue: 3 + 2 iterations in two calls
        for im, s, st, st2 in zip(imgs, ss, sts, sts2):
            o = (O.kmeans_rgb if kind == "rgb" else O.kmeans_xyrgb)(im, k, mode=O.MODE_EXACT, max_iters=5, allow_inactive=True)
            cen, wts, asg = s.get()
            assert st2.iterations == o.iterations and st.iterations == min(3, o.iterations)
            assert np.array_equal(cen, o.centroids) and np.array_equal(wts, o.weights) and np.array_equal(asg, o.assign)
            assert st2.empty_events == o.empty_events and st2.moved_total == o.moved_total
    for s in ss:
        s.close()


def test_kmeans_batch_rejects_mixed_sessions(ctx):
    a = cb.KMeansSession(ctx, cb.POINTS_RGB, 4, cb.synth_image_host(16, 8, 1, 3), 128)
    b = cb.KMeansSession(ctx, cb.POINTS_RGB, 5, cb.synth_image_host(16, 8, 2, 3), 128)
    with pytest.raises(cb.CniicError) as e:
        cb.kmeans_reset_batch([a, b])
    assert e.value.code == cb.ERR_BAD_ARG
    a.close(); b.close()


# ---- parallel Huffman decoder (huffdec.cu) against the sequential walk of huf.rs:187-206 ----
def _fibonacci_image(depth, seed):
    """Colour i occurs fib(i) times: the Huffman tree degenerates into a chain, code lengths 1 .. depth-1 bits."""
    fib = [1, 1]
    while len(fib) < depth:
        fib.append(fib[-1] + fib[-2])
    rng = np.random.default_rng(seed)
    pal = rng.integers(0, 256, size=(depth, 3), dtype=np.uint8)
    pal[:, 0] = np.arange(depth)  # distinct colours
    px = np.repeat(np.arange(depth), fib)
    rng.shuffle(px)
    w = 256
    h = len(px) // w
    return pal[px[:w * h]].reshape(h, w, 3)


@pytest.mark.parametrize("expr", ["hufman", "delta"])
def test_huffman_decoder_long_codes_over_many_chunks(ctx, expr):
    """Code words from 1 to 20+ bits, several 64-Kbit chunks: wrong starting guesses must re-synchronise inside CTAs and
    across them, and the symbol counts must add up to the exact output positions."""
    img = _fibonacci_image(25, 1)
    c = codecs.Codec.from_str(ctx, expr)
    data = c.encode(img)
    assert data == (O.encode_hufman(img) if expr == "hufman" else O.encode_delta(img))
    assert len(data) * 8 > 3 * 65536
    assert np.array_equal(c.decode(data), img)


def test_huffman_decoder_two_symbols_one_bit_each(ctx):
    """256 code words per subsequence: the densest stream there is (every bit position is a valid start)."""
    rng = np.random.default_rng(2)
    img = np.where(rng.random((96, 128, 1)) < 0.5, np.uint8(10), np.uint8(200)).repeat(3, axis=2).astype(np.uint8)
    c = codecs.Hufman(ctx)
    data = c.encode(img)
    assert data == O.encode_hufman(img)
    assert np.array_equal(c.decode(data), img)


@pytest.mark.parametrize("expr,odec", [("hufman", O.decode_hufman), ("delta", O.decode_delta)])
def test_huffman_decoder_truncated_at_every_byte(ctx, expr, odec):
    """Same verdict as the sequential decoder for every prefix of a stream: None while code words are missing
    (huf.rs:190-204), the image as soon as all n are complete (trailing padding bits are ignored)."""
    img = _fibonacci_image(12, 4)[:2, :24]
    c = codecs.Codec.from_str(ctx, expr)
    data = c.encode(img)
    for cut in range(8, len(data) + 1):
        g, o = c.decode(data[:cut]), odec(data[:cut])
        assert (g is None) == (o is None), cut
        if o is not None:
            assert np.array_equal(g, o), cut
    # garbage appended after the payload is ignored by both
    assert np.array_equal(c.decode(data + b"\xff\x00\xaa"), odec(data + b"\xff\x00\xaa"))


# ---- exact RLE along the Hilbert stream on the GPU (hilbertc.rs:99-196, 304-333) ----
@pytest.mark.parametrize("w,h", [(70, 70), (255, 3), (64, 64), (510, 1), (1, 511), (128, 96)])
def test_rle_long_runs_are_cut_at_255(ctx, w, h):
    """Runs longer than 255 pixels (and exact multiples of 255) across thread and CTA boundaries of the scan."""
    rng = np.random.default_rng(w * 1000 + h)
    c = codecs.Codec.from_str(ctx, "hilbert(rle)")
    flat = np.full((h, w, 3), 7, np.uint8)  # one run of w*h pixels
    imgs = [flat]
    lin_len = w * h
    # runs of chosen lengths laid out ALONG THE CURVE, so their lengths are exact
    xy = O.hilbert_xy(w, h)
    runs, pos = [], 0
    for L in [255, 1, 256, 510, 3, 254, 4096, 2, 765, 4097]:
        if pos + L > lin_len:
            break
        runs.append((pos, L))
        pos += L
    img = np.zeros((h, w, 3), np.uint8)
    colours = rng.integers(1, 255, size=(len(runs) + 1, 3), dtype=np.uint8)
    for i, (p0, L) in enumerate(runs):
        colours[i, 0] = i  # neighbouring runs differ
        img[xy[p0:p0 + L, 1], xy[p0:p0 + L, 0]] = colours[i]
    img[xy[pos:, 1], xy[pos:, 0]] = (250, 250, 250)
    imgs.append(img)
    for im in imgs:
        data = c.encode(im)
        assert data == O.encode_hilbert_rle(im)
        assert np.array_equal(c.decode(data), im)


def test_rle_decoder_verdicts_match_the_sequential_decoder(ctx):
    c = codecs.Codec.from_str(ctx, "hilbert(rle)")
    img = (cb.synth_image_host(24, 10, 3, 3) // 128) * 128
    data = c.encode(img)
    for cut in range(8, len(data) + 1):  # every prefix: None until the records cover all pixels
        g, o = c.decode(data[:cut]), O.decode_hilbert_rle(data[:cut])
        assert (g is None) == (o is None), cut
        if o is not None:
            assert np.array_equal(g, o), cut
    hdr, body = data[:8], data[8:]
    rec = lambda cnt, rgb, ln=3: bytes([cnt]) + int(ln).to_bytes(8, "little") + bytes(rgb)
    cases = [
        hdr + rec(0, (1, 2, 3)) + body,                      # zero-count record: paints nothing
        hdr + body + rec(9, (1, 2, 3), ln=4),                # malformed record BEHIND the last needed one: never read
        hdr + rec(5, (1, 2, 3), ln=4) + body,                # malformed record that is needed
        hdr + body[:12] + rec(200, (9, 9, 9)) + body[12:],   # more pixels than the image holds: the surplus is dropped
        hdr + rec(255, (4, 5, 6)),                           # exactly one record, 240 pixels needed
    ]
    for i, s in enumerate(cases):
        g, o = c.decode(s), O.decode_hilbert_rle(s)
        assert (g is None) == (o is None), i
        if o is not None:
            assert np.array_equal(g, o), i


# ---- first versions of the culled kernels stay selectable (A/B measurements) and agree with the second versions ----
@pytest.mark.parametrize("var,kind", [("CNIIC_RGB_CULL_V1", "rgb"), ("CNIIC_XY_CULL_V1", "xyrgb")])
def test_first_kernel_versions_agree(ctx, var, kind, monkeypatch):
    img = cb.synth_image_host(300, 170, 77, 12)
    n, k = 300 * 170, 96
    outs = []
    for v1 in (False, True):
        if v1:
            monkeypatch.setenv(var, "1")
        else:
            monkeypatch.delenv(var, raising=False)
        if kind == "rgb":
            s = cb.KMeansSession(ctx, cb.POINTS_RGB, k, img, n, flags=cb._lib.KMEANS_FORCE_CULL)
        else:
            s = cb.KMeansSession(ctx, cb.POINTS_XYRGB, k, img, n, w=300, h_local=170)
        s.reset()
        st = s.run(4)
        outs.append((s.get(), st.iterations, st.moved_total, st.pairs_scored))
        s.close()
    (c0, w0, a0), it0, m0, p0 = outs[0]
    (c1, w1, a1), it1, m1, p1 = outs[1]
    assert it0 == it1 and m0 == m1 and np.array_equal(c0, c1) and np.array_equal(w0, w1) and np.array_equal(a0, a1)
    o = (O.kmeans_rgb if kind == "rgb" else O.kmeans_xyrgb)(img, k, mode=O.MODE_EXACT, max_iters=4)
    assert np.array_equal(c0, o.centroids) and np.array_equal(a0, o.assign)
    if kind == "rgb":
        assert p0 <= p1  # the warp-level culling of the second version never scores more pairs


def test_hist_delta_counter_spill(ctx):
    """A flat 512x512 image puts 262 143 identical symbols through a few CTAs: the packed 15-bit shared counters must hand
    their full 2^15 blocks to the global bins without losing or double counting (and the neighbouring field stays intact)."""
    img = np.full((512, 512, 3), 40, np.uint8)
    img[100:140, 7:300] = 41     # some (+1,+1,+1) / (-1,-1,-1) symbols: the neighbour fields of (0,0,0)
    img[300, 300] = (200, 0, 90)  # and two symbols outside the shared cube
    keys, cnts = ctx.hist_delta(img)
    okeys, ocnts = O.hist_delta(O.delta(img))
    assert np.array_equal(keys, okeys) and np.array_equal(cnts, ocnts)
    assert int(cnts.sum()) == 512 * 512 and int(cnts.max()) > 200000


@pytest.mark.parametrize("kind", ["rgb", "xyrgb"])
def test_kernel_versions_agree_on_a_shard(ctx, kind, monkeypatch):
    """One rank's view of a row-sharded run (rows 37..96 of a 160-row image, explicit initial centroids, one iteration):
    both kernel versions must produce the same partial sums -> centroids, weights and assignment for the shard."""
    w, h, y0, hl, k = 192, 160, 37, 60, 40
    img = cb.synth_image_host(w, h, 5, 9)
    shard = np.ascontiguousarray(img[y0:y0 + hl])
    rng = np.random.default_rng(1)
    if kind == "rgb":
        init = rng.integers(0, 256, size=(k, 3)).astype(np.int32)
        kw = dict(n_total=w * h, first_index=y0 * w, flags=cb._lib.KMEANS_FORCE_CULL)
        kid = cb.POINTS_RGB
    else:
        init = np.concatenate([rng.integers(0, w, (k, 1)), rng.integers(0, h, (k, 1)), rng.integers(0, 256, (k, 3))], axis=1).astype(np.int32)
        kw = dict(n_total=w * h, first_index=y0 * w, w=w, h_local=hl, y0=y0)
        kid = cb.POINTS_XYRGB
    outs = []
    for var in (None, "CNIIC_RGB_CULL_V1", "CNIIC_XY_CULL_V1"):
        monkeypatch.delenv("CNIIC_RGB_CULL_V1", raising=False)
        monkeypatch.delenv("CNIIC_XY_CULL_V1", raising=False)
        if var:
            monkeypatch.setenv(var, "1")
        s = cb.KMeansSession(ctx, kid, k, shard, w * hl, **kw)
        s.reset(init)
        st = s.run(1)
        outs.append(s.get() + (st.moved_last,))
        s.close()
    for o in outs[1:]:
        assert all(np.array_equal(a, b) for a, b in zip(outs[0][:3], o[:3])) and outs[0][3] == o[3]
    assert int(outs[0][1].sum()) == w * hl  # every pixel of the shard is counted exactly once


# ---- curve-sharded integer stages: ranges of the Hilbert curve processed separately add up to the whole ----
@pytest.mark.parametrize("w,h,world", [(128, 128, 2), (128, 128, 3), (256, 256, 8), (100, 70, 3), (64, 64, 5)])
def test_curve_sharded_delta_and_histogram(ctx, w, h, world):
    from cniic_b200 import dist as cdist
    img = cb.synth_image_host(w, h, 23, 6)
    n = w * h
    d_img = ctx.device_alloc(n * 3)
    ctx.h2d(d_img, img)
    d_out = ctx.device_alloc(n * 6)
    parts, covered = [], 0
    whole = np.zeros((n, 3), np.int16)
    for r in range(world):
        i0, i1 = cdist.curve_shard(n, world, r)
        assert i0 == covered and i1 >= i0
        covered = i1
        if i1 > i0:
            ctx.delta_range_device(d_img, w, h, i0, i1, d_out)  # written relative to i0
            piece = np.zeros((i1 - i0, 3), np.int16)
            ctx.d2h(piece, d_out)
            whole[i0:i1] = piece
        parts.append(ctx.hist_delta_range_device(d_img, w, h, i0, i1))
    assert covered == n
    assert np.array_equal(whole, O.delta(img))
    keys, cnts = cdist.merge_histograms(parts)
    okeys, ocnts = O.hist_delta(O.delta(img))
    assert np.array_equal(keys, okeys) and np.array_equal(cnts, ocnts)
    # unaligned ranges fall back to the per-index kernel and still agree
    i0, i1 = 5, min(n, 4096 + 77)
    ctx.delta_range_device(d_img, w, h, i0, i1, d_out)
    piece = np.zeros((i1 - i0, 3), np.int16)
    ctx.d2h(piece, d_out)
    assert np.array_equal(piece, O.delta(img)[i0:i1])
    ctx.device_free(d_img); ctx.device_free(d_out)


def test_truncated_streams_follow_the_reference_decoders(ctx):
    """What the reference does with a short stream differs per codec (SURVEY 8b error behaviour): Hufman::decode returns None
    (hufc.rs:24-36); Delta and Hilbert-RLE zip their symbol iterator with the curve over a zero image, so a payload that just
    ENDS yields the pixels it reached and zeros elsewhere (hilbertc.rs:55-79, 417-431); a record cut in the middle / a bad trie
    panics there -> None here."""
    img = cb.synth_image_host(40, 24, 9, 4)
    n = 40 * 24
    xy = O.hilbert_xy(40, 24)
    # delta: cut the payload (the trie is intact: it ends where the first payload byte starts; cut a few bytes off the end)
    c = codecs.Codec.from_str(ctx, "delta")
    data = c.encode(img)
    cut = data[:len(data) - 40]
    g, o = c.decode(cut), O.decode_delta(cut)
    assert g is not None and np.array_equal(g, o)
    lin = g[xy[:, 1], xy[:, 0]]
    reached = int(np.argmax(np.any(lin != img[xy[:, 1], xy[:, 0]], axis=1)))  # first curve index that differs from the original
    assert 0 < reached < n and not lin[reached:].any()                         # ... from there on everything is zero
    # hufman: the same cut is a failure
    h = codecs.Codec.from_str(ctx, "hufman")
    hd = h.encode(img)
    assert h.decode(hd[:len(hd) - 40]) is None and O.decode_hufman(hd[:len(hd) - 40]) is None
    # hilbert-rle: whole records missing -> zeros behind the last run; half a record -> failure
    r = codecs.Codec.from_str(ctx, "hilbert(rle)")
    flat = (img // 128) * 128
    rd = r.encode(flat)
    whole = rd[:8 + 12 * ((len(rd) - 8) // 12 // 2)]
    g, o = r.decode(whole), O.decode_hilbert_rle(whole)
    assert g is not None and np.array_equal(g, o) and not g[xy[-1, 1], xy[-1, 0]].any()
    assert r.decode(whole + rd[len(whole):len(whole) + 5]) is None and O.decode_hilbert_rle(whole + rd[len(whole):len(whole) + 5]) is None
    zero_count = rd[:8] + bytes([0]) + rd[9:]
    assert r.decode(zero_count) is None and O.decode_hilbert_rle(zero_count) is None  # assert!(self.count > 0)


def test_kmeans_xyrgb_batch_equals_separate_runs(ctx):
    imgs = [cb.synth_image_host(w, h, 40 + i, 5) for i, (w, h) in enumerate([(96, 40), (64, 64), (130, 17), (24, 12)])]
    for max_iters in (3, 0):
        for im, g in zip(imgs, ctx.kmeans_xyrgb_batch(imgs, 10, max_iters=max_iters)):
            same_kmeans(g, O.kmeans_xyrgb(im, 10, mode=O.MODE_EXACT, max_iters=max_iters))


@pytest.mark.parametrize("kind,w,h,k", [("xyrgb", 200, 150, 64), ("rgb", 160, 120, 16)])
def test_kmeans_cluster_one_call_equals_the_session_calls(ctx, kind, w, h, k):
    """cniic_kmeans_cluster = open + reset + run + get + close in one FFI crossing: same answers as the oracle / the session API."""
    img = cb.synth_image_host(w, h, 17, 12)
    kid = cb.POINTS_XYRGB if kind == "xyrgb" else cb.POINTS_RGB
    o = O.kmeans_xyrgb(img, k, max_iters=4) if kind == "xyrgb" else O.kmeans_rgb(img, k, max_iters=4)
    cen, wts, asg, st = cb.kmeans_cluster(ctx, kid, k, img, w * h, max_iters=4, w=w, h_local=h, want_assign=True)
    assert st.iterations == o.iterations and np.array_equal(cen, o.centroids) and np.array_equal(asg, o.assign)
    assert np.array_equal(wts, np.bincount(o.assign, minlength=k).astype(np.uint64))
    none_cen, none_w, none_a, st2 = cb.kmeans_cluster(ctx, kid, k, img, w * h, max_iters=4, w=w, h_local=h, want_centroids=False)
    assert none_cen is None and none_a is None and st2.moved_total == st.moved_total


def test_histogram_bins_belong_to_the_device_not_the_context(ctx):
    """the 2^24-colour and 511^3-delta key spaces are allocated once per device and lent to whichever context
    is counting (bench.rs:27 runs one codec call per rayon worker, so a process holds one context per worker).  Contexts used
    in turn and -- on a real GPU -- from concurrent threads must all get the oracle's histograms, and a second context must
    not cost a second 534 MB key space."""
    import threading
    emulated = "tests/emu/_build" in str(ctx._lib._name).replace("\\", "/")
    rng = np.random.default_rng(77)
    imgs = [rng.integers(0, 256, (32, 32, 3), dtype=np.uint8) // d * d for d in (1, 16, 64)]
    want = [(O.count_freqs_rgb(im.reshape(-1, 3)), O.hist_delta(O.delta(im))) for im in imgs]

    def check(c, i):
        k, n = c.hist_rgb(imgs[i].reshape(-1, 3))
        assert np.array_equal(k, want[i][0][0]) and np.array_equal(n, want[i][0][1])
        k, n = c.hist_delta(imgs[i])
        assert np.array_equal(k, want[i][1][0]) and np.array_equal(n, want[i][1][1])
        _, cen, _ = c.cluster_colors(imgs[i], 4, max_iters=2)  # the Morton-binned pass borrows the colour bins too
        assert np.array_equal(cen, O.cluster_colors(imgs[i], 4, max_iters=2)[1])

    check(ctx, 0)  # both key spaces exist on the device from here on
    free_before = None
    if not emulated:
        import torch
        torch.cuda.synchronize()
        free_before = torch.cuda.mem_get_info()[0]
    others = [cb.Context() for _ in range(3)]
    try:
        for rep in range(2):
            for j, c in enumerate(others + [ctx]):
                check(c, (j + rep) % 3)
        if free_before is not None:
            import torch
            assert free_before - torch.cuda.mem_get_info()[0] < (200 << 20)  # three more contexts, no second copy of the key spaces
            errors = []

            def worker(c, j):
                try:
                    for rep in range(6):
                        check(c, (j + rep) % 3)
                except BaseException as e:  # noqa: BLE001 -- reported on the main thread
                    errors.append(e)
            ts = [threading.Thread(target=worker, args=(c, j)) for j, c in enumerate(others)]
            for t in ts:
                t.start()
            for t in ts:
                t.join()
            assert not errors, errors
    finally:
        for c in others:
            c.close()
    check(ctx, 1)  # the bins outlive the contexts that came and went
