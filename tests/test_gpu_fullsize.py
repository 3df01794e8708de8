"""GPU tests at BASELINE.json's full sizes, through size-independent properties (the oracle would take hours)."""
import numpy as np
import pytest

import cniic_b200 as cb

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = cb.Context()
    yield c
    c.close()


def check_kmeans_properties(pts, cen, wts, asg, k, rng, D):
    n = len(asg)
    assert int(wts.sum()) == n
    cnt = np.bincount(asg, minlength=k)
    assert np.array_equal(cnt.astype(np.uint64), wts)
    # every non-empty centroid is the truncated integer mean of its members (clusterc.rs:81-114 / 215-248)
    for j in range(D):
        sums = np.bincount(asg, weights=pts[:, j].astype(np.float64), minlength=k)  # exact: sums < 2^53
        nz = cnt > 0
        assert np.array_equal((sums[nz].astype(np.int64) // cnt[nz]), cen[nz, j])


def test_c2_cluster_colors_k256_4096(ctx):
    w = h = 4096
    k = 256
    img = cb.synth_image_host(w, 8, 0xC0FFEE + 2, 192, y0=0, h_total=h)  # only to cross-check the device generator
    d = ctx.device_alloc(w * h * 3)
    cb.synth_image_device(ctx, d, w, h, 0xC0FFEE + 2, 192)
    s = cb.KMeansSession(ctx, cb.POINTS_RGB, k, d, w * h, on_device=True)
    s.reset()
    st = s.run(3)
    cen, wts, asg = s.get()
    host = np.zeros((h, w, 3), np.uint8)
    ctx.d2h(host, d)
    assert np.array_equal(host[:8], img)  # device generator == host generator
    pts = host.reshape(-1, 3)
    assert st.iterations == 3
    check_kmeans_properties(pts, cen, wts, asg, k, None, 3)
    # idempotence: one more pass from these centroids must reproduce itself exactly (same assign kernel, same sums)
    s.reset()
    st2 = s.run(3)
    cen2, wts2, asg2 = s.get()
    assert np.array_equal(cen, cen2) and np.array_equal(asg, asg2)
    s.close()
    ctx.device_free(d)


def _session_result(ctx, kind, k, d, n, iters, flags=0, **kw):
    s = cb.KMeansSession(ctx, kind, k, d, n, on_device=True, flags=flags, **kw)
    s.reset()
    st = s.run(iters)
    cen, wts, asg = s.get()
    s.close()
    return cen, wts, asg, st


def test_c2_culled_kernels_equal_brute_force_at_full_size(ctx):
    """at the headline size the culled assignment must be THE nearest centroid, not merely self-consistent.  The
    default (colour-sorted, exactly culled, second kernel version) and the brute-force kernel (every pixel scores all 256 centroids)
    must agree bit for bit after 3 iterations -- centroids, weights, every one of the 16.7 M assignments, moved counts -- and a
    numpy int64 argmin over a sample of pixels must agree with both."""
    w = h = 4096
    k = 256
    d = ctx.device_alloc(w * h * 3)
    cb.synth_image_device(ctx, d, w, h, 0xC0FFEE + 2, 192)
    cen, wts, asg, st = _session_result(ctx, cb.POINTS_RGB, k, d, w * h, 3)
    cen_b, wts_b, asg_b, st_b = _session_result(ctx, cb.POINTS_RGB, k, d, w * h, 3, flags=cb._lib.KMEANS_NO_CULL)
    assert np.array_equal(cen, cen_b) and np.array_equal(wts, wts_b) and np.array_equal(asg, asg_b)
    assert (st.moved_last, st.moved_total, st.iterations) == (st_b.moved_last, st_b.moved_total, st_b.iterations)
    assert st.pairs_scored < st_b.pairs_scored // 10  # the culled run really culled
    # nearest-centroid sample against numpy: centroids after 2 iterations, assignment of the 3rd pass
    cen2, _, asg2, _ = _session_result(ctx, cb.POINTS_RGB, k, d, w * h, 2)
    host = np.zeros((h, w, 3), np.uint8)
    ctx.d2h(host, d)
    pts = host.reshape(-1, 3)
    rng = np.random.default_rng(0)
    idx = rng.integers(0, w * h, 20000)
    d2 = ((pts[idx].astype(np.int64)[:, None, :] - cen2[None].astype(np.int64)) ** 2).sum(-1)
    best = d2.min(1)
    assert np.array_equal(d2[np.arange(len(idx)), asg[idx]], best)  # a minimiser ...
    tied_prev = d2[np.arange(len(idx)), asg2[idx]] == best
    assert np.array_equal(asg[idx][tied_prev], asg2[idx][tied_prev])  # ... the current cluster if it ties (kmeans.rs:350-378) ...
    moved = ~tied_prev
    assert np.array_equal(asg[idx][moved], d2[moved].argmin(1))  # ... else the lowest index
    # the unique-colour path (what cluster-colors runs, clusterc.rs:19-28): same checks on its own point list
    cen_u, nu, st_u = ctx.cluster_colors_device(d, w * h, k, max_iters=3)
    assert nu == len(np.unique(pts.view(np.dtype((np.void, 3))))) and st_u.iterations == 3
    ctx.device_free(d)


def test_c3_culled_kernels_equal_brute_force_at_full_size(ctx):
    """The same for the north-star configuration: voronoi k=2048 on 7680x4320, default (three-level culling inside the assign
    kernel) against the brute-force kernel, 2 iterations (the brute-force pass takes ~10 ms each)."""
    w, h, k = 7680, 4320, 2048
    d = ctx.device_alloc(w * h * 3)
    cb.synth_image_device(ctx, d, w, h, 0xC0FFEE + 3, 2048)
    kw = dict(w=w, h_local=h)
    cen, wts, asg, st = _session_result(ctx, cb.POINTS_XYRGB, k, d, w * h, 2, **kw)
    cen_b, wts_b, asg_b, st_b = _session_result(ctx, cb.POINTS_XYRGB, k, d, w * h, 2, flags=cb._lib.KMEANS_NO_CULL, **kw)
    assert np.array_equal(cen, cen_b) and np.array_equal(wts, wts_b) and np.array_equal(asg, asg_b)
    assert (st.moved_last, st.moved_total) == (st_b.moved_last, st_b.moved_total)
    assert st.pairs_scored < st_b.pairs_scored // 100
    ctx.device_free(d)


def test_c3_voronoi_k2048_8k(ctx):
    w, h, k = 7680, 4320, 2048
    d = ctx.device_alloc(w * h * 3)
    cb.synth_image_device(ctx, d, w, h, 0xC0FFEE + 3, 2048)
    s = cb.KMeansSession(ctx, cb.POINTS_XYRGB, k, d, w * h, w=w, h_local=h, on_device=True)
    s.reset()
    st = s.run(2)
    cen, wts, asg = s.get()
    host = np.zeros((h, w, 3), np.uint8)
    ctx.d2h(host, d)
    n = w * h
    pts = np.empty((n, 5), np.int32)
    pts[:, 0] = np.arange(n, dtype=np.int64) % w
    pts[:, 1] = np.arange(n, dtype=np.int64) // w
    pts[:, 2:] = host.reshape(-1, 3)
    check_kmeans_properties(pts, cen, wts, asg, k, None, 5)
    # sample check of exact nearest-centroid: run ONE iteration, fetch its centroids, run a second, and verify a
    # random sample of second-pass assignments against numpy int64 distances to the first-pass centroids
    s.reset()
    s.run(1)
    cen1, _, asg1 = s.get()
    s.run(1)
    _, _, asg2 = s.get()
    rng = np.random.default_rng(0)
    idx = rng.integers(0, n, 4000)
    d2 = ((pts[idx].astype(np.int64)[:, None, :] - cen1[None].astype(np.int64)) ** 2).sum(-1)
    best = d2.min(1)
    assert np.array_equal(d2[np.arange(len(idx)), asg2[idx]], best)
    # keep-current tie rule: a point whose previous cluster is also a minimiser must not have moved
    tied_prev = d2[np.arange(len(idx)), asg1[idx]] == best
    assert np.array_equal(asg2[idx][tied_prev], asg1[idx][tied_prev])
    # voronoi fill at full size: sample rows against numpy
    cxy = cen[:, :2].astype(np.uint32)
    crgb = cen[:, 2:].astype(np.uint8)
    out = ctx.voronoi_fill(cxy, crgb, w, h)
    ys = rng.integers(0, h, 3)
    for y in ys:
        xs = np.arange(w, dtype=np.int64)
        dd = (cxy[None, :, 0].astype(np.int64) - xs[:, None]) ** 2 + (cxy[None, :, 1].astype(np.int64) - int(y)) ** 2
        assert np.array_equal(out[y], crgb[dd.argmin(1)])
    s.close()
    ctx.device_free(d)


def test_c5_integer_stages_8192(ctx):
    w = h = 8192
    rng = np.random.default_rng(1)
    # smooth-ish random image built cheaply on the host
    base = rng.integers(0, 256, size=(h // 64, w // 64, 3), dtype=np.uint8)
    img = np.repeat(np.repeat(base, 64, axis=0), 64, axis=1)
    img ^= rng.integers(0, 4, size=img.shape, dtype=np.uint8)
    xy = ctx.hilbert_xy(w, h)
    lin = xy[:, 1].astype(np.int64) * w + xy[:, 0]
    assert np.array_equal(np.sort(lin), np.arange(w * h))  # bijection
    assert np.abs(np.diff(xy.astype(np.int64), axis=0)).sum(axis=1).max() == 1  # true Hilbert curve on 2^n squares
    d = ctx.delta(img)
    g = img.reshape(-1, 3)[lin]
    assert np.array_equal(d[0], g[0].astype(np.int16))
    assert np.array_equal(d[1:], g[1:].astype(np.int16) - g[:-1].astype(np.int16))
    assert np.array_equal(ctx.undelta(d, w, h), img)  # encode -> decode round trip (bench.rs:57-59)
    keys, cnts = ctx.hist_delta(img)
    assert int(cnts.sum()) == w * h  # checksum of the histogram
    dk = ((d[:, 0].astype(np.int64) + 255) * 511 + (d[:, 1] + 255)) * 511 + (d[:, 2] + 255)
    uk, uc = np.unique(dk, return_counts=True)
    assert np.array_equal(keys, uk.astype(np.uint32)) and np.array_equal(cnts, uc.astype(np.uint64))
    ck, cc = ctx.hist_rgb(img)
    assert int(cc.sum()) == w * h


@pytest.mark.parametrize("expr,w,h,blobs", [("delta", 8192, 8192, 4096), ("hufman", 4096, 4096, 192), ("hilbert(rle)", 4096, 4096, 192),
                                            ("cluster-colors(256)", 4096, 4096, 192)])
def test_codecs_fullsize_roundtrip(ctx, expr, w, h, blobs):
    """Whole codecs at BASELINE sizes: the parallel Huffman decoder (thousands of 64-Kbit chunks), the run-length coder and
    the fused histogram on tens of millions of symbols.  Lossless codecs must return the image (bench.rs:57-59); for
    cluster-colors the decoded image must be the recoloured one: <= k colours, and decoding is idempotent under re-encoding."""
    from cniic_b200 import codecs
    d = ctx.device_alloc(w * h * 3)
    cb.synth_image_device(ctx, d, w, h, 0xC0FFEE + 7, blobs)
    img = np.zeros((h, w, 3), np.uint8)
    ctx.d2h(img, d)
    ctx.device_free(d)
    if expr.startswith("hilbert"):
        img = (img // 32) * 32  # give the run-length coder runs
    c = codecs.Codec.from_str(ctx, expr, 3)
    data = c.encode(img)
    dec = c.decode(data)
    assert dec is not None and dec.shape == img.shape
    if c.is_lossless():
        assert ctx.sse(img, dec) == 0 and np.array_equal(dec[::97], img[::97])
    else:
        keys, cnts = ctx.hist_rgb(dec)
        assert 1 <= len(keys) <= 256 and int(cnts.sum()) == w * h
        assert c.decode(data[:len(data) - len(data) // 3]) is None  # a truncated payload is rejected, not mis-decoded
