"""GPU results against the committed golden fixtures (tests/golden/golden_v1.npz), through the C ABI."""
import os

import numpy as np
import pytest

import cniic_b200 as cb

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))


def test_gpu_matches_committed_golden_fixtures():
    gold = np.load(os.path.join(HERE, "golden", "golden_v1.npz"))
    ctx = cb.Context()
    img = gold["img"]
    for (w, h) in [(4, 4), (8, 8), (5, 3), (3, 5), (13, 7), (16, 16)]:
        assert np.array_equal(ctx.hilbert_xy(w, h), gold[f"hilbert_{w}x{h}"])
    assert np.array_equal(ctx.delta(img), gold["delta"])
    k, c = ctx.hist_delta(img)
    assert np.array_equal(k, gold["hist_delta_keys"]) and np.array_equal(c, gold["hist_delta_counts"])
    k, c = ctx.hist_rgb(img)
    assert np.array_equal(k, gold["hist_rgb_keys"]) and np.array_equal(c, gold["hist_rgb_counts"])
    assert ctx.codec_encode("hufman", img) == gold["stream_hufman"].tobytes()
    assert ctx.codec_encode("delta", img) == gold["stream_delta_stream"].tobytes()
    assert ctx.codec_encode("hilbert(rle)", (img // 64) * 64) == gold["stream_rle"].tobytes()
    assert ctx.codec_encode("voronoi(6)", img) == gold["stream_voronoi6"].tobytes()
    assert ctx.codec_encode("cluster-colors(5)", img) == gold["stream_ccol5"].tobytes()
    for tie in (cb.TIE_KEEP_CURRENT, cb.TIE_LOWEST_INDEX):
        g = ctx.kmeans_xyrgb(img, 7, tie=tie)
        assert np.array_equal(g.centroids, gold[f"km5_t{tie}_cen"]) and np.array_equal(g.assign, gold[f"km5_t{tie}_asg"])
        assert [g.iterations, g.empty_events] == gold[f"km5_t{tie}_it"].tolist()
        g = ctx.kmeans_rgb(img, 9, tie=tie)
        assert np.array_equal(g.centroids, gold[f"km3_t{tie}_cen"]) and np.array_equal(g.assign, gold[f"km3_t{tie}_asg"])
        assert [g.iterations, g.empty_events] == gold[f"km3_t{tie}_it"].tolist()
    cxy = np.array([[2, 3], [20, 4], [11, 15], [11, 15]], np.uint32)
    crgb = np.array([[1, 2, 3], [40, 50, 60], [200, 100, 0], [9, 9, 9]], np.uint8)
    assert np.array_equal(ctx.voronoi_fill(cxy, crgb, 24, 18), gold["fill"])
    ctx.close()


def test_graft_entry_smoke_and_extras():
    import sys
    sys.path.insert(0, os.path.dirname(HERE))
    import __graft_entry__ as g
    g.smoke()
    g.smoke_extras()
