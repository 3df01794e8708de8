"""Generates tests/golden/golden_v1.npz from the CPU oracle (oracle/cniic_oracle.c).

The reference (Rust) cannot be built or run in this environment, so these are NOT outputs of the reference binary: they
freeze the oracle's answers on small seeded inputs so that later edits of the oracle (or of the deterministic stand-in rules)
cannot drift silently.  The reference's own known-answer vectors are ported separately in tests/test_oracle_kat.py.

    python tests/golden/make_golden.py        # rewrites golden_v1.npz
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle as O  # noqa: E402


def image(seed, w, h):
    """Small seeded test image that does not depend on the product library."""
    rng = np.random.default_rng(seed)
    base = rng.integers(0, 256, size=((h + 7) // 8, (w + 7) // 8, 3), dtype=np.uint8)
    img = np.repeat(np.repeat(base, 8, axis=0), 8, axis=1)[:h, :w].copy()
    img ^= rng.integers(0, 8, size=img.shape, dtype=np.uint8)
    return img


def build():
    out = {}
    for (w, h) in [(4, 4), (8, 8), (5, 3), (3, 5), (13, 7), (16, 16)]:
        out[f"hilbert_{w}x{h}"] = O.hilbert_xy(w, h)
    img = image(1, 24, 18)
    out["img"] = img
    out["delta"] = O.delta(img)
    k, c = O.hist_delta(out["delta"])
    out["hist_delta_keys"], out["hist_delta_counts"] = k, c
    k, c = O.count_freqs_rgb(img)
    out["hist_rgb_keys"], out["hist_rgb_counts"] = k, c
    for name, data in [("hufman", O.encode_hufman(img)), ("delta_stream", O.encode_delta(img)), ("rle", O.encode_hilbert_rle((img // 64) * 64)),
                       ("voronoi6", O.encode_voronoi(img, 6)), ("ccol5", O.encode_cluster_colors(img, 5))]:
        out[f"stream_{name}"] = np.frombuffer(data, np.uint8)
    for tie in (O.TIE_KEEP_CURRENT, O.TIE_LOWEST_INDEX):
        r = O.kmeans_xyrgb(img, 7, mode=O.MODE_EXACT, tie=tie)
        out[f"km5_t{tie}_cen"], out[f"km5_t{tie}_asg"], out[f"km5_t{tie}_it"] = r.centroids, r.assign, np.array([r.iterations, r.empty_events])
        r = O.kmeans_rgb(img, 9, mode=O.MODE_EXACT, tie=tie)
        out[f"km3_t{tie}_cen"], out[f"km3_t{tie}_asg"], out[f"km3_t{tie}_it"] = r.centroids, r.assign, np.array([r.iterations, r.empty_events])
    r = O.kmeans_xyrgb(img, 7, mode=O.MODE_VERBATIM)
    out["km5_verbatim_cen"], out["km5_verbatim_it"] = r.centroids, np.array([r.iterations, r.dist_evals])
    cxy = np.array([[2, 3], [20, 4], [11, 15], [11, 15]], np.uint32)
    crgb = np.array([[1, 2, 3], [40, 50, 60], [200, 100, 0], [9, 9, 9]], np.uint8)
    out["fill"] = O.voronoi_fill(cxy, crgb, 24, 18)
    return out


if __name__ == "__main__":
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_v1.npz"), **build())
    print("written")
