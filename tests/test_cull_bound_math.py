"""The exact-culling argument of DESIGN.md section 4, checked numerically on the CPU (numpy restatement of the bounds the
kernels compute): for any box of points, a centroid with LB_c > U = min_c UB_c is never a minimiser -- not even a tied one --
for any point of the box, in 3-D colour space and in 5-D (position rectangle x colour box)."""
import numpy as np
import pytest


def bounds(cen, lo, hi):
    """cen (k, D), box corners lo/hi (D,): per-centroid min and max squared distance to the box (integer arithmetic)."""
    dmin = np.maximum(0, np.maximum(lo[None] - cen, cen - hi[None]))
    dmax = np.maximum(np.abs(cen - lo[None]), np.abs(cen - hi[None]))
    return (dmin.astype(np.int64) ** 2).sum(1), (dmax.astype(np.int64) ** 2).sum(1)


@pytest.mark.parametrize("D,span", [(3, 255), (5, 4000)])
@pytest.mark.parametrize("seed", range(20))
def test_survivors_contain_every_minimiser(D, span, seed):
    rng = np.random.default_rng(1000 * D + seed)
    k = int(rng.integers(2, 300))
    hi_val = [span, span, 255, 255, 255][:D] if D == 5 else [255] * 3
    cen = np.stack([rng.integers(0, hv + 1, k) for hv in hi_val], axis=1).astype(np.int64)
    if seed % 4 == 0:
        cen[k // 2] = cen[0]  # duplicate centroids: exact ties
    lo = np.array([rng.integers(0, hv + 1) for hv in hi_val], np.int64)
    ext = np.array([rng.integers(0, 64) if j < D - 3 or D == 3 else rng.integers(0, 40) for j in range(D)], np.int64)
    hi = np.minimum(lo + ext, np.array(hi_val))
    pts = np.stack([rng.integers(lo[j], hi[j] + 1, 500) for j in range(D)], axis=1).astype(np.int64)
    lb, ub = bounds(cen, lo, hi)
    U = ub.min()
    keep = lb <= U
    d2 = ((pts[:, None, :] - cen[None]) ** 2).sum(-1)
    best = d2.min(1)
    minimisers = d2 == best[:, None]  # includes ties
    assert not (minimisers & ~keep[None]).any()
    # and the lowest-index minimiser among the survivors is the global lowest-index minimiser
    surv = np.nonzero(keep)[0]
    assert np.array_equal(surv[d2[:, surv].argmin(1)], d2.argmin(1))


def test_packed_rgb_score_orders_like_distance_and_breaks_ties_by_lowest_id():
    """km_assign_rgb_cull packs (2*dot - |c|^2) * 4096 + (4095 - id) into an int32; max() must pick the nearest, lowest id."""
    rng = np.random.default_rng(3)
    cen = rng.integers(0, 256, size=(4096, 3)).astype(np.int64)
    cen[100] = cen[7]
    cen[4095] = [255, 255, 255]
    cen[0] = [0, 0, 0]
    pts = rng.integers(0, 256, size=(300, 3)).astype(np.int64)
    pts[0] = cen[7]
    pts[1] = [255, 255, 255]
    pts[2] = [0, 0, 0]
    dot = pts @ cen.T
    packed = dot * 8192 + (-(cen ** 2).sum(1) * 4096 + 4095 - np.arange(4096))[None]
    assert packed.max() < 2 ** 31 and packed.min() >= -2 ** 31 and (dot * 8192).max() < 2 ** 31
    win = packed.argmax(1)
    d2 = ((pts[:, None, :] - cen[None]) ** 2).sum(-1)
    assert np.array_equal(win, d2.argmin(1))  # argmin returns the first (lowest id) minimiser
    assert np.array_equal(4095 - (packed.max(1) & 4095), win)
    assert np.array_equal(packed.max(1) >> 12, (2 * dot - (cen ** 2).sum(1)[None])[np.arange(len(pts)), win])


def test_xyrgb_key_fits_int32_at_the_size_limit():
    """key = 2*(p.c) - |c|^2 with coordinates up to CNIIC_MAX_DIM - 1 = 16383 must stay inside int32."""
    m = 16383
    worst_pos = 2 * (m * m + m * m + 3 * 255 * 255)
    worst_neg = -(m * m + m * m + 3 * 255 * 255)
    assert worst_pos < 2 ** 31 and worst_neg > -2 ** 31
    # tile-relative part + folded constant, as the culled kernel computes it
    rel = 2 * (63 * m + 31 * m + 3 * 255 * 255)
    const = 2 * (m * m + m * m)
    assert rel + const < 2 ** 31
