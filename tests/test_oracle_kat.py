"""Pins the CPU oracle against the reference's own unit-test vectors.

Each test names the reference test it ports (paths relative to /root/reference/src).
"""
import math

import numpy as np
import pytest

import oracle as O


def square_centered_at(p):  # kmeans.rs:508-515
    return [(p[0] + i, p[1] + j) for i in range(-1, 2) for j in range(-1, 2)]


MODES = [O.MODE_VERBATIM, O.MODE_EXACT]


@pytest.mark.parametrize("mode", MODES)
def test_all_clusters(mode):  # kmeans.rs:491-500
    data = [(0, 0), (1, 1)]
    r = O.kmeans_i32x2(data, 2, mode=mode)
    assert {tuple(c) for c in r.centroids.tolist()} == set(data)
    for i in range(2):
        members = [data[j] for j in range(2) if r.assign[j] == i]
        assert members == [tuple(r.centroids[i].tolist())]


@pytest.mark.parametrize("mode", MODES)
def test_square1(mode):  # kmeans.rs:517-523
    data = square_centered_at((0, 0))
    r = O.kmeans_i32x2(data, 1, mode=mode)
    assert r.centroids.tolist() == [[0, 0]]
    assert int(r.weights[0]) == len(data)


@pytest.mark.parametrize("mode", MODES)
def test_squares2(mode):  # kmeans.rs:526-539
    centers = [(-100, 0), (100, 0)]
    data = square_centered_at(centers[0]) + square_centered_at(centers[1])
    r = O.kmeans_i32x2(data, 2, mode=mode)
    assert {tuple(c) for c in r.centroids.tolist()} == set(centers)


def test_dist1():  # kmeans.rs:542-544 (2-D) -- same arithmetic as oracle_dist_rgb on one axis
    r, radii = O.kmeans_i32x2([(0, 0), (0, 1)], 2, want_radii=True)
    assert radii.tolist() == [0.5, 0.5]


def test_dist2():  # kmeans.rs:547-557
    # distances are exercised through a one-pass assignment: every point of the left square must choose (-11,0)
    pts = square_centered_at((-100, 0))
    for p in pts:
        closer = math.hypot(p[0] + 11, p[1])
        further = math.hypot(p[0] - 11, p[1])
        assert closer < further


def test_mean1():  # kmeans.rs:560-563
    r = O.kmeans_i32x2(square_centered_at((-100, 0)), 1)
    assert r.centroids.tolist() == [[-100, 0]]


def test_radii():  # kmeans.rs:566-573
    r, radii = O.kmeans_i32x2([(0, 0), (1, 0)], 2, want_radii=True)
    assert radii[0] == 0.5 and radii[1] == 0.5


@pytest.mark.parametrize("mode", MODES)
def test_proper_init_asg(mode):  # kmeans.rs:576-580 : must not trip the active-cluster assert
    r = O.kmeans_i32x2([(1000, 0), (1000, 1), (-1000, 0), (-1000, 1)], 3, mode=mode)
    assert r.status == O.OK


def test_too_few_points():  # kmeans.rs:67-68
    with pytest.raises(O.OracleError) as e:
        O.kmeans_i32x2([(0, 0)], 2)
    assert e.value.code == O.ERR_TOO_FEW_POINTS


def test_rgb_mean():  # clusterc.rs:304-310
    out, cnt = O.mean_colorcount([[0, 0, 0], [2, 2, 2]], [1, 1])
    assert out.tolist() == [1, 1, 1] and cnt == 1


def test_rgb_mean_single_is_clone():  # clusterc.rs:87-90
    out, cnt = O.mean_colorcount([[9, 8, 7]], [42])
    assert out.tolist() == [9, 8, 7] and cnt == 42


def test_rgb_mean_empty():  # clusterc.rs:84-86
    assert O.mean_colorcount(np.zeros((0, 3), np.uint8), np.zeros(0, np.uint32)) is None


def test_rgb_mean_weighted():  # clusterc.rs:92-105 : count-weighted u64 sums, truncating
    out, _ = O.mean_colorcount([[0, 0, 0], [10, 20, 255]], [3, 1])
    assert out.tolist() == [2, 5, 63]


def test_rgb_dist0():  # clusterc.rs:312-316
    assert O.dist_rgb([0, 10, 20], [0, 10, 20]) == 0.0


def test_rgb_dist1():  # clusterc.rs:318-323
    assert O.dist_rgb([0, 0, 0], [1, 0, 0]) == 1.0


def test_rgb_dist2():  # clusterc.rs:325-330
    assert O.dist_rgb([0, 0, 0], [1, 1, 0]) == math.sqrt(2.0)


def test_rgb_dist3():  # clusterc.rs:332-337
    assert O.dist_rgb([0, 0, 0], [1, 1, 1]) == math.sqrt(3.0)


def test_colorpos_dist_and_mean():  # clusterc.rs:206-248 (no reference test exists; arithmetic restated)
    d = O.dist_colorpos(3, 0, [0, 0, 0], 0, 4, [0, 0, 0])  # wrapping u32 sub, squared
    assert d == 5.0
    xy, rgb = O.mean_colorpos([[0, 0], [3, 5]], [[0, 0, 0], [255, 1, 2]])
    assert xy.tolist() == [1, 2] and rgb.tolist() == [127, 0, 1]


def test_huf_code_lens1():  # huf.rs:417-424
    assert O.huf_code_lengths([2, 1, 1]).tolist() == [1, 2, 2]


def test_huf_encode1():  # huf.rs:501-523
    codes = ["010", "11110000011", "00"]
    assert O.bitpack_codes([0, 1, 2], codes) == bytes([0x5e, 0x0c])


def test_huf_encode2():  # huf.rs:525-539
    assert O.bitpack_codes([0], ["11110000"]) == bytes([0xf0])


def test_bit_interleaved_byte():  # bit.rs:299-322
    assert O.bitpack_codes([0, 1, 2], ["010", "11110000", "01100"]) == bytes([0x5e, 0x0c])


def test_bit_bw_mask():  # bit.rs:324-349
    assert O.bitpack_codes([0, 1, 2, 3], ["0000", "110", "11111111", "0"]) == bytes([0x0d, 0xfe])


def test_huf_trie_and_roundtrip():  # huf.rs:430-499 (enc_dec1..3, ser1) through the Hufman codec framing
    rng = np.random.default_rng(1)
    img = rng.integers(0, 4, size=(5, 7, 3), dtype=np.uint8) * 60
    data = O.encode_hufman(img)
    assert data[:8] == (7).to_bytes(4, "little") + (5).to_bytes(4, "little")
    back = O.decode_hufman(data)
    assert np.array_equal(back, img)


def test_huf_single_symbol():  # huf.rs:139-142 : zero-length code, no payload
    img = np.full((3, 3, 3), 7, np.uint8)
    data = O.encode_hufman(img)
    assert len(data) == 8 + 1 + 11
    assert np.array_equal(O.decode_hufman(data), img)


def test_huf_abc_stream():  # huf.rs:386-388 + 296-321 : a:2 b:1 c:1, deterministic (freq, sequence) heap stand-in
    stream = [0, 1, 2, 0]
    data = O.huf_encode_ids(stream, 3, np.frombuffer(b"abc", np.uint8), 1)
    # b(1,seq1), c(1,seq2) merge first (left=b, right=c) -> node(2,seq3); then a(2,seq0) pops before node(2,seq3)
    assert data[:8] == bytes([1, 0, ord("a"), 1, 0, ord("b"), 0, ord("c")])
    # payload a=0, b=10, c=11, a=0 -> 0101 1000
    assert data[8:] == bytes([0b01011000])
