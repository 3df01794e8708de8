"""ctypes binding of the CPU oracle (oracle/cniic_oracle.c).

TEST INFRASTRUCTURE ONLY.  May be imported by tests/, __graft_entry__.smoke() and the cpu_baseline /
``--impl reference`` legs of bench.py -- never by the product package ``cniic_b200``.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")

MODE_EXACT, MODE_VERBATIM = 0, 1
TIE_KEEP_CURRENT, TIE_LOWEST_INDEX = 0, 1
OK, ERR_BAD_ARG, ERR_TOO_FEW_POINTS, ERR_TOO_FEW_ACTIVE = 0, 1, 2, 3


class OracleError(RuntimeError):
    def __init__(self, code):
        super().__init__(f"oracle status {code}")
        self.code = code


class _Stats(C.Structure):
    _fields_ = [("iterations", C.c_uint32), ("empty_events", C.c_uint32), ("moved_last", C.c_uint64),
                ("dist_evals", C.c_uint64), ("moved_total", C.c_uint64)]


@dataclass
class KMeansResult:
    centroids: np.ndarray  # (k, D) int64
    weights: np.ndarray    # (k,) uint64
    assign: np.ndarray     # (n,) uint32
    iterations: int
    empty_events: int
    moved_last: int
    dist_evals: int
    moved_total: int
    status: int = 0


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "cniic_oracle.c")
    hdr = os.path.join(_HERE, "cniic_oracle.h")
    if force or not os.path.exists(_SO) or (
            os.path.exists(src) and os.path.getmtime(_SO) < max(os.path.getmtime(src), os.path.getmtime(hdr))):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.oracle_dist_rgb.restype = C.c_double
        _lib.oracle_dist_colorpos.restype = C.c_double
        _lib.oracle_mse.restype = C.c_double
        _lib.oracle_sse.restype = C.c_uint64
        for name in ("oracle_count_freqs_rgb", "oracle_hist_delta", "oracle_rle_exact", "oracle_huf_encode_ids",
                     "oracle_bitpack_codes", "oracle_encode_hufman", "oracle_encode_delta", "oracle_encode_voronoi",
                     "oracle_encode_cluster_colors", "oracle_encode_hilbert_rle"):
            getattr(_lib, name).restype = C.c_size_t
    return _lib


def _p(a, t=C.c_void_p):
    return None if a is None else a.ctypes.data_as(t)


def _u8(a):
    return np.ascontiguousarray(a, dtype=np.uint8)


def _stats(st, cen, wts, asg, rc):
    return KMeansResult(cen, wts, asg, st.iterations, st.empty_events, st.moved_last, st.dist_evals, st.moved_total, rc)


def _check(rc, allow=()):
    if rc != OK and rc not in allow:
        raise OracleError(rc)


def kmeans_i32x2(pts, k, mode=MODE_VERBATIM, tie=TIE_KEEP_CURRENT, max_iters=0, want_radii=False):
    pts = np.ascontiguousarray(pts, dtype=np.int32).reshape(-1, 2)
    n = len(pts)
    cen = np.zeros((k, 2), np.int32)
    asg = np.zeros(n, np.uint32)
    radii = np.zeros(k, np.float64) if want_radii else None
    st = _Stats()
    rc = lib().oracle_kmeans_i32x2(_p(pts), C.c_size_t(n), C.c_size_t(k), mode, tie, C.c_uint32(max_iters), _p(cen),
                                   _p(asg), _p(radii), C.byref(st))
    _check(rc)
    res = _stats(st, cen.astype(np.int64), np.bincount(asg, minlength=k).astype(np.uint64), asg, rc)
    return (res, radii) if want_radii else res


def kmeans_rgb(rgb, k, counts=None, mode=MODE_EXACT, tie=TIE_KEEP_CURRENT, max_iters=0, allow_inactive=False):
    rgb = _u8(rgb).reshape(-1, 3)
    n = len(rgb)
    cnt = None if counts is None else np.ascontiguousarray(counts, dtype=np.uint32)
    cen = np.zeros((k, 3), np.uint8)
    wts = np.zeros(k, np.uint64)
    asg = np.zeros(n, np.uint32)
    st = _Stats()
    rc = lib().oracle_kmeans_rgb(_p(rgb), _p(cnt), C.c_size_t(n), C.c_size_t(k), mode, tie, C.c_uint32(max_iters),
                                 _p(cen), _p(wts), _p(asg), C.byref(st))
    _check(rc, (ERR_TOO_FEW_ACTIVE,) if allow_inactive else ())
    return _stats(st, cen.astype(np.int64), wts, asg, rc)


def kmeans_xyrgb(img, k, mode=MODE_EXACT, tie=TIE_KEEP_CURRENT, max_iters=0, allow_inactive=False):
    img = _u8(img)
    h, w = img.shape[:2]
    cxy = np.zeros((k, 2), np.uint32)
    crgb = np.zeros((k, 3), np.uint8)
    wts = np.zeros(k, np.uint64)
    asg = np.zeros(h * w, np.uint32)
    st = _Stats()
    rc = lib().oracle_kmeans_xyrgb(_p(img), C.c_uint32(w), C.c_uint32(h), C.c_size_t(k), mode, tie,
                                   C.c_uint32(max_iters), _p(cxy), _p(crgb), _p(wts), _p(asg), C.byref(st))
    _check(rc, (ERR_TOO_FEW_ACTIVE,) if allow_inactive else ())
    cen = np.concatenate([cxy.astype(np.int64), crgb.astype(np.int64)], axis=1)
    return _stats(st, cen, wts, asg, rc)


def dist_rgb(a, b):
    a, b = _u8(a), _u8(b)
    return lib().oracle_dist_rgb(_p(a), _p(b))


def dist_colorpos(ax, ay, a, bx, by, b):
    a, b = _u8(a), _u8(b)
    return lib().oracle_dist_colorpos(C.c_uint32(ax), C.c_uint32(ay), _p(a), C.c_uint32(bx), C.c_uint32(by), _p(b))


def mean_colorcount(rgb, counts=None):
    rgb = _u8(rgb).reshape(-1, 3)
    cnt = None if counts is None else np.ascontiguousarray(counts, dtype=np.uint32)
    out = np.zeros(3, np.uint8)
    oc = C.c_uint32(0)
    ok = lib().oracle_mean_colorcount(_p(rgb), _p(cnt), C.c_size_t(len(rgb)), _p(out), C.byref(oc))
    return (out, oc.value) if ok else None


def mean_colorpos(xy, rgb):
    xy = np.ascontiguousarray(xy, dtype=np.uint32).reshape(-1, 2)
    rgb = _u8(rgb).reshape(-1, 3)
    oxy = np.zeros(2, np.uint32)
    orgb = np.zeros(3, np.uint8)
    ok = lib().oracle_mean_colorpos(_p(xy), _p(rgb), C.c_size_t(len(xy)), _p(oxy), _p(orgb))
    return (oxy, orgb) if ok else None


def count_freqs_rgb(rgb):
    rgb = _u8(rgb).reshape(-1, 3)
    n = len(rgb)
    cap = max(1, min(n, 1 << 24))
    keys = np.zeros(cap, np.uint32)
    cnts = np.zeros(cap, np.uint64)
    u = lib().oracle_count_freqs_rgb(_p(rgb), C.c_size_t(n), _p(keys), _p(cnts))
    return keys[:u].copy(), cnts[:u].copy()


def cluster_colors(img, k, mode=MODE_EXACT, tie=TIE_KEEP_CURRENT, max_iters=0):
    img = _u8(img)
    h, w = img.shape[:2]
    out = np.zeros_like(img)
    cen = np.zeros((k, 3), np.uint8)
    st = _Stats()
    rc = lib().oracle_cluster_colors(_p(img), C.c_uint32(w), C.c_uint32(h), C.c_size_t(k), mode, tie,
                                     C.c_uint32(max_iters), _p(out), _p(cen), C.byref(st))
    _check(rc)
    return out, cen, st.iterations


def voronoi_fill(cxy, crgb, w, h):
    cxy = np.ascontiguousarray(cxy, dtype=np.uint32).reshape(-1, 2)
    crgb = _u8(crgb).reshape(-1, 3)
    out = np.zeros((h, w, 3), np.uint8)
    lib().oracle_voronoi_fill(_p(cxy), _p(crgb), C.c_size_t(len(cxy)), C.c_uint32(w), C.c_uint32(h), _p(out))
    return out


def mse(a, b):
    a, b = _u8(a), _u8(b)
    h, w = a.shape[:2]
    return lib().oracle_mse(_p(a), _p(b), C.c_uint32(w), C.c_uint32(h))


def sse(a, b):
    a, b = _u8(a), _u8(b)
    return int(lib().oracle_sse(_p(a), _p(b), C.c_size_t(a.size // 3)))


def hilbert_xy(w, h):
    out = np.zeros((w * h, 2), np.uint32)
    lib().oracle_hilbert_xy(C.c_uint32(w), C.c_uint32(h), _p(out))
    return out


def hilbert_gather(img):
    img = _u8(img)
    h, w = img.shape[:2]
    out = np.zeros((w * h, 3), np.uint8)
    lib().oracle_hilbert_gather(_p(img), C.c_uint32(w), C.c_uint32(h), _p(out))
    return out


def delta(img):
    img = _u8(img)
    h, w = img.shape[:2]
    out = np.zeros((w * h, 3), np.int16)
    lib().oracle_delta(_p(img), C.c_uint32(w), C.c_uint32(h), _p(out))
    return out


def undelta(diff, w, h):
    diff = np.ascontiguousarray(diff, dtype=np.int16)
    out = np.zeros((h, w, 3), np.uint8)
    if lib().oracle_undelta(_p(diff), C.c_uint32(w), C.c_uint32(h), _p(out)):
        return None  # FromDiff's try_into().unwrap() panics (hilbertc.rs:503-506)
    return out


def hist_delta(diff):
    diff = np.ascontiguousarray(diff, dtype=np.int16).reshape(-1, 3)
    n = len(diff)
    keys = np.zeros(max(n, 1), np.uint32)
    cnts = np.zeros(max(n, 1), np.uint64)
    u = lib().oracle_hist_delta(_p(diff), C.c_size_t(n), _p(keys), _p(cnts), C.c_size_t(n))
    return keys[:u].copy(), cnts[:u].copy()


def rle_exact(stream):
    s = _u8(stream).reshape(-1, 3)
    n = len(s)
    cnt = np.zeros(max(n, 1), np.uint8)
    col = np.zeros((max(n, 1), 3), np.uint8)
    r = lib().oracle_rle_exact(_p(s), C.c_size_t(n), _p(cnt), _p(col))
    return cnt[:r].copy(), col[:r].copy()


def huf_code_lengths(freqs):
    f = np.ascontiguousarray(freqs, dtype=np.uint64)
    out = np.zeros(len(f), np.uint32)
    lib().oracle_huf_code_lengths(_p(f), C.c_size_t(len(f)), _p(out))
    return out


def huf_encode_ids(stream, nsym, sym_bytes, sym_size):
    s = np.ascontiguousarray(stream, dtype=np.uint32)
    sb = _u8(sym_bytes)
    need = lib().oracle_huf_encode_ids(_p(s), C.c_size_t(len(s)), C.c_size_t(nsym), _p(sb), C.c_size_t(sym_size), None,
                                       C.c_size_t(0))
    out = np.zeros(max(need, 1), np.uint8)
    lib().oracle_huf_encode_ids(_p(s), C.c_size_t(len(s)), C.c_size_t(nsym), _p(sb), C.c_size_t(sym_size), _p(out),
                                C.c_size_t(need))
    return out[:need].tobytes()


def bitpack_codes(stream, codes):
    s = np.ascontiguousarray(stream, dtype=np.uint32)
    arr = (C.c_char_p * len(codes))(*[c.encode() for c in codes])
    out = np.zeros(max(1, sum(len(c) for c in codes) * len(s)), np.uint8)
    n = lib().oracle_bitpack_codes(_p(s), C.c_size_t(len(s)), arr, _p(out), C.c_size_t(len(out)))
    return out[:n].tobytes()


def _encode(fn, img, *args):
    img = _u8(img)
    h, w = img.shape[:2]
    need = fn(_p(img), C.c_uint32(w), C.c_uint32(h), *args, None, C.c_size_t(0))
    out = np.zeros(max(need, 1), np.uint8)
    fn(_p(img), C.c_uint32(w), C.c_uint32(h), *args, _p(out), C.c_size_t(need))
    return out[:need].tobytes()


def _decode(fn, data, max_px=1 << 26):
    buf = np.frombuffer(data, np.uint8)
    w, h = C.c_uint32(0), C.c_uint32(0)
    # dims are the first 8 bytes (ser.rs:146-151)
    if len(buf) < 8:
        return None
    ww, hh = int(np.frombuffer(data[:4], "<u4")[0]), int(np.frombuffer(data[4:8], "<u4")[0])
    if ww * hh > max_px:
        return None
    out = np.zeros((hh, ww, 3), np.uint8)
    rc = fn(_p(buf), C.c_size_t(len(buf)), C.byref(w), C.byref(h), _p(out), C.c_size_t(ww * hh))
    return out if rc == 0 else None


def encode_hufman(img):
    return _encode(lib().oracle_encode_hufman, img)


def decode_hufman(data):
    return _decode(lib().oracle_decode_hufman, data)


def encode_delta(img):
    return _encode(lib().oracle_encode_delta, img)


def decode_delta(data):
    return _decode(lib().oracle_decode_delta, data)


def encode_voronoi(img, k, mode=MODE_EXACT, tie=TIE_KEEP_CURRENT, max_iters=0):
    return _encode(lib().oracle_encode_voronoi, img, C.c_size_t(k), mode, tie, C.c_uint32(max_iters))


def decode_voronoi(data):
    return _decode(lib().oracle_decode_voronoi, data)


def encode_cluster_colors(img, k, mode=MODE_EXACT, tie=TIE_KEEP_CURRENT, max_iters=0):
    return _encode(lib().oracle_encode_cluster_colors, img, C.c_size_t(k), mode, tie, C.c_uint32(max_iters))


def encode_hilbert_rle(img):
    return _encode(lib().oracle_encode_hilbert_rle, img)


def decode_hilbert_rle(data):
    return _decode(lib().oracle_decode_hilbert_rle, data)
