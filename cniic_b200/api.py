"""Thin Python host layer over the C ABI (numpy in / numpy out).

Mirrors the reference's inner seams for the hot path: ``kmeans::cluster`` (kmeans.rs:21),
``utils::count_freqs`` (utils.rs:4), ``hilbert::iter`` (hilbert.rs:40), the voronoi fill (clusterc.rs:179-186) and the
``Codec`` trait (codec.rs:14-19, see codecs.py).  All compute happens in libcniic_b200.so on the GPU.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib as L


class CniicError(RuntimeError):
    def __init__(self, code: int, msg: str = ""):
        super().__init__(f"cniic_b200 status {code}: {msg}")
        self.code = code


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _u8(a):
    return np.ascontiguousarray(a, dtype=np.uint8)


@dataclass
class KMeansResult:
    centroids: np.ndarray   # (k, D) int64 ; D = 3: r,g,b ; D = 5: x,y,r,g,b
    weights: np.ndarray     # (k,) uint64
    assign: np.ndarray | None  # (n,) uint16
    iterations: int
    empty_events: int
    moved_last: int
    moved_total: int
    converged: bool
    gpu_launches: int
    device_ms: float
    status: int = 0


class Context:
    """One CUDA device + stream (cniic_ctx).  Not thread-safe; one Context per thread."""

    def __init__(self, device: int = -1, rank: int = 0, world: int = 1, nccl_unique_id: bytes | None = None):
        self._lib = L.lib()
        h = C.c_void_p()
        if world > 1:
            if nccl_unique_id is None or len(nccl_unique_id) != 128:
                raise ValueError("a 128-byte ncclUniqueId is required for world > 1")
            buf = (C.c_uint8 * 128).from_buffer_copy(nccl_unique_id)
            rc = self._lib.cniic_ctx_create_dist(device, rank, world, buf, C.byref(h))
        else:
            rc = self._lib.cniic_ctx_create(device, C.byref(h))
        if rc != L.OK:
            raise CniicError(rc, "cniic_ctx_create failed (no CUDA device? there is no CPU fallback)")
        self.h = h
        self.rank, self.world = rank, world

    @staticmethod
    def nccl_unique_id() -> bytes:
        buf = (C.c_uint8 * 128)()
        rc = L.lib().cniic_nccl_unique_id(buf)
        if rc != L.OK:
            raise CniicError(rc, "ncclGetUniqueId failed")
        return bytes(buf)

    def p2p_export(self) -> bytes:
        """CUDA IPC handle (64 bytes) of this rank's exchange region for the peer-memory all-reduce."""
        buf = (C.c_uint8 * 64)()
        self.check(self._lib.cniic_ctx_p2p_export(self.h, buf))
        return bytes(buf)

    def p2p_connect(self, handles: bytes):
        """handles = world x 64 bytes in rank order (every rank's p2p_export())."""
        buf = (C.c_uint8 * len(handles)).from_buffer_copy(handles)
        self.check(self._lib.cniic_ctx_p2p_connect(self.h, buf))

    def close(self):
        if getattr(self, "h", None):
            self._lib.cniic_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # ---- helpers ----
    def check(self, rc: int, allow=()):
        if rc != L.OK and rc not in allow:
            raise CniicError(rc, (self._lib.cniic_last_error(self.h) or b"").decode(errors="replace"))
        return rc

    @property
    def stream(self) -> int:
        return self._lib.cniic_ctx_stream(self.h) or 0

    @property
    def launches(self) -> int:
        return int(self._lib.cniic_ctx_launches(self.h))

    def sync(self):
        self.check(self._lib.cniic_ctx_sync(self.h))

    def set_max_iters(self, n: int):
        self.check(self._lib.cniic_ctx_set_max_iters(self.h, C.c_uint32(n)))

    def device_alloc(self, nbytes: int) -> int:
        p = self._lib.cniic_device_alloc(self.h, nbytes)
        if not p:
            raise CniicError(L.ERR_CUDA, "device allocation failed")
        return p

    def device_free(self, p: int):
        self._lib.cniic_device_free(self.h, p)

    def h2d(self, dptr: int, arr: np.ndarray):
        arr = np.ascontiguousarray(arr)
        self.check(self._lib.cniic_memcpy_h2d(self.h, dptr, _ptr(arr), arr.nbytes))

    def d2h(self, arr: np.ndarray, dptr: int):
        self.check(self._lib.cniic_memcpy_d2h(self.h, _ptr(arr), dptr, arr.nbytes))

    # ---- K-means: kmeans::cluster (kmeans.rs:21-39) ----
    def _result(self, st, cen, wts, asg, rc):
        return KMeansResult(cen, wts, asg, st.iterations, st.empty_events, st.moved_last, st.moved_total,
                            bool(st.converged), st.gpu_launches, st.device_ms, rc)

    def kmeans_rgb(self, rgb, k, counts=None, max_iters=0, tie=L.TIE_KEEP_CURRENT, want_assign=True,
                   allow_inactive=False) -> KMeansResult:
        """ColorCount points (clusterc.rs:68-114); counts=None clusters the points unweighted (per pixel)."""
        rgb = _u8(rgb).reshape(-1, 3)
        n = len(rgb)
        cnt = None if counts is None else np.ascontiguousarray(counts, dtype=np.uint32)
        cen = np.zeros((max(k, 1), 3), np.uint8)
        wts = np.zeros(max(k, 1), np.uint64)
        asg = np.zeros(n, np.uint16) if want_assign else None
        st = L.KMeansStats()
        rc = self._lib.cniic_kmeans_rgb(self.h, _ptr(rgb), _ptr(cnt), C.c_size_t(n), C.c_uint32(k),
                                        C.c_uint32(max_iters), tie, _ptr(cen), _ptr(wts), _ptr(asg), C.byref(st))
        self.check(rc, (L.ERR_TOO_FEW_ACTIVE,) if allow_inactive else ())
        return self._result(st, cen[:k].astype(np.int64), wts[:k], asg, rc)

    def kmeans_xyrgb(self, img, k, max_iters=0, tie=L.TIE_KEEP_CURRENT, want_assign=True,
                     allow_inactive=False) -> KMeansResult:
        """ColorPos points (clusterc.rs:148-153, 200-248): one point per pixel of img (h, w, 3)."""
        img = _u8(img)
        h, w = img.shape[:2]
        cxy = np.zeros((max(k, 1), 2), np.uint32)
        crgb = np.zeros((max(k, 1), 3), np.uint8)
        wts = np.zeros(max(k, 1), np.uint64)
        asg = np.zeros(h * w, np.uint16) if want_assign else None
        st = L.KMeansStats()
        rc = self._lib.cniic_kmeans_xyrgb(self.h, _ptr(img), C.c_uint32(w), C.c_uint32(h), C.c_uint32(k),
                                          C.c_uint32(max_iters), tie, _ptr(cxy), _ptr(crgb), _ptr(wts), _ptr(asg),
                                          C.byref(st))
        self.check(rc, (L.ERR_TOO_FEW_ACTIVE,) if allow_inactive else ())
        cen = np.concatenate([cxy[:k].astype(np.int64), crgb[:k].astype(np.int64)], axis=1)
        return self._result(st, cen, wts[:k], asg, rc)

    def kmeans_rgb_batch(self, images, k, max_iters=0, tie=L.TIE_KEEP_CURRENT, want_assign=True, allow_inactive=False):
        """`count` independent RGB images (per-pixel points), one K-means each, advanced in lock step with one launch per
        stage for the whole batch (cniic_kmeans_rgb_batch; bench.rs:27 runs one image per worker).  Returns a list of
        KMeansResult, identical to calling kmeans_rgb on every image."""
        imgs = [_u8(im).reshape(-1, 3) for im in images]
        count = len(imgs)
        ptrs = (C.c_void_p * count)(*[im.ctypes.data for im in imgs])
        ns = (C.c_size_t * count)(*[len(im) for im in imgs])
        cen = np.zeros((count, max(k, 1), 3), np.uint8)
        wts = np.zeros((count, max(k, 1)), np.uint64)
        asgs = [np.zeros(len(im), np.uint16) for im in imgs] if want_assign else None
        aptrs = (C.c_void_p * count)(*[a.ctypes.data for a in asgs]) if want_assign else None
        sts = (L.KMeansStats * count)()
        rc = self._lib.cniic_kmeans_rgb_batch(self.h, ptrs, ns, C.c_uint32(count), C.c_uint32(k), C.c_uint32(max_iters), tie,
                                              _ptr(cen), _ptr(wts), aptrs, sts)
        self.check(rc, (L.ERR_TOO_FEW_ACTIVE,) if allow_inactive else ())
        return [self._result(sts[i], cen[i, :k].astype(np.int64), wts[i, :k], asgs[i] if want_assign else None, rc)
                for i in range(count)]

    def kmeans_xyrgb_batch(self, images, k, max_iters=0, tie=L.TIE_KEEP_CURRENT, want_assign=True, allow_inactive=False):
        """The batch form for ColorPos points (cniic_kmeans_xyrgb_batch): images of any sizes, results as kmeans_xyrgb per image."""
        imgs = [_u8(im) for im in images]
        count = len(imgs)
        ptrs = (C.c_void_p * count)(*[im.ctypes.data for im in imgs])
        ws = (C.c_uint32 * count)(*[im.shape[1] for im in imgs])
        hs = (C.c_uint32 * count)(*[im.shape[0] for im in imgs])
        cxy = np.zeros((count, max(k, 1), 2), np.uint32)
        crgb = np.zeros((count, max(k, 1), 3), np.uint8)
        wts = np.zeros((count, max(k, 1)), np.uint64)
        asgs = [np.zeros(im.shape[0] * im.shape[1], np.uint16) for im in imgs] if want_assign else None
        aptrs = (C.c_void_p * count)(*[a.ctypes.data for a in asgs]) if want_assign else None
        sts = (L.KMeansStats * count)()
        rc = self._lib.cniic_kmeans_xyrgb_batch(self.h, ptrs, ws, hs, C.c_uint32(count), C.c_uint32(k), C.c_uint32(max_iters), tie,
                                                _ptr(cxy), _ptr(crgb), _ptr(wts), aptrs, sts)
        self.check(rc, (L.ERR_TOO_FEW_ACTIVE,) if allow_inactive else ())
        return [self._result(sts[i], np.concatenate([cxy[i, :k].astype(np.int64), crgb[i, :k].astype(np.int64)], axis=1), wts[i, :k],
                             asgs[i] if want_assign else None, rc) for i in range(count)]

    def kmeans_session(self, **kw) -> "KMeansSession":
        return KMeansSession(self, **kw)

    # ---- utils::count_freqs (utils.rs:4-16) in canonical ascending-key order ----
    def hist_rgb(self, rgb):
        """Distinct colours and their counts; keys = r<<16|g<<8|b ascending (clusterc.rs:21, huf.rs:30)."""
        rgb = _u8(rgb).reshape(-1, 3)
        n = len(rgb)
        cap = max(1, min(n, 1 << 24))
        keys = np.zeros(cap, np.uint32)
        cnts = np.zeros(cap, np.uint64)
        u = C.c_size_t(0)
        self.check(self._lib.cniic_hist_rgb(self.h, _ptr(rgb), C.c_size_t(n), _ptr(keys), _ptr(cnts), C.c_size_t(cap),
                                            C.byref(u)))
        return keys[:u.value].copy(), cnts[:u.value].copy()

    def hist_delta(self, img):
        """Histogram of the joint SignedColor symbols of the delta stream (hilbertc.rs:409-414 via huf.rs:30)."""
        img = _u8(img)
        h, w = img.shape[:2]
        cap = max(1, h * w)
        keys = np.zeros(cap, np.uint32)
        cnts = np.zeros(cap, np.uint64)
        u = C.c_size_t(0)
        self.check(self._lib.cniic_hist_delta(self.h, _ptr(img), C.c_uint32(w), C.c_uint32(h), _ptr(keys), _ptr(cnts),
                                              C.c_size_t(cap), C.byref(u)))
        return keys[:u.value].copy(), cnts[:u.value].copy()

    def recolor_rgb(self, rgb, keys, assign, centroids):
        """clusterc.rs:31-47: map every pixel to the colour of the centroid its colour belongs to."""
        rgb = _u8(rgb)
        flat = rgb.reshape(-1, 3)
        keys = np.ascontiguousarray(keys, dtype=np.uint32)
        assign = np.ascontiguousarray(assign, dtype=np.uint16)
        cen = _u8(centroids).reshape(-1, 3)
        out = np.zeros_like(flat)
        self.check(self._lib.cniic_recolor_rgb(self.h, _ptr(flat), C.c_size_t(len(flat)), _ptr(keys), _ptr(assign),
                                               C.c_size_t(len(keys)), _ptr(cen), C.c_uint32(len(cen)), _ptr(out)))
        return out.reshape(rgb.shape)

    def cluster_colors(self, img, k, max_iters=0, tie=L.TIE_KEEP_CURRENT, want_image=True):
        """Front half of ClusterColors::encode (clusterc.rs:19-47): returns (recoloured image or None, centroids, stats)."""
        img = _u8(img)
        h, w = img.shape[:2]
        out = np.zeros_like(img) if want_image else None
        cen = np.zeros((max(k, 1), 3), np.uint8)
        st = L.KMeansStats()
        self.check(self._lib.cniic_cluster_colors(self.h, _ptr(img), C.c_uint32(w), C.c_uint32(h), C.c_uint32(k),
                                                  C.c_uint32(max_iters), tie, _ptr(out), _ptr(cen), C.byref(st)))
        return out, cen[:k], st

    def cluster_colors_device(self, d_rgb: int, n_pixels: int, k: int, max_iters=0, tie=L.TIE_KEEP_CURRENT, d_out: int | None = None):
        """cniic_cluster_colors on an image resident in HBM: returns (centroids (k, 3) int32, number of unique colours, stats)."""
        cen = np.zeros((max(k, 1), 3), np.int32)
        st = L.KMeansStats()
        nu = C.c_size_t(0)
        self.check(self._lib.cniic_cluster_colors_device(self.h, C.c_void_p(d_rgb), C.c_size_t(n_pixels), C.c_uint32(k), C.c_uint32(max_iters), tie,
                                                         C.c_void_p(d_out) if d_out else None, _ptr(cen), C.byref(nu), C.byref(st)))
        return cen[:k], int(nu.value), st

    # ---- voronoi decode fill (clusterc.rs:179-186) ----
    def voronoi_fill(self, cxy, crgb, w, h):
        cxy = np.ascontiguousarray(cxy, dtype=np.uint32).reshape(-1, 2)
        crgb = _u8(crgb).reshape(-1, 3)
        out = np.zeros((h, w, 3), np.uint8)
        self.check(self._lib.cniic_voronoi_fill(self.h, _ptr(cxy), _ptr(crgb), C.c_uint32(len(cxy)), C.c_uint32(w),
                                                C.c_uint32(h), _ptr(out)))
        return out

    def voronoi_fill_rows(self, cxy, crgb, w, h, y0, h_local):
        """Rows [y0, y0 + h_local) of the fill: one rank's share of a row-sharded decode (no collective)."""
        cxy = np.ascontiguousarray(cxy, dtype=np.uint32).reshape(-1, 2)
        crgb = _u8(crgb).reshape(-1, 3)
        out = np.zeros((h_local, w, 3), np.uint8)
        d_cxy, d_crgb, d_out = self.device_alloc(cxy.nbytes), self.device_alloc(crgb.nbytes), self.device_alloc(max(16, out.nbytes))
        try:
            self.h2d(d_cxy, cxy)
            self.h2d(d_crgb, crgb)
            self.check(self._lib.cniic_voronoi_fill_device(self.h, C.c_void_p(d_cxy), C.c_void_p(d_crgb), C.c_uint32(len(cxy)), C.c_uint32(w),
                                                           C.c_uint32(h), C.c_uint32(y0), C.c_uint32(h_local), C.c_void_p(d_out)))
            if out.nbytes:
                self.d2h(out, d_out)
        finally:
            for p in (d_cxy, d_crgb, d_out):
                self.device_free(p)
        return out

    # ---- hilbert::iter / linearize (hilbert.rs:34-43), DiffStream (hilbertc.rs:449-477) ----
    def hilbert_xy(self, w, h):
        out = np.zeros((w * h, 2), np.uint32)
        self.check(self._lib.cniic_hilbert_xy(self.h, C.c_uint32(w), C.c_uint32(h), _ptr(out)))
        return out

    def hilbert_gather(self, img):
        img = _u8(img)
        h, w = img.shape[:2]
        out = np.zeros((w * h, 3), np.uint8)
        self.check(self._lib.cniic_hilbert_gather_rgb(self.h, _ptr(img), C.c_uint32(w), C.c_uint32(h), _ptr(out)))
        return out

    def delta(self, img):
        img = _u8(img)
        h, w = img.shape[:2]
        out = np.zeros((w * h, 3), np.int16)
        self.check(self._lib.cniic_delta_i16(self.h, _ptr(img), C.c_uint32(w), C.c_uint32(h), _ptr(out)))
        return out

    def delta_range_device(self, d_rgb: int, w: int, h: int, i_begin: int, i_end: int, d_out: int):
        """DiffStream of curve indices [i_begin, i_end) of a device-resident image into a device buffer (one rank's share)."""
        self.check(self._lib.cniic_delta_i16_range_device(self.h, C.c_void_p(d_rgb), C.c_uint32(w), C.c_uint32(h),
                                                          C.c_uint64(i_begin), C.c_uint64(i_end), C.c_void_p(d_out)))

    def hist_delta_range_device(self, d_rgb: int, w: int, h: int, i_begin: int, i_end: int):
        """Partial histogram (keys ascending, counts) of the delta symbols of curve indices [i_begin, i_end)."""
        cap = max(1, i_end - i_begin)
        keys = np.zeros(cap, np.uint32)
        cnts = np.zeros(cap, np.uint64)
        u = C.c_size_t(0)
        self.check(self._lib.cniic_hist_delta_range_device(self.h, C.c_void_p(d_rgb), C.c_uint32(w), C.c_uint32(h), C.c_uint64(i_begin),
                                                           C.c_uint64(i_end), _ptr(keys), _ptr(cnts), C.c_size_t(cap), C.byref(u)))
        return keys[:u.value].copy(), cnts[:u.value].copy()

    def undelta(self, diff, w, h):
        diff = np.ascontiguousarray(diff, dtype=np.int16)
        out = np.zeros((h, w, 3), np.uint8)
        rc = self._lib.cniic_undelta_rgb(self.h, _ptr(diff), C.c_uint32(w), C.c_uint32(h), _ptr(out))
        if rc == L.ERR_DECODE:
            return None  # a reconstructed channel left 0..255: FromDiff's try_into().unwrap() panics (hilbertc.rs:503-506)
        self.check(rc)
        return out

    def sse(self, a, b) -> int:
        """Exact integer sum of squared channel errors; MSE = sse / (w*h) (bench.rs:95-104)."""
        a, b = _u8(a), _u8(b)
        out = C.c_uint64(0)
        self.check(self._lib.cniic_sse_rgb(self.h, _ptr(a), _ptr(b), C.c_size_t(a.size // 3), C.byref(out)))
        return int(out.value)

    # ---- Codec::encode / Codec::decode (codec.rs:14-19) ----
    def codec_encode(self, codec: str, img) -> bytes:
        img = _u8(img)
        h, w = img.shape[:2]
        need = C.c_size_t(0)
        cap = 1 << 16
        out = np.zeros(cap, np.uint8)
        rc = self._lib.cniic_codec_encode(self.h, codec.encode(), _ptr(img), C.c_uint32(w), C.c_uint32(h), _ptr(out),
                                          C.c_size_t(cap), C.byref(need))
        if rc == L.ERR_BUFFER_TOO_SMALL and need.value > cap:
            # the finished stream waits in the ctx: fetch it (no second encoding pass)
            cap = need.value
            out = np.zeros(cap, np.uint8)
            rc = self._lib.cniic_codec_encode_fetch(self.h, _ptr(out), C.c_size_t(cap), C.byref(need))
        self.check(rc)
        return out[:need.value].tobytes()

    def codec_decode(self, codec: str, data: bytes):
        """Returns the decoded (h, w, 3) image or None (Codec::decode -> Option<Img>)."""
        buf = np.frombuffer(data, np.uint8)
        w, h = C.c_uint32(0), C.c_uint32(0)
        rc = self._lib.cniic_codec_decode(self.h, codec.encode(), _ptr(buf), C.c_size_t(len(buf)), C.byref(w),
                                          C.byref(h), None, C.c_size_t(0))
        if rc == L.ERR_DECODE:
            return None
        self.check(rc)
        if w.value * h.value > (1 << 28):
            return None
        out = np.zeros((h.value, w.value, 3), np.uint8)
        rc = self._lib.cniic_codec_decode(self.h, codec.encode(), _ptr(buf), C.c_size_t(len(buf)), C.byref(w),
                                          C.byref(h), _ptr(out), C.c_size_t(out.size // 3))
        if rc == L.ERR_DECODE:
            return None
        self.check(rc)
        return out


class KMeansSession:
    """Points resident in HBM across iterations (cniic_kmeans_open/reset/run/get)."""

    def __init__(self, ctx: Context, kind: int, k: int, rgb, n_local: int, n_total: int | None = None,
                 first_index: int = 0, w: int = 0, h_local: int = 0, y0: int = 0, weights=None,
                 tie: int = L.TIE_KEEP_CURRENT, on_device: bool = False, flags: int = 0):
        self.ctx = ctx
        self.k, self.D = k, (5 if kind == L.POINTS_XYRGB else 3)
        self.n_local = n_local
        d = L.KMeansDesc()
        d.kind, d.k, d.tie_rule = kind, k, tie
        d.n_local = n_local
        d.n_total = n_local if n_total is None else n_total
        d.first_index, d.w, d.h_local, d.y0 = first_index, w, h_local, y0
        self._keep = []
        if on_device:
            d.rgb = int(rgb)
            d.weights = int(weights) if weights is not None else None
        else:
            a = _u8(rgb)
            self._keep.append(a)
            d.rgb = a.ctypes.data
            if weights is not None:
                wa = np.ascontiguousarray(weights, dtype=np.uint32)
                self._keep.append(wa)
                d.weights = wa.ctypes.data
        d.points_on_device = 1 if on_device else 0
        d.flags = flags
        h = C.c_void_p()
        ctx.check(ctx._lib.cniic_kmeans_open(ctx.h, C.byref(d), C.byref(h)))
        self.h = h

    def reset(self, init_centroids=None):
        ic = None if init_centroids is None else np.ascontiguousarray(init_centroids, dtype=np.int32)
        self.ctx.check(self.ctx._lib.cniic_kmeans_reset(self.h, _ptr(ic)))

    def run(self, max_iters: int = 0) -> L.KMeansStats:
        st = L.KMeansStats()
        self.ctx.check(self.ctx._lib.cniic_kmeans_run(self.h, C.c_uint32(max_iters), C.byref(st)))
        return st

    def get(self, want_assign=True):
        cen = np.zeros((self.k, self.D), np.int32)
        wts = np.zeros(self.k, np.uint64)
        asg = np.zeros(self.n_local, np.uint16) if want_assign else None
        self.ctx.check(self.ctx._lib.cniic_kmeans_get(self.h, _ptr(cen), _ptr(wts), _ptr(asg)))
        return cen.astype(np.int64), wts, asg

    def close(self):
        if getattr(self, "h", None):
            self.ctx._lib.cniic_kmeans_close(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def kmeans_cluster(ctx: Context, kind: int, k: int, rgb, n_local: int, max_iters: int = 0, init_centroids=None, n_total: int | None = None,
                   first_index: int = 0, w: int = 0, h_local: int = 0, y0: int = 0, tie: int = L.TIE_KEEP_CURRENT, on_device: bool = False,
                   flags: int = 0, want_centroids: bool = True, want_assign: bool = False):
    """cniic_kmeans_cluster: kmeans::cluster on the described points in one C call (open + reset + run + get + close).
    Returns (centroids (k, D) int64 or None, weights or None, assign or None, stats)."""
    D = 5 if kind == L.POINTS_XYRGB else 3
    d = L.KMeansDesc()
    d.kind, d.k, d.tie_rule = kind, k, tie
    d.n_local = n_local
    d.n_total = n_local if n_total is None else n_total
    d.first_index, d.w, d.h_local, d.y0 = first_index, w, h_local, y0
    keep = None
    if on_device:
        d.rgb = int(rgb)
    else:
        keep = _u8(rgb)
        d.rgb = keep.ctypes.data
    d.points_on_device = 1 if on_device else 0
    d.flags = flags
    ic = None if init_centroids is None else np.ascontiguousarray(init_centroids, dtype=np.int32)
    cen = np.zeros((k, D), np.int32) if want_centroids else None
    wts = np.zeros(k, np.uint64) if want_centroids else None
    asg = np.zeros(n_local, np.uint16) if want_assign else None
    st = L.KMeansStats()
    ctx.check(ctx._lib.cniic_kmeans_cluster(ctx.h, C.byref(d), _ptr(ic), C.c_uint32(max_iters), _ptr(cen), _ptr(wts), _ptr(asg), C.byref(st)))
    return (None if cen is None else cen.astype(np.int64)), wts, asg, st


def kmeans_reset_batch(sessions):
    """cniic_kmeans_reset_batch: chunked init of every session, one launch per stage for the whole batch."""
    ctx = sessions[0].ctx
    hs = (C.c_void_p * len(sessions))(*[s.h.value for s in sessions])
    ctx.check(ctx._lib.cniic_kmeans_reset_batch(hs, C.c_uint32(len(sessions))))


def kmeans_run_batch(sessions, max_iters: int = 0):
    """cniic_kmeans_run_batch: Lloyd iterations of all sessions in lock step.  Returns one KMeansStats per session
    (device_ms / assign_ms_avg / gpu_launches describe the whole batch)."""
    ctx = sessions[0].ctx
    hs = (C.c_void_p * len(sessions))(*[s.h.value for s in sessions])
    sts = (L.KMeansStats * len(sessions))()
    ctx.check(ctx._lib.cniic_kmeans_run_batch(hs, C.c_uint32(len(sessions)), C.c_uint32(max_iters), sts))
    return list(sts)


# ---- synthetic images: identical on host and device ----
def synth_image_host(w: int, h: int, seed: int, n_blobs: int, y0: int = 0, h_total: int | None = None) -> np.ndarray:
    out = np.zeros((h, w, 3), np.uint8)
    rc = L.lib().cniic_synth_image_host(_ptr(out), C.c_uint32(w), C.c_uint32(h), C.c_uint32(y0),
                                        C.c_uint32(h_total or h), C.c_uint64(seed), C.c_uint32(n_blobs))
    if rc != L.OK:
        raise CniicError(rc, "synth_image_host")
    return out


def synth_image_device(ctx: Context, dptr: int, w: int, h: int, seed: int, n_blobs: int, y0: int = 0,
                       h_total: int | None = None):
    ctx.check(ctx._lib.cniic_synth_image_device(ctx.h, C.c_void_p(dptr), C.c_uint32(w), C.c_uint32(h), C.c_uint32(y0),
                                                C.c_uint32(h_total or h), C.c_uint64(seed), C.c_uint32(n_blobs)))
