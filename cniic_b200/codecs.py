"""Host-side mirror of the reference's ``Codec`` trait (codec.rs:14-19) for the codecs on the hot path.

Same names, argument meaning and error behaviour as the reference: ``encode`` returns the byte stream
(``io::Write`` sink), ``decode`` returns an image or ``None`` (``Option<Img>``), ``name()`` / ``is_lossless()`` match
clusterc.rs:59-65,191-197, hilbertc.rs:81-95,433-439 and hufc.rs:42-48.  ``from_str`` accepts the reference's codec
expressions (``AnyCodec::from_str``, codec.rs:43-58).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L
from .api import Context


def codec_name(expr: str) -> str:
    buf = C.create_string_buffer(64)
    rc = L.lib().cniic_codec_name(expr.encode(), buf, C.c_size_t(64))
    if rc != L.OK:
        raise ValueError(f"unknown codec expression {expr!r}")
    return buf.value.decode()


class Codec:
    """One codec bound to a GPU context.  ``max_iters`` = 0 reproduces the reference (run until converged)."""

    LOSSLESS = {"delta", "Hufman", "hilbert-rle"}

    def __init__(self, ctx: Context, expr: str, max_iters: int = 0):
        self.ctx, self.expr, self.max_iters = ctx, expr, max_iters
        self._name = codec_name(expr)

    @classmethod
    def from_str(cls, ctx: Context, expr: str, max_iters: int = 0) -> "Codec":
        return cls(ctx, expr, max_iters)

    def name(self) -> str:
        return self._name

    def is_lossless(self) -> bool:
        return self._name in self.LOSSLESS

    def encode(self, img: np.ndarray) -> bytes:
        self.ctx.set_max_iters(self.max_iters)
        return self.ctx.codec_encode(self.expr, img)

    def decode(self, data: bytes):
        return self.ctx.codec_decode(self.expr, data)


def ClusterColors(ctx, k, max_iters=0):  # clusterc.rs:14
    return Codec(ctx, f"cluster-colors({k})", max_iters)


def VoronoiCluster(ctx, k, max_iters=0):  # clusterc.rs:145
    return Codec(ctx, f"voronoi({k})", max_iters)


def Delta(ctx):  # hilbertc.rs:400
    return Codec(ctx, "delta")


def Hufman(ctx):  # hufc.rs:9
    return Codec(ctx, "hufman")


def HilbertRle(ctx):  # hilbertc.rs:12 with CompressionMethod::RLE(0.0)
    return Codec(ctx, "hilbert(rle)")
