#!/usr/bin/env python
"""fuzz_decode.py -- TEST INFRASTRUCTURE: hostile streams through the decoders on the emulated kernels.

    python tests/emu/fuzz_decode.py [seconds] [seed]           (under AddressSanitizer: see run_asan.sh for the environment)

Valid streams of every codec are truncated, extended and bit-flipped (header, trie and payload alike).  A decoder may reject a
stream (None / CniicError) or return an image of the advertised size -- and must agree with the sequential oracle decoder
whenever the oracle accepts the stream; it must never crash, hang or touch memory outside its buffers."""
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
ODEC = {"hufman": O.decode_hufman, "delta": O.decode_delta, "hilbert(rle)": O.decode_hilbert_rle, "voronoi(6)": O.decode_voronoi,
        "cluster-colors(5)": O.decode_hufman}
t0, cases, rejected, agreed = time.time(), 0, 0, 0
while time.time() - t0 < budget:
    w, h = int(rng.integers(1, 40)), int(rng.integers(1, 30))
    img = (rng.integers(0, 4, size=(h, w, 3)) * 60).astype(np.uint8) if rng.random() < 0.5 else rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)
    for expr, odec in ODEC.items():
        if "voronoi" in expr and w * h < 6 or "cluster" in expr and len(np.unique(img.reshape(-1, 3), axis=0)) < 5:
            continue
        c = codecs.Codec.from_str(ctx, expr, 3)
        data = bytearray(c.encode(img))
        for _ in range(6):
            d = bytearray(data)
            kind = rng.integers(0, 4)
            if kind == 0 and len(d) > 9:
                d = d[:int(rng.integers(8, len(d)))]
            elif kind == 1:
                d += bytes(rng.integers(0, 256, int(rng.integers(1, 40)), dtype=np.uint8))
            else:
                lo = 8 if rng.random() < 0.8 else 0  # mostly keep the dimensions (a flipped size is just another image size)
                for _ in range(int(rng.integers(1, 4))):
                    if len(d) > lo:
                        i = int(rng.integers(lo, len(d)))
                        d[i] ^= 1 << int(rng.integers(0, 8))
            d = bytes(d)
            if len(d) >= 8:
                dw, dh = int.from_bytes(d[:4], "little"), int.from_bytes(d[4:8], "little")
                if dw * dh > 1 << 16:
                    continue  # keep the emulated work small
            cases += 1
            try:
                g = c.decode(d)
            except cb.CniicError:
                g = None
            try:
                o = odec(d)
            except Exception:
                o = None
            if g is None:
                rejected += 1
                if o is not None and expr != "voronoi(6)":  # the library also rejects k > CNIIC_MAX_K, the oracle does not
                    print("REJECTED a stream the oracle decodes:", expr, "seed", seed, "case", cases)
                    sys.exit(1)
            elif o is not None:
                if g.shape != o.shape or not np.array_equal(g, o):
                    print("MISMATCH", expr, "seed", seed, "case", cases)
                    sys.exit(1)
                agreed += 1
print(f"decode fuzz ok: {cases} hostile streams ({rejected} rejected, {agreed} decoded like the oracle) in {time.time() - t0:.0f} s, seed {seed}")
