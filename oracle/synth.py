"""TEST INFRASTRUCTURE (part of oracle/): numpy restatement of the synthetic image generator of bench.py's workloads
(cniic_b200/csrc/synth.cu -- two octaves of hashed-lattice value noise plus +-8 per-channel hashed noise, integer only).

It exists so that `bench.py --impl reference` and the cpu_baseline leg can build the workload image WITHOUT loading the
product library (the reference arm imported cniic_b200 only for this).  tests/test_cpu_host.py checks that it
equals cniic_synth_image_host bit for bit.  Not part of the reference: hkapp/cniic reads image files (bench.rs:30).
"""
from __future__ import annotations

import numpy as np

_U = np.uint64
_M = (1 << 64) - 1


def _splitmix64(x: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        x = x + _U(0x9E3779B97F4A7C15)
        x = (x ^ (x >> _U(30))) * _U(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> _U(27))) * _U(0x94D049BB133111EB)
        return x ^ (x >> _U(31))


def _isqrt(v: int) -> int:
    import math
    return math.isqrt(v)


def _cell_of(w: int, h_total: int, n_blobs: int) -> int:
    c = _isqrt(w * h_total // max(1, n_blobs))
    return max(8, c)


def _bilerp(seed: int, x: np.ndarray, y: np.ndarray, cell: int) -> list[np.ndarray]:
    seed = _U(seed & _M)
    cell_u = _U(cell)
    gx, gy, fx, fy = x // cell_u, y // cell_u, x % cell_u, y % cell_u

    def lattice(ax, ay):
        return _splitmix64(seed ^ ((ay << _U(32)) | ax)) & _U(0xFFFFFFFF)

    c00, c10, c01, c11 = lattice(gx, gy), lattice(gx + _U(1), gy), lattice(gx, gy + _U(1)), lattice(gx + _U(1), gy + _U(1))
    w00, w10, w01, w11 = (cell_u - fx) * (cell_u - fy), fx * (cell_u - fy), (cell_u - fx) * fy, fx * fy
    out = []
    for ch in range(3):
        sh = _U(8 * ch)
        v = (w00 * ((c00 >> sh) & _U(0xFF)) + w10 * ((c10 >> sh) & _U(0xFF)) + w01 * ((c01 >> sh) & _U(0xFF)) + w11 * ((c11 >> sh) & _U(0xFF)))
        out.append(v // (cell_u * cell_u))
    return out


def synth_image(w: int, h: int, seed: int, n_blobs: int, y0: int = 0, h_total: int | None = None) -> np.ndarray:
    """(h, w, 3) uint8 rows [y0, y0 + h) of the w x h_total image of `seed` -- same bytes as cniic_synth_image_host."""
    h_total = h if not h_total else h_total
    cell = _cell_of(w, h_total, n_blobs)
    fine = max(2, cell // 4)
    out = np.empty((h, w, 3), np.uint8)
    rows_per_block = max(1, (1 << 22) // max(1, w))
    xs = np.arange(w, dtype=_U)
    for r0 in range(0, h, rows_per_block):
        r1 = min(h, r0 + rows_per_block)
        ys = np.arange(y0 + r0, y0 + r1, dtype=_U)
        x = np.broadcast_to(xs[None, :], (r1 - r0, w))
        y = np.broadcast_to(ys[:, None], (r1 - r0, w))
        a = _bilerp(seed, x, y, cell)
        b = _bilerp((seed & _M) ^ 0xA5A5A5A5DEADBEEF, x, y, fine)
        with np.errstate(over="ignore"):
            nz = _splitmix64(_U(seed & _M) ^ (_U(0x51ED270B4C3D) + y * _U(w) + x))
        for ch in range(3):
            v = ((_U(3) * a[ch] + b[ch]) // _U(4)).astype(np.int64) + ((nz >> _U(8 * ch)) % _U(17)).astype(np.int64) - 8
            out[r0:r1, :, ch] = np.clip(v, 0, 255).astype(np.uint8)
    return out
