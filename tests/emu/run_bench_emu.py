#!/usr/bin/env python
"""run_bench_emu.py -- TEST INFRASTRUCTURE: run bench.py's real main() on the emulated kernels with tiny workloads.

    python tests/emu/run_bench_emu.py --workload c2 --steps 2 --warmup 1

Checks the driver's control flow and JSON contract without a GPU (bench.py cannot otherwise run in the build container).
The numbers it prints are host timings of the emulation: meaningless as performance."""
import ctypes
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(HERE, "fake_torch"))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

import build_emu  # noqa: E402
from cniic_b200 import _lib as L  # noqa: E402

L._lib = L._declare(ctypes.CDLL(build_emu.build()))

import bench  # noqa: E402
import bench_stages  # noqa: E402

# same shapes of work, sizes the emulation finishes in seconds
bench.WORKLOADS = {
    "c1": ("rgb", 64, 48, 16, 6, "emulated c1", "weak"),
    "c2": ("rgb", 160, 120, 256, 12, "emulated c2", "weak"),
    "c3": ("xyrgb", 192, 96, 64, 16, "emulated c3", "strong"),
    "c4": ("rgb", 64, 32, 16, 4, "emulated c4", "weak"),
}
bench.C4_BATCH = 3
bench_stages.SIZES = {"c5": (256, 256, 64), "fill": (192, 96, 64)}
if "--batch" not in sys.argv:
    sys.argv += ["--batch", "3"]
sys.exit(bench.main())
