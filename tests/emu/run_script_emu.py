#!/usr/bin/env python
"""run_script_emu.py -- TEST INFRASTRUCTURE: run a GPU-box helper script (tools/*.py) against the emulated kernels.

    python tests/emu/run_script_emu.py tools/sanitize_small.py

The script sees the normal cniic_b200 package; only the ctypes loader is pointed at tests/emu/_build/libcniic_emu.so."""
import ctypes
import os
import runpy
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import build_emu  # noqa: E402
from cniic_b200 import _lib as L  # noqa: E402

L._lib = L._declare(ctypes.CDLL(build_emu.build()))
script = sys.argv[1]
sys.argv = sys.argv[1:]
runpy.run_path(script, run_name="__main__")
