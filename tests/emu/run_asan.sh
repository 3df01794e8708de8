#!/bin/bash
# TEST INFRASTRUCTURE: the emulated kernel tests under AddressSanitizer (out-of-bounds reads and writes of device blocks,
# shared-memory arrays and host buffers).  The interpreter is not instrumented, so libasan is preloaded.
#   bash tests/emu/run_asan.sh [pytest args]          e.g.  -k "rle or huffman"
cd "$(dirname "$0")/../.."
export CNIIC_EMU_ASAN=1
export ASAN_OPTIONS=detect_leaks=0:abort_on_error=1:detect_stack_use_after_return=0
export LD_PRELOAD="$(gcc -print-file-name=libasan.so)"
exec python -m pytest tests/test_emu_kernels.py -q -x -p no:cacheprovider "$@"
