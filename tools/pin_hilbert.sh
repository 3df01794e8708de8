#!/bin/bash
# One-command pin of the Hilbert order on a box that has cargo and network access (neither exists in the build image):
#   tools/pin_hilbert.sh /path/to/hkapp-cniic-checkout
# Builds the reference, lets it dump its own linearisation of coordinate-coded images (cniic --special=hilbert, main.rs:29-49) and
# diffs the orders against oracle.hilbert_xy -- the function the GPU kernels are tested against.  Exit 0 = pinned.
set -euo pipefail
REF=${1:?usage: pin_hilbert.sh <path to a checkout of hkapp/cniic>}
HERE=$(cd "$(dirname "$0")/.." && pwd)
WORK=$(mktemp -d)
python "$HERE/tools/pin_hilbert.py" make "$WORK/img"
(cd "$REF" && mkdir -p output && cargo run --release -- --special=hilbert "$WORK"/img/*.png)
make -C "$HERE/oracle" >/dev/null
python "$HERE/tools/pin_hilbert.py" check "$REF/output"
