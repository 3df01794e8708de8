#!/usr/bin/env python
"""pin_hilbert.py -- pin kit for the one part of the path whose parity is UNPINNED: the Hilbert order.

The reference takes its curve from the un-vendored crate zhang_hilbert 0.1.1 (hilbert.rs:3, 40-43); neither the crate nor cargo
exists in the build image, so oracle and library implement a generalized Hilbert scan that is only pinned by invariants.  On ANY
box with cargo + network this kit pins (or refutes) it in one command -- see tools/pin_hilbert.sh:

    python tools/pin_hilbert.py make  <dir>         coordinate-coded PNGs (pixel (x, y) = (x & 255, y & 255, x >> 8 | (y >> 8) << 4))
    cniic --special=hilbert <dir>/*.png             the reference dumps its three linearisations as CSV (main.rs:29-49)
    python tools/pin_hilbert.py check <csv dir>     decodes <name>.rect.hilbert.csv and compares with oracle.hilbert_xy(w, h)

`check` prints, per shape: MATCH, or the first differing index -- and whether one of the 8 dihedral images of our curve, forwards or
backwards, equals the reference's (then the difference is orientation only and the fix is the one transform named in the output,
applied inside hilbert_d2xy / oracle_hilbert_xy).  Exit status 0 only if every shape matches as is.
"""
import csv
import glob
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

# squares 2^n (what the tile kernels and BASELINE config 5 use), odd / even rectangles, both aspect ratios, thin strips
SHAPES = [(4, 4), (8, 8), (16, 16), (64, 64), (256, 256), (5, 3), (3, 5), (6, 4), (13, 29), (29, 13), (100, 7), (7, 100), (123, 77),
          (640, 360), (360, 640), (1, 9), (9, 1), (2, 2), (3, 3), (1024, 1024)]

TRANSFORMS = {
    "identity": lambda x, y, w, h: (x, y), "flip-x": lambda x, y, w, h: (w - 1 - x, y), "flip-y": lambda x, y, w, h: (x, h - 1 - y),
    "rotate-180": lambda x, y, w, h: (w - 1 - x, h - 1 - y),
    # the four that swap the axes only map a w x h curve onto a w x h rectangle when w == h
    "transpose": lambda x, y, w, h: (y, x), "anti-transpose": lambda x, y, w, h: (h - 1 - y, w - 1 - x),
    "rotate-90": lambda x, y, w, h: (h - 1 - y, x), "rotate-270": lambda x, y, w, h: (y, w - 1 - x),
}


def make(out_dir):
    from PIL import Image
    os.makedirs(out_dir, exist_ok=True)
    for w, h in SHAPES:
        x, y = np.meshgrid(np.arange(w), np.arange(h))
        img = np.stack([x & 255, y & 255, (x >> 8) | ((y >> 8) << 4)], axis=-1).astype(np.uint8)
        Image.fromarray(img, "RGB").save(os.path.join(out_dir, f"coords_{w}x{h}.png"))
    print(f"wrote {len(SHAPES)} images to {out_dir}")


def decode_csv(path):
    rows = list(csv.reader(open(path)))[1:]  # header "red,blue,green" (sic, main.rs:38) -- the values are r, g, b
    a = np.array(rows, dtype=np.int64)
    return np.stack([a[:, 0] | ((a[:, 2] & 15) << 8), a[:, 1] | ((a[:, 2] >> 4) << 8)], axis=1)


def check(csv_dir):
    import oracle as O
    bad = 0
    for path in sorted(glob.glob(os.path.join(csv_dir, "coords_*x*.rect.hilbert.csv")) + glob.glob(os.path.join(csv_dir, "coords_*x*.png.rect.hilbert.csv"))):
        name = os.path.basename(path)
        w, h = (int(v) for v in name.split("_")[1].split(".")[0].split("x"))
        ref = decode_csv(path)
        ours = O.hilbert_xy(w, h).astype(np.int64)
        if len(ref) != w * h:
            print(f"{w}x{h}: reference CSV has {len(ref)} rows, expected {w * h}")
            bad += 1
            continue
        if np.array_equal(ref, ours):
            print(f"{w}x{h}: MATCH")
            continue
        bad += 1
        first = int(np.nonzero((ref != ours).any(axis=1))[0][0])
        found = []
        for tname, t in TRANSFORMS.items():
            if w != h and tname in ("transpose", "anti-transpose", "rotate-90", "rotate-270"):
                continue
            tx, ty = t(ours[:, 0], ours[:, 1], w, h)
            cand = np.stack([tx, ty], axis=1)
            if np.array_equal(ref, cand):
                found.append(tname)
            if np.array_equal(ref, cand[::-1]):
                found.append(tname + " + reversed")
        print(f"{w}x{h}: DIFFERS at index {first}: reference {tuple(ref[first])}, ours {tuple(ours[first])}; "
              + (f"equal up to: {', '.join(found)}" if found else "no dihedral image / reversal of our curve matches -- a different scan (Zhang's block scan)"))
    print("all shapes match: the Hilbert order is PINNED" if bad == 0 else f"{bad} shape(s) differ")
    return 1 if bad else 0


if __name__ == "__main__":
    if len(sys.argv) == 3 and sys.argv[1] == "make":
        make(sys.argv[2])
    elif len(sys.argv) == 3 and sys.argv[1] == "check":
        sys.exit(check(sys.argv[2]))
    else:
        sys.exit(__doc__)
