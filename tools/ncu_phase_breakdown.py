#!/usr/bin/env python
"""ncu_phase_breakdown.py -- where a kernel's instructions and stall samples go, by source line.

    python tools/ncu_phase_breakdown.py <report.ncu-rep> <object.o> <mangled-name-fragment> [top_n]

Joins the per-instruction page of an `ncu --set full --import-source on` report (Instructions Executed, # Samples) with the
line table of the same kernel in the object file (`nvdisasm -g`; the object must be the build the report was taken from --
the script checks that the opcodes agree instruction by instruction) and prints the hottest source lines.  Inlined
helper code is attributed to the call site in the kernel body."""
import csv
import io
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict


def main():
    rep, obj, frag = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
    # NCU_FILTER (optional), e.g. "-k regex:km_assign_rgb_cull2 -s 1": selects ONE launch of a report that holds several
    page = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"] + os.environ.get("NCU_FILTER", "").split(),
                          capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(page)))
    hdr = rows[1]
    i_s, i_i = hdr.index("# Samples"), hdr.index("Instructions Executed")
    # a report with several launches repeats the two header rows per launch: take the first launch whose kernel name contains
    # NCU_KERNEL (default: the first launch)
    want = os.environ.get("NCU_KERNEL", "")
    starts = [j for j, r in enumerate(rows) if r and r[0] == "Kernel Name"]
    pick = next((j for j in starts if want in " ".join(rows[j][1:])), starts[0])
    nxt = next((j for j in starts if j > pick), len(rows))
    body = rows[pick + 2:nxt]
    prof = [(r[1].strip(), int(r[i_s]), int(r[i_i])) for r in body if len(r) > i_i]
    with tempfile.TemporaryDirectory() as td:
        subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=td, stdout=subprocess.DEVNULL)
        cubin = [f for f in os.listdir(td) if f.endswith(".cubin")][0]
        txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(td, cubin)], capture_output=True, text=True).stdout.split("\n")
    start = [i for i, l in enumerate(txt) if re.match(r"_Z\S*" + re.escape(frag) + r"\S*:$", l)][0]
    ins, cur = [], None
    for l in txt[start + 1:]:
        if l.startswith("//-----"):
            break
        m = re.search(r'//## File "([^"]+)", line (\d+)(?: inlined at "[^"]+", line (\d+))?', l)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)), int(m.group(3)) if m.group(3) else None)
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4})\*/\s+(.*?);", l)
        if m:
            ins.append((m.group(2).strip(), cur))
    op = lambda t: [w for w in t.split() if not w.startswith("@")][0]
    if len(ins) != len(prof) or any(op(a[0]) != op(b[0]) for a, b in zip(ins, prof)):
        sys.exit(f"object and report disagree ({len(ins)} vs {len(prof)} instructions): not the same build")
    tot, ts = sum(p[2] for p in prof), sum(p[1] for p in prof)
    by = defaultdict(lambda: [0, 0, 0])
    main_file = None
    for (text, li), (_, smp, n) in zip(ins, prof):
        if li is None:
            key = ("?", 0)
        else:
            main_file = main_file or (li[0] if li[2] is None else None)
            key = (li[0], li[1]) if li[2] is None else ("<inlined into>", li[2])
        by[key][0] += n
        by[key][1] += smp
        by[key][2] += 1
    print(f"{frag}: {len(ins)} SASS instructions, {tot} warp-instructions executed, {ts} stall samples")
    print(f"{'instr %':>8} {'smp %':>7} {'static':>6}  line")
    for (f, ln), v in sorted(by.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{100 * v[0] / tot:8.1f} {100 * v[1] / ts:7.1f} {v[2]:6d}  {f}:{ln}")


if __name__ == "__main__":
    main()
