#!/usr/bin/env python
"""sass_diff.py -- compare the SASS of two object files kernel by kernel (addresses and encodings stripped).

    python tools/sass_diff.py old.o new.o        (or two saved `cuobjdump -sass` dumps)

Used to prove that a source refactor (e.g. moving a kernel body into a __device__ function shared by two __global__
wrappers) leaves the machine code of the existing kernels untouched when no GPU is at hand to re-measure them."""
import re
import subprocess
import sys


def load(path):
    text = open(path).read() if path.endswith(".sass") else subprocess.check_output(["cuobjdump", "-sass", path], text=True)
    funcs, cur = {}, None
    for line in text.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.check_output(["c++filt", m.group(1)], text=True).strip()
            name = re.sub(r"\(anonymous namespace\)::", "", name)
            name = re.sub(r"_GLOBAL__N__\w+::", "", name)
            cur = funcs.setdefault(name, [])
            continue
        if cur is None:
            continue
        line = re.sub(r"/\*[0-9a-f]{4,}\*/", "", line)
        line = re.sub(r"/\* 0x[0-9a-f]+ \*/", "", line).strip()
        if line and not line.startswith("."):
            cur.append(line)
    return funcs


def main():
    a, b = load(sys.argv[1]), load(sys.argv[2])
    rc = 0
    for name in sorted(set(a) | set(b)):
        if name not in a:
            print(f"NEW        {len(b[name]):6d} instr  {name}")
        elif name not in b:
            print(f"GONE       {len(a[name]):6d} instr  {name}")
            rc = 1
        elif a[name] == b[name]:
            print(f"identical  {len(a[name]):6d} instr  {name}")
        else:
            print(f"DIFFERENT  {len(a[name]):6d} -> {len(b[name]):6d} instr  {name}")
            rc = 1
    return rc


if __name__ == "__main__":
    sys.exit(main())
