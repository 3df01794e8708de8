#!/usr/bin/env python
"""build_emu.py -- TEST INFRASTRUCTURE: compile cniic_b200/csrc/*.cu for the HOST against tests/emu/shim.

The kernel sources are used as they are; three purely syntactic rewrites make them C++ that g++ accepts:
  1. `kernel<<<grid, block, smem, stream>>>(args);`  ->  `emu::launch(emu::LaunchCfg(grid, block, smem, stream), [&]{ kernel(args); }, "kernel");`
  2. `extern __shared__ T name[];`                   ->  `T *name = reinterpret_cast<T *>(emu::dyn_smem());`
  3. `asm("dp4a.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));`  ->  `d = emu_ptx_dp4a_u32_u32(a, b, c);`
     (`asm volatile("griddepcontrol.wait;" ::: "memory");` and `...launch_dependents...` -> nothing: launches are sequential here)
Output: tests/emu/_build/libcniic_emu.so (git-ignored).  Loaded only by tests/test_emu_kernels.py -- never by cniic_b200.
"""
from __future__ import annotations

import glob
import hashlib
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "cniic_b200", "csrc")
OUT_DIR = os.path.join(HERE, "_build")
OUT_SO = os.path.join(OUT_DIR, "libcniic_emu.so")


def _match_paren(text: str, i: int) -> int:
    """text[i] == '(' -> index of the matching ')'."""
    depth = 0
    j = i
    while j < len(text):
        c = text[j]
        if c == "(":
            depth += 1
        elif c == ")":
            depth -= 1
            if depth == 0:
                return j
        elif c == '"':
            j += 1
            while text[j] != '"':
                j += 2 if text[j] == "\\" else 1
        j += 1
    raise ValueError("unbalanced parentheses")


def rewrite_launches(text: str) -> str:
    out, pos = [], 0
    while True:
        i = text.find("<<<", pos)
        if i < 0:
            break
        # kernel expression: identifier (with ::) optionally followed by <template args>, scanning backwards
        j = i
        while j > 0 and text[j - 1].isspace():
            j -= 1
        if text[j - 1] == ">":
            depth, j = 1, j - 1
            while depth:
                j -= 1
                depth += {">": 1, "<": -1}.get(text[j], 0)
        k = j
        while k > 0 and (text[k - 1].isalnum() or text[k - 1] in "_:"):
            k -= 1
        kernel = text[k:i].strip()
        m = re.compile(r">>>\s*\(").search(text, i)
        if not m:
            raise ValueError("launch without >>>(")
        cfg = text[i + 3:m.start()]
        a0 = m.end() - 1
        a1 = _match_paren(text, a0)
        args = text[a0 + 1:a1]
        name = re.sub(r"\s+", "", kernel)
        out.append(text[pos:k])
        out.append(f'emu::launch(emu::LaunchCfg({cfg}), [&]() {{ {kernel}({args}); }}, "{name}")')
        pos = a1 + 1
    out.append(text[pos:])
    return "".join(out)


ASM_RE = re.compile(r'asm\s*(?:volatile)?\s*\(\s*"([a-z0-9_.]+)\s+%0,\s*%1,\s*%2,\s*%3;"\s*:\s*"=r"\((\w+)\)\s*:\s*"r"\((\w+)\),\s*"r"\((\w+)\),\s*"r"\((\w+)\)\s*\)\s*;')
# programmatic-dependent-launch control instructions have no effect when launches run one after another
NOOP_ASM_RE = re.compile(r'asm\s*(?:volatile)?\s*\(\s*"(?:griddepcontrol\.(?:wait|launch_dependents);"\s*:::\s*"memory"|prefetch\.global\.L1 \[%0\];"\s*::\s*"l"\(\w+\))\s*\)\s*;')
TIMER_ASM_RE = re.compile(r'asm\s*volatile\s*\(\s*"mov\.u64 %0, %%globaltimer;"\s*:\s*"=l"\((\w+)\)\s*\)\s*;')  # timeline instrumentation: no clock here
SHARED_RE = re.compile(r"extern\s+__shared__\s+(?:__align__\(\d+\)\s+)?([\w:<> ]+?)\s+(\w+)\s*\[\s*\]\s*;")


# `#ifdef __CUDACC__ ... #endif  // __CUDACC__` blocks hold code only nvcc can compile (TMA / mbarrier inline PTX in tma.cuh); the shim
# provides host stand-ins of the same names with the same semantics (box copy with zero fill, phase-parity barrier)
CUDACC_ONLY_RE = re.compile(r"#ifdef __CUDACC__\n.*?#endif  // __CUDACC__\n", re.S)


def transform(text: str) -> str:
    text = CUDACC_ONLY_RE.sub("", text)
    text = ASM_RE.sub(lambda m: f"{m.group(2)} = emu_ptx_{m.group(1).replace('.', '_')}({m.group(3)}, {m.group(4)}, {m.group(5)});", text)
    text = NOOP_ASM_RE.sub(";", text)
    text = TIMER_ASM_RE.sub(lambda m: f"{m.group(1)} = 0;", text)
    text = SHARED_RE.sub(lambda m: f"{m.group(1)} *{m.group(2)} = reinterpret_cast<{m.group(1)} *>(emu::dyn_smem());", text)
    if "asm" in re.sub(r"//.*", "", text) and re.search(r"\basm\s*(volatile)?\s*\(", text):
        raise ValueError("an inline asm statement was not rewritten: extend ASM_RE / the shim")
    return rewrite_launches(text)


def build(verbose: bool = False, asan: bool | None = None) -> str:
    """asan (default: environment CNIIC_EMU_ASAN): AddressSanitizer build, loadable only into a process started with
    LD_PRELOAD=libasan (tests/emu/run_asan.sh); it also catches out-of-bounds READS of device blocks."""
    if asan is None:
        asan = bool(os.environ.get("CNIIC_EMU_ASAN"))
    os.makedirs(OUT_DIR, exist_ok=True)
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    deps = srcs + sorted(glob.glob(os.path.join(CSRC, "*.cuh"))) + sorted(glob.glob(os.path.join(CSRC, "*.h"))) + [
        os.path.join(ROOT, "include", "cniic_b200.h"), os.path.join(HERE, "emu_runtime.cpp"), os.path.join(HERE, "shim", "cuda_runtime.h"),
        os.path.join(HERE, "shim", "cub", "device", "device_radix_sort.cuh"), os.path.abspath(__file__)]
    h = hashlib.sha256()
    for p in deps:
        h.update(open(p, "rb").read())
    out_so = OUT_SO[:-3] + "_asan.so" if asan else OUT_SO
    stamp = os.path.join(OUT_DIR, "stamp_asan" if asan else "stamp")
    if os.path.exists(out_so) and os.path.exists(stamp) and open(stamp).read() == h.hexdigest():
        return out_so
    objs, procs = [], []
    # -fsanitize=alignment: the GPU faults on a misaligned 64/128-bit access, x86 would not notice -- let UBSan stand in
    flags = ["-std=c++17", "-O1", "-g", "-fPIC", "-fno-strict-aliasing", "-fsanitize=alignment", "-fno-sanitize-recover=alignment", "-Wall", "-Wno-unused-function", "-Wno-unknown-pragmas", "-Wno-unused-variable",
             "-Wno-sign-compare", "-I", os.path.join(HERE, "shim")]
    if asan:
        flags += ["-fsanitize=address", "-fno-omit-frame-pointer", "-DEMU_ASAN"]
    gen_dir = os.path.join(OUT_DIR, "gen_asan" if asan else "gen")
    os.makedirs(gen_dir, exist_ok=True)
    # headers are transformed too (common.cuh holds the inline asm); they keep their names so the #includes resolve
    for p in glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "*.h")):
        open(os.path.join(gen_dir, os.path.basename(p)), "w").write(transform(open(p).read()).replace('"../../include/cniic_b200.h"', f'"{ROOT}/include/cniic_b200.h"'))
    for p in srcs:
        g = os.path.join(gen_dir, os.path.basename(p)[:-3] + ".emu.cpp")
        open(g, "w").write(f'#line 1 "{p}"\n' + transform(open(p).read()))
        o = g[:-4] + ".o"
        objs.append(o)
        procs.append((g, subprocess.Popen(["g++", *flags, "-I", gen_dir, "-c", g, "-o", o], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    o = os.path.join(gen_dir, "emu_runtime.o")
    objs.append(o)
    procs.append(("emu_runtime.cpp", subprocess.Popen(["g++", *flags, "-c", os.path.join(HERE, "emu_runtime.cpp"), "-o", o], stdout=subprocess.PIPE,
                                                      stderr=subprocess.STDOUT, text=True)))
    failed = False
    for name, pr in procs:
        outp = pr.communicate()[0]
        if pr.returncode != 0 or (verbose and outp.strip()):
            sys.stderr.write(f"--- {name}\n{outp}\n")
        failed |= pr.returncode != 0
    if failed:
        raise RuntimeError("emulation build failed")
    subprocess.check_call(["g++", "-shared", "-fsanitize=alignment", *(["-fsanitize=address"] if asan else []), "-o", out_so, *objs, "-ldl"])
    open(stamp, "w").write(h.hexdigest())
    return out_so


def build_selftest() -> str:
    """tests/emu/selftest.cu -> tests/emu/_build/selftest (an executable exercising the emulation's own diagnostics)."""
    os.makedirs(os.path.join(OUT_DIR, "gen"), exist_ok=True)
    src = os.path.join(HERE, "selftest.cu")
    exe = os.path.join(OUT_DIR, "selftest")
    deps = [src, os.path.join(HERE, "emu_runtime.cpp"), os.path.join(HERE, "shim", "cuda_runtime.h"), os.path.abspath(__file__)]
    if os.path.exists(exe) and all(os.path.getmtime(exe) >= os.path.getmtime(d) for d in deps):
        return exe
    g = os.path.join(OUT_DIR, "gen", "selftest.emu.cpp")
    open(g, "w").write(f'#line 1 "{src}"\n' + transform(open(src).read()))
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-g", "-Wall", "-fsanitize=alignment", "-fno-sanitize-recover=alignment", "-I", os.path.join(HERE, "shim"), g, os.path.join(HERE, "emu_runtime.cpp"), "-o", exe])
    return exe


if __name__ == "__main__":
    print(build(verbose=True))
