"""The CUDA kernel sources, compiled for the HOST and run under the fiber emulation of tests/emu/ -- logic tests of the
real kernels (barriers, warp collectives, shared-memory protocols, index arithmetic) on a machine without a GPU.

TEST INFRASTRUCTURE ONLY.  tests/emu/build_emu.py compiles cniic_b200/csrc/*.cu against a stand-in <cuda_runtime.h>
into tests/emu/_build/libcniic_emu.so; this module points the ctypes loader at that library for its own duration
(module-scoped fixture, restored afterwards) and re-runs the bodies of the GPU parity tests against the oracle.
The product never loads the emulated library and still has no CPU path (tests/test_cpu_host.py checks both); a
green run here says nothing about performance and does not replace `pytest -m gpu` on a B200.

Sizes: the slow cases (dense 2^24 / 511^3 histogram bins, k = 2048 tables) are skipped unless CNIIC_EMU_FULL=1
(`CNIIC_EMU_FULL=1 python -m pytest tests/test_emu_kernels.py` also runs the BASELINE-size cases of test_gpu_fullsize.py and
takes ~20 minutes).
"""
import ctypes
import os
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "emu"))

import cniic_b200 as cb  # noqa: E402
from cniic_b200 import _lib as L  # noqa: E402

if sys.platform != "linux":
    pytest.skip("the emulation needs Linux (mmap'ed fiber stacks, ELF shared object)", allow_module_level=True)


@pytest.fixture(scope="module", autouse=True)
def emu_lib():
    import build_emu
    so = build_emu.build()
    saved = L._lib
    L._lib = L._declare(ctypes.CDLL(so))
    yield so
    L._lib = saved


@pytest.fixture(scope="module")
def ctx(emu_lib):
    c = cb.Context()
    yield c
    c.close()


def _reexport(module):
    """Collect the GPU test functions of `module` here, under the emulated context (module-level gpu marks stay behind)."""
    for name, fn in vars(module).items():
        if name.startswith("test_") and callable(fn):
            globals()["test_emu_" + name[5:]] = fn


import test_gpu_fullsize  # noqa: E402
import test_gpu_golden  # noqa: E402
import test_gpu_parity  # noqa: E402

_reexport(test_gpu_parity)
_reexport(test_gpu_golden)
_reexport(test_gpu_fullsize)  # BASELINE sizes: CNIIC_EMU_FULL=1 only (minutes per case even on the host)


def test_emu_library_is_the_emulated_one(ctx, emu_lib):
    assert "tests/emu/_build" in emu_lib.replace(os.sep, "/")
    assert ctx._lib._name == emu_lib
    assert cb.Context is not None and L.SO_PATH.endswith("libcniic_b200.so")  # the product path is untouched


def test_emu_detects_out_of_bounds_and_deadlocks():
    """The emulation's own checks fire: run tiny bad kernels in a child process and expect an abort with a diagnosis."""
    import subprocess
    import build_emu
    exe = build_emu.build_selftest()
    for case, needle in (("oob", "OUT-OF-BOUNDS WRITE"), ("deadlock", "DEADLOCK"), ("divergent", "DIFFERENT collectives"), ("misaligned", "misaligned address"),
                         ("ok", "selftest ok")):
        r = subprocess.run([exe, case], capture_output=True, text=True, timeout=60)
        assert needle in (r.stderr + r.stdout), (case, r.stderr, r.stdout)
        assert (r.returncode == 0) == (case == "ok")
    # a missing __syncthreads() that the forward sweep hides is exposed by the other scheduling orders
    assert subprocess.run([exe, "race"], capture_output=True, env=dict(os.environ, EMU_SCHED="forward")).returncode == 0
    assert subprocess.run([exe, "race"], capture_output=True, env=dict(os.environ, EMU_SCHED="reverse")).returncode == 3
    assert subprocess.run([exe, "race"], capture_output=True, env=dict(os.environ, EMU_SCHED="random:1")).returncode == 3


@pytest.mark.parametrize("workload", ["c5", "fill"])
def test_emu_bench_stage_workloads_run(workload):
    """bench_stages.run() (integer stages, voronoi fill) on the emulated kernels: control flow + JSON contract."""
    import json
    import subprocess
    r = subprocess.run([sys.executable, os.path.join(HERE, "emu", "run_bench_emu.py"), "--workload", workload, "--steps", "2", "--warmup", "1", "--no-cpu"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["metric"] == "Mpix/s" and line["value"] > 0 and line["n_gpus"] == 1 and line["steps"] == 2 and line["warmup"] >= 3
    rf = line["roofline"]
    assert rf["bound"] == "hbm" and rf["achieved"] > 0 and abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-12
    assert line["e2e"]["value"] > 0 and line["e2e"]["h2d_bytes_per_step"] > 0 and line["e2e"]["d2h_bytes_per_step"] > 0
    assert line["gpu_launches"] > 0 and "workload" in line["config"]


@pytest.mark.parametrize("workload", ["c5", "fill"])
def test_emu_bench_stage_workloads_sharded_code_path(workload):
    """The N > 1 branches of the stage benchmarks (curve / row shards) walked by lone processes that pretend to be ranks 0 and 2 of 3
    (stand-in torch.distributed: collectives are identities): rank 0 prints the line, the others stay silent."""
    import json
    import subprocess
    for rank in (0, 2):
        env = dict(os.environ, RANK=str(rank), LOCAL_RANK="0", WORLD_SIZE="3")
        r = subprocess.run([sys.executable, os.path.join(HERE, "emu", "run_bench_emu.py"), "--workload", workload, "--gpus", "3", "--steps", "2",
                            "--warmup", "1", "--no-cpu"], capture_output=True, text=True, timeout=600, env=env)
        assert r.returncode == 0, r.stderr[-2000:]
        if rank:
            assert r.stdout.strip() == ""
        else:
            line = json.loads(r.stdout.strip().splitlines()[-1])
            assert line["n_gpus"] == 3 and line["scaling"] == "strong" and "sharded over 3 GPUs" in line["config"]["workload"]
            assert line["e2e"]["d2h_bytes_per_step"] > 0


@pytest.mark.parametrize("workload", ["c2", "c4", "c3", None])
def test_emu_bench_main_runs_and_keeps_the_json_contract(workload):
    """bench.py's real main() on the emulated kernels (tiny workloads, stand-in torch): control flow + JSON contract.
    workload None = the default run: C3 (north-star sharded configuration) as the line, C2 as its `secondary` object."""
    import json
    import subprocess
    r = subprocess.run([sys.executable, os.path.join(HERE, "emu", "run_bench_emu.py")] + (["--workload", workload] if workload else []) +
                       ["--steps", "2", "--warmup", "1"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
                "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"):
        assert key in line, key
    assert line["steps"] == 2 and line["warmup"] >= 3 and line["n_gpus"] == 1 and line["value"] > 0
    assert line["gpu_launches"] > 0 and line["iterations_run"] == 2 * 10
    rf = line["roofline"]
    assert rf["bound"] == "hbm" and rf["unit"] == "GB/s" and rf["achieved"] > 0 and abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-12
    assert rf["fp32"]["brute_force_kernel"]["launch_ms"] > 0
    e = line["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["value"] > 0
    assert "workload" in line["config"]
    assert line["config"]["name"] == (workload or "c3") and "per_iteration_overhead_ms" in line["breakdown"]
    if workload is None:
        sec = line["secondary"]
        assert line["scaling"] == "strong" and sec["config"]["name"] == "c2" and sec["value"] > 0 and sec["e2e"]["value"] > 0
    else:
        assert "secondary" not in line
    if workload == "c4":
        assert rf["kernel"].endswith("_batch") and line["config"]["images_per_step_per_gpu"] == 3 and e["api"] == "cniic_kmeans_rgb_batch"


def test_emu_graft_entry_smoke_runs(ctx):
    """__graft_entry__.smoke() (what the driver runs on the GPU box before the bench) walks on the emulated kernels too."""
    sys.path.insert(0, os.path.dirname(HERE))
    import __graft_entry__ as g
    g.smoke()
    g.smoke_extras()
