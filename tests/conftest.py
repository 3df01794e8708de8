import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


# emulated-kernel cases that take more than ~5 s on the host (tests/test_emu_kernels.py); run them with CNIIC_EMU_FULL=1
EMU_SLOW = ("many_symbols", "333-257-256", "640-360-2048", "unique_colours[256]", "hist_delta_extremes", "513-65-4096", "100003-1000",
            "0-512-512-16", "1024-96-100", "pipeline[64]", "k256_4096", "k2048_8k", "stages_8192", "codecs_fullsize",
            "sharded_code_path[c5]", "stage_workloads_run[c5]", "histogram[256-256-8]", "histogram[64-64-5]", "at_full_size",
            "edge_cases[2]", "edge_cases[3]", "edge_cases[8]", "edge_cases[13]", "json_contract[c2]")


def pytest_collection_modifyitems(config, items):
    if not os.environ.get("CNIIC_EMU_FULL"):
        slow = pytest.mark.skip(reason="slow under the host emulation; set CNIIC_EMU_FULL=1")
        for item in items:
            if item.module.__name__ == "test_emu_kernels" and any(s in item.name for s in EMU_SLOW):
                item.add_marker(slow)
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
