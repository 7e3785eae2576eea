import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


# FPL_EMULATE=1: no GPU needed — binding.Engine loads the EMULATED build of the library (tests/simt_emu.py:build_library: every
# .cu of fastplong_b200/csrc compiled for the host behind the SIMT emulator, the C ABI and fpl_api.cu's host code included) and the
# `-m gpu` tests run on it, except the ones that need a second GPU or sizes the emulator takes minutes for (tests that hand the library
# device pointers use torch CPU tensors then: the emulated library's device memory is host memory).
EMULATE = os.environ.get("FPL_EMULATE", "") not in ("", "0")
NOT_UNDER_EMULATION = ("test_size_independent_properties_large", "test_gpu_multigpu", "test_config1_full_size_bit_exact",
                       "test_config1_shape_vs_oracle")


def pytest_sessionstart(session):
    if EMULATE:
        sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
        import simt_emu
        from fastplong_b200 import binding
        binding.LIB_PATH = simt_emu.build_library()
        binding._lib = None
        os.environ["FPL_GPU_BIN"] = simt_emu.build_binary()      # tests/test_gpu_binary.py: the drop-in CLI on the emulated library


def pytest_collection_modifyitems(config, items):
    if EMULATE:
        skip = pytest.mark.skip(reason="not under the emulator (CUDA memory from torch, a second GPU, or too large)")
        for item in items:
            if "gpu" in item.keywords and any(k in item.nodeid for k in NOT_UNDER_EMULATION):
                item.add_marker(skip)
        return
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
