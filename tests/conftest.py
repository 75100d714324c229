import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run under gpurun)")


def _native_library_present():
    """The product has no CPU fallback: importing the package without libmvae_b200.so raises.  The tests that only exercise
    host logic / the C-ABI surface build it when a compiler is around (nvcc cross-compiles without a GPU) and are skipped,
    not failed, on a machine that has neither the built library nor nvcc."""
    lib = os.path.join(ROOT, "molecular-vae_b200", "libmvae_b200.so")
    if os.path.exists(lib):
        return True
    try:
        import __graft_entry__ as ge
        ge.build()
    except Exception:
        return False
    return os.path.exists(lib)


NEEDS_LIBRARY = ("test_cabi_cpu", "test_checkpoint_cpu", "test_ddp_gloo", "test_featurizer", "test_host_logic_cpu")


def pytest_collection_modifyitems(config, items):
    if not _native_library_present():
        skip_lib = pytest.mark.skip(reason="libmvae_b200.so is not built and cannot be built here (no nvcc)")
        for item in items:
            if any(n in item.nodeid for n in NEEDS_LIBRARY) or "gpu" in item.keywords:
                item.add_marker(skip_lib)
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
