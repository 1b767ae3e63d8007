import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `pytest -m gpu`)")


def _cuda_device_present() -> bool:
    if not os.path.exists("/dev/nvidiactl") and not os.path.exists("/dev/nvidia0"):
        return False
    try:
        import torch
        return bool(torch.cuda.is_available())
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    """`pytest tests` on a box without a GPU: skip the gpu-marked tests instead of failing in chs_create (the product has no
    CPU fallback, so they cannot run there). On a GPU box nothing is skipped: a missing CUDA extension fails loudly."""
    gpu_items = [it for it in items if it.get_closest_marker("gpu")]
    if gpu_items and not _cuda_device_present():
        skip = pytest.mark.skip(reason="no CUDA device on this box (the CUDA path has no CPU fallback)")
        for it in gpu_items:
            it.add_marker(skip)
