"""pytest configuration: registers the `gpu` marker and puts the repo root on sys.path so that
`oracle` (test infrastructure) and the package loader `sbp_b200` resolve."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a real CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    # GPU tests are selected with `-m gpu`; if someone runs the full suite on a CPU-only
    # machine they are skipped rather than failed.
    try:
        import torch

        has_gpu = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
