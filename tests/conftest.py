"""pytest config: marker registration and import paths.

``-m "not gpu"`` runs on a CPU-only box; ``-m gpu`` needs a B200 and calls the
CUDA path through the C ABI (libcmx.so).
"""
import os
import pathlib
import sys

import pytest

ROOT = pathlib.Path(__file__).resolve().parent.parent
PKG = ROOT / "codemix-dense-retrieval_b200"
for p in (str(ROOT), str(PKG)):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def pytest_collection_modifyitems(config, items):
    try:
        import torch

        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device visible")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(params=["rescore", "split"])
def precision(request):
    """Run a GPU test under both tensor-path arithmetic modes."""
    from cmx import _lib

    _lib.set_default_precision(request.param)
    yield request.param
    _lib.set_default_precision("rescore")
