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


def collapse_groups_ref(D, I, codes, ndocs):
    """numpy restatement of cmx_collapse_max (csrc/collapse.cu): per query the (base code, sign * rint(|score| * 1e6))
    groups in final order -- max over the 6-decimal values, descending, equal values in first-seen order."""
    import numpy as np

    nq, k = D.shape
    out_c = np.full((nq, k), -1, np.int32)
    out_v = np.zeros((nq, k), np.int64)
    out_n = np.zeros((nq,), np.int32)
    for r in range(nq):
        first, best = {}, {}
        for j in range(k):
            ix = int(I[r, j])
            if ix < 0 or ix >= ndocs:
                continue
            s = float(D[r, j])
            v = int(np.rint(abs(s) * 1e6)) * (-1 if s < 0 else 1)
            c = int(codes[ix])
            if c not in first:
                first[c], best[c] = j, v
            elif v > best[c]:
                best[c] = v
        order = sorted(first, key=lambda c: (-best[c], first[c]))
        out_n[r] = len(order)
        for g, c in enumerate(order):
            out_c[r, g], out_v[r, g] = c, best[c]
    return out_c, out_v, out_n
