"""cmx: B200-native vector-mix + flat inner-product top-k search (host side).

``import cmx.faiss as faiss`` gives the subset of the faiss Python API the
reference's onepass run scripts use; ``cmx.runloop`` holds the per-alpha run loop;
``cmx.engine`` is the thin wrapper over the C ABI (include/cmx.h, libcmx.so).
"""
from . import _lib  # noqa: F401
from .engine import Shard, merge_topk, mix_normalize  # noqa: F401
from .runloop import (  # noqa: F401
    format_alpha,
    parse_alpha_list,
    run_alpha_sweep,
    run_alpha_sweep_bilingual,
    collapse_run_max,
)

__version__ = "0.1.0"
