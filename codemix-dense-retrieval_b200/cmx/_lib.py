"""ctypes binding of libcmx.so (the C ABI declared in include/cmx.h).

The product path has no CPU fallback: if the shared library is missing this
module raises, and every compute entry point raises RuntimeError when no CUDA
device is usable.
"""
from __future__ import annotations

import ctypes as C
import os
import pathlib

_HERE = pathlib.Path(__file__).resolve().parent
LIB_PATH = pathlib.Path(os.environ.get("CMX_LIB", _HERE.parent / "lib" / "libcmx.so"))

PATH_AUTO, PATH_STREAM, PATH_TENSOR = 0, 1, 2
PRECISION_SPLIT, PRECISION_RESCORE = 0, 1
MAX_K = 2048


class SearchStats(C.Structure):
    _fields_ = [
        ("path", C.c_int32),
        ("slabs", C.c_int32),
        ("reruns", C.c_int32),
        ("launches", C.c_int32),
        ("nq", C.c_int64),
        ("ntotal", C.c_int64),
        ("score_ms", C.c_float),
        ("select_ms", C.c_float),
        ("total_ms", C.c_float),
        ("score_launches", C.c_int32),
        ("select_launches", C.c_int32),
    ]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise RuntimeError(
            f"libcmx.so not found at {LIB_PATH}: build it with "
            "`python codemix-dense-retrieval_b200/build.py` (nvcc, sm_100a). There is no CPU fallback."
        )
    L = C.CDLL(str(LIB_PATH))
    vp, i32, i64, f32p = C.c_void_p, C.c_int, C.c_int64, C.c_void_p
    L.cmx_last_error.restype = C.c_char_p
    L.cmx_last_error.argtypes = []
    L.cmx_version.restype = i32
    L.cmx_launch_count.restype = C.c_uint64
    L.cmx_device_count.argtypes = [C.POINTER(i32)]
    L.cmx_index_create.argtypes = [i32, i32, C.POINTER(vp)]
    L.cmx_index_free.argtypes = [vp]
    L.cmx_index_reserve.argtypes = [vp, i64]
    L.cmx_index_add.argtypes = [vp, f32p, i64, i32]
    L.cmx_index_reset.argtypes = [vp]
    L.cmx_index_add_from_file.argtypes = [vp, C.c_char_p, i64, i64, i32, vp]
    L.cmx_index_add_gather.argtypes = [vp, vp, vp, i64]
    L.cmx_read_file.argtypes = [C.c_char_p, i64, i64, vp, i32]
    L.cmx_index_ntotal.argtypes = [vp, C.POINTER(i64)]
    L.cmx_index_memory.argtypes = [vp, vp]
    L.cmx_index_dim.argtypes = [vp, C.POINTER(i32)]
    L.cmx_index_device.argtypes = [vp, C.POINTER(i32)]
    L.cmx_index_reconstruct.argtypes = [vp, i64, i64, f32p, i32]
    L.cmx_index_data.argtypes = [vp, C.POINTER(vp)]
    L.cmx_index_search.argtypes = [vp, f32p, i64, i32, f32p, vp, i32, i64, i32, vp]
    L.cmx_mix_normalize.argtypes = [f32p, f32p, i64, i32, C.POINTER(C.c_double), i32, f32p, vp, i32, i32, vp]
    L.cmx_search_mixed.argtypes = [vp, f32p, f32p, i64, C.POINTER(C.c_double), i32, i32, f32p, vp, vp, i32, i64, i32, vp]
    L.cmx_search_prepare.argtypes = [vp, f32p, f32p, i64, C.POINTER(C.c_double), i32, C.POINTER(vp), vp]
    L.cmx_index_export_bounds.argtypes = [vp, vp, vp]
    L.cmx_search_begin.argtypes = [vp, f32p, i64, i32, i64, vp, i32, C.c_float, vp, vp, vp]
    L.cmx_search_end.argtypes = [vp, vp, i32, f32p, vp, vp]
    L.cmx_union_kth.argtypes = [vp, i32, i64, i32, i64, i64, vp, i32, vp, i32, vp, i32, vp]
    L.cmx_peer_broadcast.argtypes = [vp, vp, i32, i64, i32, vp]
    L.cmx_host_register.argtypes = [vp, i64, C.POINTER(vp)]
    L.cmx_host_unregister.argtypes = [vp]
    L.cmx_enable_peer_access.argtypes = [i32, i32]
    L.cmx_merge_topk.argtypes = [f32p, vp, i32, i64, i32, f32p, vp, i32, i32, vp]
    L.cmx_merge_topk_peers.argtypes = [vp, vp, i32, i64, i32, i64, i64, vp, vp, i32, i32, vp]
    cpp, i64p = C.POINTER(C.c_char_p), C.POINTER(i64)
    L.cmx_trec_mono.argtypes = [vp, vp, i64, i32, C.c_char_p, vp, C.c_char_p, vp, vp, i64, C.c_char_p, i32,
                                C.POINTER(vp), i64p]
    L.cmx_trec_bilingual.argtypes = [vp, vp, i64, i32, C.c_char_p, vp, C.c_char_p, vp, i64, vp, C.c_char_p, vp, i64,
                                     C.c_char_p, i32, C.POINTER(vp), i64p, C.POINTER(vp), i64p]
    L.cmx_trec_mono_file.argtypes = [vp, vp, i64, i32, C.c_char_p, vp, C.c_char_p, vp, vp, i64, C.c_char_p, i32,
                                     C.c_char_p, i64p]
    L.cmx_trec_bilingual_file.argtypes = [vp, vp, i64, i32, C.c_char_p, vp, C.c_char_p, vp, i64, vp, C.c_char_p, vp, i64,
                                          C.c_char_p, i32, C.c_char_p, C.c_char_p, i64p, i64p]
    L.cmx_trec_bilingual_file_pre.argtypes = [vp, vp, i64, i32, C.c_char_p, vp, C.c_char_p, vp, i64, vp, C.c_char_p, vp, i64,
                                              vp, vp, vp, C.c_char_p, i32, C.c_char_p, C.c_char_p, i64p, i64p]
    L.cmx_collapse_max.argtypes = [vp, vp, i64, i32, vp, i64, vp, vp, vp, C.POINTER(i32), i32, vp]
    L.cmx_free_text.argtypes = [vp]
    L.cmx_free_text.restype = None
    L.cmx_index_last_stats.argtypes = [vp, C.POINTER(SearchStats)]
    L.cmx_set_profiling.argtypes = [i32]
    L.cmx_index_set_cand_capacity.argtypes = [vp, i32]
    L.cmx_index_set_precision.argtypes = [vp, i32]
    L.cmx_set_default_precision.argtypes = [i32]
    L.cmx_index_error_bounds.argtypes = [vp, vp]
    L.cmx_index_raise_error_bounds.argtypes = [vp, vp]
    L.cmx_debug_set_tensor_tile.argtypes = [i32]
    L.cmx_debug_set_stream_variant.argtypes = [i32]
    L.cmx_debug_set_tensor_flags.argtypes = [i32]
    L.cmx_debug_set_tensor_pair.argtypes = [i32]
    L.cmx_debug_set_tensor_small.argtypes = [i32]
    L.cmx_debug_set_tensor_window.argtypes = [i32]
    L.cmx_debug_set_block_order.argtypes = [i32]
    L.cmx_debug_set_speculate.argtypes = [i32]
    L.cmx_debug_set_small_first.argtypes = [i32]
    L.cmx_debug_set_mapped_outputs.argtypes = [i32]
    L.cmx_debug_plan_slabs.argtypes = [i64, i32, i32, i32, i32, i32, vp, i32, vp, vp, vp]
    L.cmx_debug_plan_ranks.argtypes = [i64, i32, i32, i32, i32, i32, vp, i32]
    L.cmx_debug_set_prescore.argtypes = [i32]
    L.cmx_debug_inject_begin_status.argtypes = [i32, C.c_uint32]
    L.cmx_debug_set_prescore_params.argtypes = [C.c_double, i32, i32]
    L.cmx_debug_set_prescore_min_rows.argtypes = [i64]
    L.cmx_debug_approx_scores.argtypes = [vp, f32p, i64, i64, i64, f32p, vp, f32p, vp]
    L.cmx_debug_block_perm.argtypes = [i64]
    L.cmx_debug_block_perm.restype = C.c_uint64
    for name in (
        "cmx_device_count cmx_index_create cmx_index_free cmx_index_reserve cmx_index_add cmx_index_reset cmx_index_add_from_file cmx_index_add_gather cmx_read_file "
        "cmx_index_ntotal cmx_index_memory cmx_index_dim cmx_index_device cmx_index_reconstruct cmx_index_data cmx_index_search "
        "cmx_mix_normalize cmx_search_mixed cmx_search_prepare cmx_index_export_bounds cmx_search_begin cmx_search_end cmx_union_kth cmx_peer_broadcast cmx_host_register cmx_host_unregister cmx_enable_peer_access cmx_debug_plan_ranks cmx_debug_set_prescore cmx_debug_inject_begin_status cmx_debug_set_prescore_params cmx_debug_set_prescore_min_rows cmx_debug_approx_scores cmx_merge_topk cmx_merge_topk_peers cmx_trec_mono cmx_trec_bilingual cmx_trec_mono_file cmx_trec_bilingual_file cmx_trec_bilingual_file_pre cmx_collapse_max cmx_index_last_stats cmx_set_profiling "
        "cmx_index_set_cand_capacity cmx_index_error_bounds cmx_index_raise_error_bounds cmx_index_set_precision cmx_set_default_precision cmx_debug_set_tensor_tile cmx_debug_set_stream_variant cmx_debug_set_tensor_flags cmx_debug_set_tensor_pair cmx_debug_set_tensor_small cmx_debug_set_tensor_window cmx_debug_set_block_order cmx_debug_set_speculate cmx_debug_set_small_first cmx_debug_set_mapped_outputs cmx_debug_plan_slabs"
    ).split():
        getattr(L, name).restype = i32
    _lib = L
    return L


def check(rc: int) -> None:
    """Map a non-zero C status to RuntimeError (faiss raises RuntimeError from C++ too)."""
    if rc != 0:
        msg = lib().cmx_last_error()
        raise RuntimeError(f"cmx: {msg.decode('utf-8', 'replace') if msg else 'error'} (code {rc})")


def device_count() -> int:
    n = C.c_int(0)
    rc = lib().cmx_device_count(C.byref(n))
    return n.value if rc == 0 else 0


def launch_count() -> int:
    return int(lib().cmx_launch_count())


def set_default_precision(mode) -> None:
    """'rescore' (default) or 'split' for indexes created afterwards."""
    m = {"split": PRECISION_SPLIT, "rescore": PRECISION_RESCORE}.get(mode, mode)
    check(lib().cmx_set_default_precision(int(m)))


def set_profiling(on: bool) -> None:
    check(lib().cmx_set_profiling(1 if on else 0))
