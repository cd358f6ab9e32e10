"""How often the pool epilogue's overflow paths run in the regimes of test_early_slab_epilogue_paths_agree
(needs the -DCMX_TC_TIMERS build: CMX_LIB=.../lib/libcmx_timers.so)."""
import ctypes as C, json, sys, pathlib
ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "codemix-dense-retrieval_b200"), str(ROOT / "tests")]
import torch
from cmx import _lib
from cmx.engine import Shard
from test_gpu_round2 import _unit_cuda
L = _lib.lib()
L.cmx_debug_tc_timers.argtypes = [C.POINTER(C.c_ulonglong), C.c_int]
N, d, k, nq = 400_000, 256, 1000, 300
X = _unit_cuda(N, d, 71); Q = _unit_cuda(nq, d, 72)
buf = (C.c_ulonglong * 16)()
for precision, cap, small_first, spec in (("rescore", 0, 1, 1), ("rescore", 0, 0, 0), ("rescore", 4096, 0, 0), ("split", 2048, 0, 0)):
    sh = Shard(d, 0); sh.set_precision(precision); sh.add(X)
    _lib.check(L.cmx_debug_set_small_first(small_first)); _lib.check(L.cmx_debug_set_speculate(spec)); sh.set_cand_capacity(cap)
    sh.search(Q, k, path="tensor"); torch.cuda.synchronize()
    L.cmx_debug_tc_timers(buf, 1)
    sh.search(Q, k, path="tensor"); torch.cuda.synchronize()
    L.cmx_debug_tc_timers(buf, 1)
    print(json.dumps({"precision": precision, "cap": cap or 8192, "small_first": small_first, "speculate": spec, "slabs": sh.last_stats()["slabs"],
                      "tiles": int(buf[2]), "pool_emptied_in_place": int(buf[10]), "chunks_lane_by_lane": int(buf[11])}), flush=True)
