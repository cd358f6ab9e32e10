"""Where the single-CTA scoring kernel's roles spend their cycles, slab by slab (needs the -DCMX_TC_TIMERS build:
CMX_LIB=.../lib/libcmx_timers.so).  Runs one search with a host sync after every slab (CMX_DEBUG_SLABS profiling keeps
the launches apart), reading and resetting the device counters around each scoring launch through a tiny hook:
the counters are read after whole searches restricted to a prefix of the plan is not possible, so instead the corpus
is searched with plans of 1, 2 and 3 slabs... -- simpler: run the search, then print totals; per-slab attribution
comes from running the same search on corpora that END after slab 0 / slab 1 (rows argument)."""
import ctypes as C, json, os, sys, pathlib
ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "codemix-dense-retrieval_b200")]
import torch, bench
from cmx import _lib
from cmx.engine import Shard
dev = torch.device("cuda", 0)
d, nq, k = 1024, 6980, 1000
L = _lib.lib()
L.cmx_debug_tc_timers.argtypes = [C.POINTER(C.c_ulonglong), C.c_int]
names = ["mma_wait_acc", "mma_wait_operands", "mma_tiles", "-", "epi_wait_acc", "epi_hold_acc", "-", "-", "epi_after_release", "epi_tiles"]
P, S = bench.make_queries(nq, d, dev)
for rows in [int(a) for a in sys.argv[1:]] or [1_105_228]:
    sh = Shard(d, 0); sh.reserve(rows); bench.fill_rows(sh.add, 0, rows, d, dev, rows)
    for _ in range(2): sh.search_mixed(P, S, [0.5], k)
    torch.cuda.synchronize()
    buf = (C.c_ulonglong * 16)()
    assert L.cmx_debug_tc_timers(buf, 1) == 0
    sh.search_mixed(P, S, [0.5], k); torch.cuda.synchronize()
    assert L.cmx_debug_tc_timers(buf, 1) == 0
    out = {n: int(buf[i]) for i, n in enumerate(names) if n != "-"}
    tiles = max(out["epi_tiles"], 1)
    per_tile = {n: round(v / tiles) for n, v in out.items() if n.startswith("epi_") and n != "epi_tiles"}
    mt = max(out["mma_tiles"], 1)
    print(json.dumps({"rows": rows, "totals": out, "epi_cycles_per_tile(group0,warp2)": per_tile,
                      "mma_wait_acc_per_tile": round(out["mma_wait_acc"] / mt), "mma_wait_operands_per_tile": round(out["mma_wait_operands"] / mt)}), flush=True)
    del sh
