"""A/B inside one process: epilogue of the early slabs (two-pass count-then-store vs one-pass staged) and the
small dense first slab, on a shard-sized corpus.  Per-slab times come from CMX_DEBUG_SLABS (stderr)."""
import json, os, sys, pathlib, statistics
os.environ["CMX_DEBUG_SLABS"] = "1"
ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "codemix-dense-retrieval_b200")]
import torch, bench
from cmx import _lib
from cmx.engine import Shard
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1_105_228
dev = torch.device("cuda", 0)
d, nq, k = 1024, 6980, 1000
sh = Shard(d, 0); sh.reserve(rows); bench.fill_rows(sh.add, 0, rows, d, dev, rows)
P, S = bench.make_queries(nq, d, dev)
_lib.set_profiling(True)
L = _lib.lib()
variants = {"default": dict(flags=0, small=1), "one_pass": dict(flags=32, small=1), "big_first": dict(flags=0, small=0),
            "big_first_one_pass": dict(flags=32, small=0)}
rounds, block = 4, 4
res = {n: [] for n in variants}
ref = None
for rnd in range(rounds):
    for name, v in variants.items():
        _lib.check(L.cmx_debug_set_tensor_flags(v["flags"])); _lib.check(L.cmx_debug_set_small_first(v["small"]))
        print(f"## {name} round {rnd}", file=sys.stderr, flush=True)
        D, I = sh.search_mixed(P, S, [0.5], k)
        if ref is None: ref = (D.clone(), I.clone())
        assert torch.equal(D, ref[0]) and torch.equal(I, ref[1]), name
        sc = tot = 0.0
        for _ in range(block):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); sh.search_mixed(P, S, [0.5], k); e1.record(); torch.cuda.synchronize()
            sc += sh.last_stats()["score_ms"]; tot += e0.elapsed_time(e1)
        res[name].append((tot / block, sc / block))
_lib.check(L.cmx_debug_set_tensor_flags(0)); _lib.check(L.cmx_debug_set_small_first(1))
for name in variants:
    t = [a for a, _ in res[name]]; s = [b for _, b in res[name]]
    print(json.dumps({"variant": name, "rows": rows, "ms_per_step": [round(x, 2) for x in t], "median_ms": round(statistics.median(t), 2),
                      "median_score_ms": round(statistics.median(s), 2)}), flush=True)
