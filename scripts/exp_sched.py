"""A/B inside one process (same box, same clocks): dynamic tile scheduler vs static round-robin + throttle, alternating."""
import json, sys, pathlib, statistics
ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "codemix-dense-retrieval_b200")]
import torch, bench
from cmx import _lib
from cmx.engine import Shard
rows = int(sys.argv[1]) if len(sys.argv) > 1 else bench.N_FULL
dev = torch.device("cuda", 0)
d, nq, k = 1024, 6980, 1000
sh = Shard(d, 0); sh.reserve(rows); bench.fill_rows(sh.add, 0, rows, d, dev, rows)
P, S = bench.make_queries(nq, d, dev)
_lib.set_profiling(True)
L = _lib.lib()
res = {512: [], 0: []}
for rnd in range(5):
    for flag in (512, 0):
        _lib.check(L.cmx_debug_set_tensor_flags(flag))
        sh.search_mixed(P, S, [0.5], k)
        sc = tot = 0.0
        for _ in range(4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); sh.search_mixed(P, S, [0.5], k); e1.record(); torch.cuda.synchronize()
            sc += sh.last_stats()["score_ms"]; tot += e0.elapsed_time(e1)
        res[flag].append((tot / 4, sc / 4))
_lib.check(L.cmx_debug_set_tensor_flags(0))
for flag, name in ((512, "dynamic"), (0, "static+throttle")):
    t = [a for a, _ in res[flag]]; s = [b for _, b in res[flag]]
    print(json.dumps({"scheduler": name, "rows": rows, "ms_per_step": [round(v, 2) for v in t], "score_ms": [round(v, 2) for v in s],
                      "median_ms": round(statistics.median(t), 2), "median_score_ms": round(statistics.median(s), 2)}), flush=True)
