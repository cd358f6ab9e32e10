"""A/B inside one process: tensor-kernel flag sets (cmx_debug_set_tensor_flags), alternating.
    python scripts/exp_flags.py <rows> <name=flags,...> [rounds] [block]"""
import json, sys, pathlib, statistics, random
ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "codemix-dense-retrieval_b200")]
import torch, bench
from cmx import _lib
from cmx.engine import Shard
rows = int(sys.argv[1])
variants = {kv.split("=")[0]: int(kv.split("=")[1]) for kv in sys.argv[2].split(",")}
rounds = int(sys.argv[3]) if len(sys.argv) > 3 else 5
block = int(sys.argv[4]) if len(sys.argv) > 4 else 4
dev = torch.device("cuda", 0)
d, nq, k = 1024, 6980, 1000
sh = Shard(d, 0); sh.reserve(rows); bench.fill_rows(sh.add, 0, rows, d, dev, rows)
P, S = bench.make_queries(nq, d, dev)
_lib.set_profiling(True)
L = _lib.lib()
random.seed(1)
res = {n: [] for n in variants}
ref = None
for rnd in range(rounds):
    order = list(variants.items()); random.shuffle(order)
    for name, fl in order:
        _lib.check(L.cmx_debug_set_tensor_flags(fl))
        D, I = sh.search_mixed(P, S, [0.5], k)
        if ref is None: ref = (D.clone(), I.clone())
        assert torch.equal(D, ref[0]) and torch.equal(I, ref[1]), name
        sc = tot = 0.0
        for _ in range(block):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); sh.search_mixed(P, S, [0.5], k); e1.record(); torch.cuda.synchronize()
            sc += sh.last_stats()["score_ms"]; tot += e0.elapsed_time(e1)
        res[name].append((tot / block, sc / block))
_lib.check(L.cmx_debug_set_tensor_flags(0))
for name in variants:
    t = [a for a, _ in res[name]]; s = [b for _, b in res[name]]
    print(json.dumps({"variant": name, "flags": variants[name], "rows": rows, "ms_per_step": [round(x, 2) for x in t],
                      "median_ms": round(statistics.median(t), 2), "median_score_ms": round(statistics.median(s), 2)}), flush=True)
