"""GPU experiment: tensor-path scoring time vs L2 cache-hint flags / tile width / corpus size."""
import json, sys, pathlib
ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "codemix-dense-retrieval_b200")]
import torch
from cmx import _lib
from cmx.engine import Shard
import bench

sizes = [int(v) for v in sys.argv[1].split(",")] if len(sys.argv) > 1 else [1_000_000, 4_420_912]
d, dev = 1024, torch.device("cuda", 0)
P, S = bench.make_queries(6980, d, dev)
_lib.set_profiling(True)
for N in sizes:
    sh = Shard(d, 0); sh.reserve(N)
    c = 0
    while c * bench.CHUNK < N:
        x = bench.corpus_chunk(c, d, dev)
        sh.add(x[: min(bench.CHUNK, N - c * bench.CHUNK)]); del x; c += 1
    for prec, pair, flags, bn in (("rescore", 0, 3, 256), ("rescore", 0, 2, 256), ("rescore", 0, 4, 256), ("rescore", 0, 6, 256),
                                  ("rescore", 0, 0, 256), ("rescore", 1, 3, 256), ("rescore", 1, 2, 256), ("rescore", 0, 3, 128),
                                  ("rescore", 0, 3, 256)):
        sh.set_precision(prec)
        _lib.check(_lib.lib().cmx_debug_set_tensor_pair(pair))
        _lib.check(_lib.lib().cmx_debug_set_tensor_tile(bn))
        for _ in (0,):
            _lib.check(_lib.lib().cmx_debug_set_tensor_window(flags))
            for _ in range(2):
                sh.search_mixed(P, S, [0.5], 1000, path="tensor")
            sc = se = tot = 0.0
            reps = 4
            for _ in range(reps):
                sh.search_mixed(P, S, [0.5], 1000, path="tensor"); st = sh.last_stats()
                sc += st["score_ms"]; se += st["select_ms"]; tot += st["total_ms"]
            tf = (3 if prec == "split" else 1) * 2.0 * 6980 * N * d / (sc / reps / 1e3) / 1e12
            print(json.dumps({"N": N, "precision": prec, "pair": pair, "window": flags, "bn": bn, "reruns": st["reruns"], "qps": round(6980 / (tot / reps / 1e3)), "score_ms": round(sc / reps, 2), "select_ms": round(se / reps, 2),
                              "total_ms": round(tot / reps, 2), "exec_TFLOPs": round(tf, 1)}), flush=True)
    del sh
    torch.cuda.empty_cache()
