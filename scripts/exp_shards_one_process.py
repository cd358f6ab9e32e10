"""One process driving G GPUs (cmx.faiss.IndexShardsIP over a LocalFabric) on the C3 shape (17 683 646 x 1024):
ms per alpha-pass with HOST query vectors in and HOST (D, I) out -- what `cmx.cli ... --gpus G` runs -- to set beside
the one-process-per-GPU ShardedIndex e2e number of bench.py's extra_configs.C3.  Also an 11-alpha batch (> 8192 queries:
chunked two-phase pass)."""
import json, sys, time, pathlib
ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "codemix-dense-retrieval_b200")]
import numpy as np
import torch
import bench
import cmx.faiss as faiss

G = int(sys.argv[1]) if len(sys.argv) > 1 else torch.cuda.device_count()
rows = int(sys.argv[2]) if len(sys.argv) > 2 else 2 * bench.N_FULL
d, nq, k = 1024, 6980, 1000
idx = faiss.IndexShardsIP(d, list(range(G)))
b = faiss.IndexShardsIP.split(rows, G)
for g in range(G):
    dev = torch.device("cuda", g)
    with torch.cuda.device(dev):
        idx.shards[g].reserveMemory(b[g + 1] - b[g])
        bench.fill_rows(idx.shards[g].add, b[g], b[g + 1], d, dev, rows)
P, S = bench.make_queries(nq, d, torch.device("cuda", 0))
P, S = P.cpu().pin_memory(), S.cpu().pin_memory()
for _ in range(3):
    D, I = idx.search_mixed(P, S, [0.5], k)
steps = 6
t0 = time.perf_counter()
for _ in range(steps):
    D, I = idx.search_mixed(P, S, [0.5], k)
ms = (time.perf_counter() - t0) * 1e3 / steps
ok = bool((D[0, :, 1:] <= D[0, :, :-1]).all()) and int(I.min()) >= 0 and int(I.max()) < rows
t0 = time.perf_counter()
D11, I11 = idx.search_mixed(P, S, bench.SWEEP11, k)
ms11 = (time.perf_counter() - t0) * 1e3
same = bool(np.array_equal(I11[5], I[0]) and np.array_equal(D11[5], D[0]))
print(json.dumps({"n_gpus": G, "rows": rows, "ms_per_step_host_in_host_out": ms, "queries_per_s": nq / (ms / 1e3), "sorted_in_range": ok,
                  "two_phase_used": idx.two_phase_used, "fallback_steps": [v.fallback_steps for v in idx._views],
                  "sweep11_ms": ms11, "sweep11_alpha0.5_equals_single_alpha": same}), flush=True)
