"""GPU experiment: the COMPLETE per-alpha loop of the mono run script at the full C2 shape --
fused mix+search on the device, D2H of (D, I), TREC text formatting and file write -- i.e.
what replaces onepass_dense_mix_run_custom_lang.py:844-890 for one job (7 alphas, k=100 as the
reference runs it, and k=1000 as BASELINE asks)."""
import json, sys, time, pathlib, tempfile
ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "codemix-dense-retrieval_b200")]
import numpy as np
import torch
import cmx.faiss as faiss
from cmx import runloop
import bench

N = int(sys.argv[1]) if len(sys.argv) > 1 else bench.N_FULL
d, dev = 1024, torch.device("cuda", 0)
flat = faiss.GpuIndexFlatIP(d)
flat.reserveMemory(N)
c = 0
while c * bench.CHUNK < N:
    x = bench.corpus_chunk(c, d, dev)
    flat.add(x[: min(bench.CHUNK, N - c * bench.CHUNK)]); del x; c += 1
index = faiss.IndexIDMap(faiss.GpuIndexFlatIP(d))
index.index = flat
index._ids = [np.arange(N, dtype=np.int64)]
P, S = bench.make_queries(6980, d, dev)
P_h, S_h = P.cpu().numpy(), S.cpu().numpy()
qids = [str(1048585 + 7 * i) for i in range(6980)]
t0 = time.perf_counter()
docs = runloop.DocTable([str(i) for i in range(N)])
t_doc = time.perf_counter() - t0
alphas = [0, 0.1, 0.3, 0.5, 0.7, 0.9, 1]
for k in (100, 1000):
    with tempfile.TemporaryDirectory() as td:
        runloop.run_alpha_sweep(index, docs, qids, P_h, S_h, [0.5], td, k=k)  # warm-up (planes, workspaces)
        t0 = time.perf_counter()
        files = runloop.run_alpha_sweep(index, docs, qids, P_h, S_h, alphas, td, k=k)
        dt = time.perf_counter() - t0
        size = sum(f.stat().st_size for f in files)
        # split of one alpha
        t1 = time.perf_counter(); D, I = index.search_mixed(P_h, S_h, [0.5], k); t2 = time.perf_counter()
        b = runloop.mono_trec_bytes(qids, D[0], I[0], docs); t3 = time.perf_counter()
        (pathlib.Path(td) / "x.trec").write_bytes(b); t4 = time.perf_counter()
    print(json.dumps({"rows": N, "k": k, "alphas": len(alphas), "sweep_s": round(dt, 3), "s_per_alpha": round(dt / len(alphas), 3),
                      "files_MB": round(size / 1e6, 1), "doc_table_build_s": round(t_doc, 2),
                      "one_alpha": {"search_host_io_s": round(t2 - t1, 3), "format_s": round(t3 - t2, 3), "write_s": round(t4 - t3, 3)}}), flush=True)
