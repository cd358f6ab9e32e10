"""GPU experiment: time the stream variants and the tensor path at small batches on the
full C2 corpus; report corpus GB/s (4*N*d bytes per pass) and per-stage times."""
import json, sys, time, pathlib
ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "codemix-dense-retrieval_b200")]
import torch
from cmx import _lib
from cmx.engine import Shard
import bench

N = int(sys.argv[1]) if len(sys.argv) > 1 else bench.N_FULL
d, dev = 1024, torch.device("cuda", 0)
sh = Shard(d, 0); sh.reserve(N)
c = 0
while c * bench.CHUNK < N:
    x = bench.corpus_chunk(c, d, dev)
    sh.add(x[: min(bench.CHUNK, N - c * bench.CHUNK)]); del x; c += 1
P, S = bench.make_queries(6980, d, dev)
_lib.set_profiling(True)
gb = 4.0 * N * d / 1e9

def run(nq, k, path, reps=5, variant=None):
    if variant is not None and path == "stream":
        _lib.check(_lib.lib().cmx_debug_set_stream_variant(variant))
    q = P[:nq].contiguous()
    for _ in range(2):
        sh.search(q, k, path=path)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); sc = se = 0.0
    for _ in range(reps):
        sh.search(q, k, path=path); st = sh.last_stats(); sc += st["score_ms"]; se += st["select_ms"]
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    passes = (nq + 7) // 8 if path == "stream" else 1
    print(json.dumps({"nq": nq, "k": k, "path": path, "variant": variant, "ms": round(ms, 3), "score_ms": round(sc / reps, 3),
                      "select_ms": round(se / reps, 3), "corpus_GBps_total": round(gb * passes / (ms / 1e3), 1),
                      "corpus_GBps_score_kernel": round(gb * passes / (sc / reps / 1e3), 1), "slabs": st["slabs"]}), flush=True)

for prec in ("rescore", "split"):
    sh.set_precision(prec)
    for nq in (1, 2, 4, 8, 16, 32, 64, 128, 256):
        run(nq, 100, "tensor", variant=prec)
    run(32, 1000, "tensor", variant=prec)
sh.set_precision("rescore")
for nq in (1, 4):
    run(nq, 100, "stream")
    run(nq, 1000, "stream")
