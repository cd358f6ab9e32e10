"""1-GPU reproduction of an 8-GPU sharded 11-alpha batch: for every shard's rows (regenerated here) and every
8192-query chunk, run cmx_search_begin and print the status word; then the plain search of the same chunk (reruns)."""
import json, sys, pathlib
ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "codemix-dense-retrieval_b200")]
import torch, bench
from cmx.engine import Shard, mix_normalize
from cmx.dist import shard_bounds
G = int(sys.argv[1]) if len(sys.argv) > 1 else 8
shards = [int(v) for v in sys.argv[2].split(",")] if len(sys.argv) > 2 else list(range(G))
dev = torch.device("cuda", 0)
N, d, nq, k = bench.N_FULL, 1024, 6980, 1000
b = shard_bounds(N, G)
P, S = bench.make_queries(nq, d, dev)
Q = mix_normalize(P, S, bench.SWEEP11).reshape(-1, d).contiguous()
bnd = torch.zeros((2,), dtype=torch.float32, device=dev)
flag = torch.zeros((1,), dtype=torch.int32, device=dev)
asc = torch.empty((8192 * k,), dtype=torch.float32, device=dev)
for r in shards:
    sh = Shard(d, 0); sh.reserve(b[r + 1] - b[r]); bench.fill_rows(sh.add, b[r], b[r + 1], d, dev, N)
    sh.export_bounds(bnd.data_ptr())
    out = []
    for c0 in range(0, Q.shape[0], 8192):
        q = Q[c0:c0 + 8192].contiguous()
        sh.search_begin(q.data_ptr(), q.shape[0], k, b[r], [bnd.data_ptr()], 1.0 / G, asc.data_ptr(), flag.data_ptr())
        torch.cuda.synchronize()
        f = int(flag.item())
        sh.search(q, k, path="tensor")
        st = sh.last_stats()
        out.append((c0 // 8192, f, st["reruns"]))
    print(json.dumps({"shard": r, "rows": b[r + 1] - b[r], "chunk,flag,plain_reruns": out}), flush=True)
    del sh
