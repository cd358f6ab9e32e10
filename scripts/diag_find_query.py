import json, sys, pathlib
ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "codemix-dense-retrieval_b200")]
import torch, bench
from cmx.engine import Shard, mix_normalize
from cmx.dist import shard_bounds
dev = torch.device("cuda", 0)
N, d, nq, k = bench.N_FULL, 1024, 6980, 1000
G, r, chunk = 8, 6, 6
b = shard_bounds(N, G)
P, S = bench.make_queries(nq, d, dev)
Q = mix_normalize(P, S, bench.SWEEP11).reshape(-1, d).contiguous()
sh = Shard(d, 0); sh.reserve(b[r + 1] - b[r]); bench.fill_rows(sh.add, b[r], b[r + 1], d, dev, N)
q = Q[chunk * 8192:(chunk + 1) * 8192].contiguous()
def reruns(sub):
    sh.search(sub.contiguous(), k, path="tensor"); return sh.last_stats()["reruns"]
print("whole chunk reruns", reruns(q))
# which 128-query groups overflow when searched in their original batch context?  search groups of 1024
bad = []
for g0 in range(0, 8192, 1024):
    if reruns(q[g0:g0 + 1024]):
        bad.append(g0)
print("bad 1024-groups (searched alone)", bad)
lo, hi = (bad[0], bad[0] + 1024) if bad else (0, 8192)
while hi - lo > 64 and bad:
    mid = (lo + hi) // 2
    if reruns(q[lo:mid]): hi = mid
    elif reruns(q[mid:hi]): lo = mid
    else: break
print("narrowed to", lo, hi)
X = sh.reconstruct_n(0, sh.ntotal, torch.empty((sh.ntotal, d), dtype=torch.float32, device=dev))
sub = q[lo:hi]
sc = sub @ X.T
top = torch.topk(sc, 8192, dim=1).values
print("query norms", sub.norm(dim=1).min().item(), sub.norm(dim=1).max().item())
print("1000th score min/max over group", top[:, 999].min().item(), top[:, 999].max().item())
print("8000th score min/max", top[:, 7999].min().item(), top[:, 7999].max().item())
gi = chunk * 8192 + lo
print("global flattened query index", gi, "alpha idx", gi // nq, "query", gi % nq)
