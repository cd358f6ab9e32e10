import json, sys, pathlib, ctypes
ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "codemix-dense-retrieval_b200")]
import torch, bench
from cmx import _lib
from cmx.engine import Shard, mix_normalize
from cmx.dist import shard_bounds
dev = torch.device("cuda", 0)
N, d, nq, k = bench.N_FULL, 1024, 6980, 1000
G, r, chunk = 8, 6, 6
b = shard_bounds(N, G)
P, S = bench.make_queries(nq, d, dev)
Q = mix_normalize(P, S, bench.SWEEP11).reshape(-1, d).contiguous()
sh = Shard(d, 0); sh.reserve(b[r + 1] - b[r]); bench.fill_rows(sh.add, b[r], b[r + 1], d, dev, N)
n = sh.ntotal
q = Q[chunk * 8192 + 2368: chunk * 8192 + 2432].contiguous()
X = sh.reconstruct_n(0, n, torch.empty((n, d), dtype=torch.float32, device=dev))
nblk = (n + 255) // 256
Pm = int(_lib.lib().cmx_debug_block_perm(nblk))
pos_blocks = (torch.arange(nblk, device=dev, dtype=torch.int64) * Pm) % nblk
def rows_of(b0, b1):
    blks = pos_blocks[b0:b1]
    rows = (blks[:, None] * 256 + torch.arange(256, device=dev)[None, :]).reshape(-1)
    return rows[rows < n]
# margins from the library
sc = torch.empty((64, 8192), dtype=torch.float32, device=dev); rw = torch.empty((64, 8192), dtype=torch.int64, device=dev); mg = torch.empty((64,), dtype=torch.float32, device=dev)
_lib.check(_lib.lib().cmx_debug_approx_scores(sh._h, q.data_ptr(), 64, 0, 8192, sc.data_ptr(), rw.data_ptr(), mg.data_ptr(), None))
print("margin min/max", mg.min().item(), mg.max().item())
Xh = X.half().float()  # approx scores ~ fp16-rounded operands (scale-free approximation)
qh = q.half().float()
s0 = qh @ Xh[rows_of(0, 32)].T
s1 = qh @ Xh[rows_of(32, 113)].T
s2 = qh @ Xh[rows_of(113, nblk)].T
kth0 = torch.topk(s0, k, dim=1).values[:, -1]
tau0 = kth0 - mg
keep0 = (s0 >= tau0[:, None]).sum(1)
surv1 = (s1 > tau0[:, None]).sum(1)
u = torch.cat([s0, s1], dim=1)
tk = torch.topk(u, k, dim=1).values
kth1 = tk[:, -1]; tau1 = kth1 - mg
keep1 = (u >= tau1[:, None]).sum(1)
spec = torch.topk(u, 106, dim=1).values[:, -1] - mg
surv2 = (s2 > spec[:, None]).sum(1)
print("keep0", keep0.min().item(), keep0.max().item(), "surv1 max", surv1.max().item(), "keep1 max", keep1.max().item())
print("surv2 min/mean/max", surv2.min().item(), surv2.float().mean().item(), surv2.max().item(), "total max", (keep1 + surv2).max().item())
i = int(torch.argmax(keep1 + surv2))
print("worst query", i, "keep1", keep1[i].item(), "surv2", surv2[i].item(), "spec", spec[i].item(), "kth1", kth1[i].item(), "tau0", tau0[i].item())
ranks_ = torch.topk(u[i], 200).values
print("its sample top scores (1,10,50,106,200):", ranks_[0].item(), ranks_[9].item(), ranks_[49].item(), ranks_[105].item(), ranks_[199].item())
