"""torchrun correctness check (N GPUs): sharded search with the fused peer-memory merge ==
sharded search with NCCL all_gather + merge == single-GPU search of the whole corpus."""
import os, sys, pathlib
ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "codemix-dense-retrieval_b200")]
import torch
import torch.distributed as dist
from cmx.dist import ShardedIndex
from cmx.engine import Shard

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
N, d, nq, k = 300_007, 256, 777, 200
g = torch.Generator(device=dev).manual_seed(5)
X = torch.nn.functional.normalize(torch.randn((N, d), generator=g, device=dev), dim=1)
X[200_000:200_500] = X[100:600]  # ties across shards
P = torch.nn.functional.normalize(torch.randn((nq, d), generator=g, device=dev), dim=1)
S = torch.nn.functional.normalize(torch.randn((nq, d), generator=g, device=dev), dim=1)
res = {}
for mode, prec in (("allgather", "rescore"), ("p2p", "rescore"), ("allgather", "split"), ("p2p", "split")):
    idx = ShardedIndex(d, N, device=lr, exchange=mode)
    idx.set_precision(prec)
    idx.add_local(X[idx.row0:idx.row1].contiguous())
    D, I = idx.search_mixed(P, S, [0.0, 0.5], k)
    D2, I2 = idx.search(P[:5].contiguous(), k)
    res[(mode, prec)] = (D.clone(), I.clone(), D2.clone(), I2.clone(), idx.exchange_used, idx.two_phase_used)
    del idx
ok = True
ok1 = True
for prec in ("rescore", "split"):
    ok = ok and all(torch.equal(a, b) for a, b in zip(res[("allgather", prec)][:4], res[("p2p", prec)][:4]))
    single = Shard(d, lr)
    single.set_precision(prec)
    single.add(X)
    Ds, Is = single.search_mixed(P, S, [0.0, 0.5], k)
    ok1 = ok1 and torch.equal(Ds, res[("p2p", prec)][0]) and torch.equal(Is, res[("p2p", prec)][1])
    del single
flag = torch.tensor([int(ok and ok1)], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("check_dist:", "OK" if int(flag) == 1 else "MISMATCH", "modes used:", {k: v[4:] for k, v in res.items()},
          "p2p==allgather", ok, "sharded==single", ok1, flush=True)
dist.destroy_process_group()
sys.exit(0 if int(flag) == 1 else 1)
