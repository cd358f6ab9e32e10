"""torchrun correctness check (N GPUs, one process each): sharded search over peer memory (two-phase cut,
fused merge, host results in ONE shared pinned buffer, query upload split over the ranks) == NCCL
all_gather + merge == single-GPU search of the whole corpus, bit for bit.  Also: a batch larger than one
two-phase chunk, an index with EMPTY shards (fewer rows than ranks) and a rank whose shard cannot run the
one-pass arithmetic (everybody falls back together instead of deadlocking)."""
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
P_h, S_h = P.cpu().pin_memory(), S.cpu().pin_memory()
res, info = {}, {}
checks = {}
for mode, prec in (("allgather", "rescore"), ("p2p", "rescore"), ("allgather", "split"), ("p2p", "split")):
    idx = ShardedIndex(d, N, device=lr, exchange=mode)
    idx.set_precision(prec)
    idx.add_local(X[idx.row0:idx.row1].contiguous())
    D, I = idx.search_mixed(P, S, [0.0, 0.5], k)
    D2, I2 = idx.search(P[:5].contiguous(), k)
    res[(mode, prec)] = (D.clone(), I.clone(), D2.clone(), I2.clone())
    info[(mode, prec)] = (idx.exchange_used, idx.two_phase_used, idx.fallback_steps)
    if mode == "p2p":
        out = idx.host_output(2, nq, k)
        Dh, Ih = idx.search_mixed_host(P_h, S_h, [0.0, 0.5], k, out=out)
        checks[f"host_result_{prec}"] = bool(torch.equal(Dh, D.cpu()) and torch.equal(Ih, I.cpu()))
        if prec == "rescore":
            many = [i / 11.0 for i in range(12)]  # 12 x 777 = 9324 queries > 8192: two chunks
            Dm, Im = idx.search_mixed(P, S, many, 50)
            res["many"] = (Dm.clone(), Im.clone())
            info["many"] = (idx.two_phase_used, idx.fallback_steps)
            # a candidate-buffer overflow reported by ONE rank in the SECOND chunk: only that chunk is redone, by everybody
            from cmx import _lib
            if rank == world - 1:
                _lib.check(_lib.lib().cmx_debug_inject_begin_status(1, 1))
            fc0 = idx.fallback_chunks
            Dj, Ij = idx.search_mixed(P, S, many, 50)
            checks["chunk_fallback"] = bool(idx.fallback_chunks == fc0 + 1 and idx.fallback_steps == 0 and idx.last_status == 1
                                            and torch.equal(Dj, res["many"][0]) and torch.equal(Ij, res["many"][1]))
    del idx
ok = True
ok1 = True
for prec in ("rescore", "split"):
    ok = ok and all(torch.equal(a, b) for a, b in zip(res[("allgather", prec)], res[("p2p", prec)]))
    single = Shard(d, lr)
    single.set_precision(prec)
    single.add(X)
    Ds, Is = single.search_mixed(P, S, [0.0, 0.5], k)
    ok1 = ok1 and torch.equal(Ds, res[("p2p", prec)][0]) and torch.equal(Is, res[("p2p", prec)][1])
    if prec == "rescore":
        Dm, Im = single.search_mixed(P, S, [i / 11.0 for i in range(12)], 50)
        checks["chunked_batch"] = bool(torch.equal(Dm, res["many"][0]) and torch.equal(Im, res["many"][1]))
    del single
checks["p2p==allgather"] = ok
checks["sharded==single"] = ok1
checks["two_phase_used"] = bool(info[("p2p", "rescore")][1]) and info[("p2p", "rescore")][2] == 0

# fewer rows than ranks: most shards are empty
tiny = ShardedIndex(d, 3, device=lr, exchange="p2p")
tiny.add_local(X[tiny.row0:tiny.row1].contiguous()) if tiny.row1 > tiny.row0 else None
Dt, It = tiny.search_mixed(P, S, [0.5], 5)
one = Shard(d, lr)
one.add(X[:3].contiguous())
D1, I1 = one.search_mixed(P, S, [0.5], 5)
checks["empty_shards"] = bool(torch.equal(Dt, D1) and torch.equal(It, I1))
del tiny, one

# one rank in split precision (a shard-local reason not to run two-phase): the status words make every rank fall back
odd = ShardedIndex(d, N, device=lr, exchange="p2p")
odd.add_local(X[odd.row0:odd.row1].contiguous())
if rank == world - 1:
    odd.local.set_precision("split")
Do, Io = odd.search_mixed(P, S, [0.0, 0.5], k)
checks["collective_fallback"] = bool(odd.fallback_steps == 1 and torch.allclose(Do, res[("p2p", "rescore")][0], rtol=1e-5, atol=1e-6)
                                     and float((Io == res[("p2p", "rescore")][1]).float().mean()) > 0.999)
del odd

good = all(checks.values())
flag = torch.tensor([int(good)], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("check_dist:", "OK" if int(flag) == 1 else "MISMATCH", checks, "modes:", {str(k2): v for k2, v in info.items()}, flush=True)
else:
    if not good:
        print(f"check_dist rank {rank}:", checks, flush=True)
dist.destroy_process_group()
sys.exit(0 if int(flag) == 1 else 1)
