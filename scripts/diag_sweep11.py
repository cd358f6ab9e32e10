"""torchrun diagnostic: an 11-alpha batch (76 780 queries = 10 two-phase chunks); prints every shard's status word per chunk."""
import os, sys, pathlib, json
ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "codemix-dense-retrieval_b200")]
import torch, torch.distributed as dist
import bench
from cmx.dist import ShardedIndex
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1_105_228 * world
d, nq, k = 1024, 6980, 1000
idx = ShardedIndex(d, rows, device=lr, exchange="p2p")
bench.fill_shard(idx, d, dev, rows)
P, S = bench.make_queries(nq, d, dev)
orig = idx.local.search_begin
log = []
def traced(q_ptr, nqc, k_, id_base, bounds, est, scores_ptr, flag_ptr):
    orig(q_ptr, nqc, k_, id_base, bounds, est, scores_ptr, flag_ptr)
    torch.cuda.synchronize()
    b = next(iter(idx._bufs.values()))
    f = int(b["flags"][0][0].item())
    cnt = None
    log.append((len(log), nqc, f))
idx.local.search_begin = traced
for alphas in ([0.5], bench.SWEEP11):
    log.clear()
    D, I = idx.search_mixed(P, S, alphas, k)
    print(json.dumps({"rank": rank, "nA": len(alphas), "status": idx.last_status, "fallback_steps": idx.fallback_steps,
                      "chunks(nq,flag)": [(n, f) for _, n, f in log]}), flush=True)
# the overflowing chunk alone, through the plain single-shard search: does the shard itself rerun?
from cmx.engine import mix_normalize
Q = mix_normalize(P, S, bench.SWEEP11).reshape(-1, d)
for c0 in range(0, Q.shape[0], 8192):
    idx.local.search(Q[c0:c0 + 8192].contiguous(), k, path="tensor")
    st = idx.local.last_stats()
    if st["reruns"]:
        print(json.dumps({"rank": rank, "chunk_start": c0, "plain_search_reruns": st["reruns"], "slabs": st["slabs"]}), flush=True)
dist.destroy_process_group()
