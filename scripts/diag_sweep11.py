"""torchrun diagnostic: an 11-alpha batch (76 780 queries = 10 two-phase chunks) on shards of ~1.1 M rows each."""
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
for alphas in ([0.5], bench.SWEEP11[:2], bench.SWEEP11, [0.5] * 3):
    D, I = idx.search_mixed(P, S, alphas, k)
    if rank == 0:
        print(json.dumps({"nA": len(alphas), "status": idx.last_status, "fallback_steps": idx.fallback_steps, "two_phase_used": idx.two_phase_used,
                          "local_stats": idx.local.last_stats()}), flush=True)
dist.destroy_process_group()
