"""Experiment (1 GPU, C2 shape): what the last slab costs with / without prescoring beside it.
mode 0 = one scoring launch + all exact scores afterwards; 2 = the last slab cut into launches, no prescoring;
1 = prescoring on the second stream (depth = est rank / expected rank of the final k-th best, pad = dynamic smem per
prescore CTA -> how many of them share an SM with the scoring CTA, max_sub = scoring launches of the last slab)."""
import json, os, sys, pathlib
ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "codemix-dense-retrieval_b200")]
import torch
import bench
from cmx import _lib
from cmx.engine import Shard

dev = torch.device("cuda", 0)
N, d, nq, k = bench.N_FULL, 1024, 6980, 1000
sh = Shard(d, 0)
sh.reserve(N)
bench.fill_rows(sh.add, 0, N, d, dev, N)
P, S = bench.make_queries(nq, d, dev)
_lib.set_profiling(True)
L = _lib.lib()
configs = [(0, 1.5, 0, 8), (2, 1.5, 0, 8), (1, 1.5, 0, 8), (1, 1.0, 0, 8), (1, 1.0, 12288, 8), (1, 0.7, 12288, 8),
           (1, 1.0, 20480, 8), (1, 1.0, 12288, 4), (1, 1.0, 12288, 2), (0, 1.5, 0, 8)]
if len(sys.argv) > 1:
    configs = [tuple(float(v) if i == 1 else int(v) for i, v in enumerate(c.split(","))) for c in sys.argv[1:]]
ref = None
_lib.check(L.cmx_debug_set_tensor_flags(512))  # prescoring is measured with the dynamic tile scheduler (see profiles/r02_prescore_experiments.md)
for mode, depth, pad, msub in configs:
    _lib.check(L.cmx_debug_set_prescore(mode))
    _lib.check(L.cmx_debug_set_prescore_params(depth, pad, msub))
    for _ in range(2):
        D, I = sh.search_mixed(P, S, [0.5], k)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sc = se = 0.0
    e0.record()
    steps = 4
    for _ in range(steps):
        D, I = sh.search_mixed(P, S, [0.5], k)
        st = sh.last_stats()
        sc += st["score_ms"]; se += st["select_ms"]
    e1.record(); torch.cuda.synchronize()
    if ref is None:
        ref = (D.clone(), I.clone())
    same = bool(torch.equal(D, ref[0]) and torch.equal(I, ref[1]))
    print(json.dumps({"mode": mode, "depth": depth, "pad": pad, "max_sub": msub, "ms_per_step": e0.elapsed_time(e1) / steps,
                      "score_ms": sc / steps, "select_ms": se / steps, "slabs": st["slabs"], "launches": st["score_launches"],
                      "reruns": st["reruns"], "same_result": same}), flush=True)
