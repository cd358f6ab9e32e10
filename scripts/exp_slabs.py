"""per-slab timing of one search_mixed at several corpus sizes (CMX_DEBUG_SLABS=1)"""
import os, sys, pathlib
os.environ["CMX_DEBUG_SLABS"] = "1"
ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "codemix-dense-retrieval_b200")]
import torch
from cmx import _lib
from cmx.engine import Shard
import bench
d, dev = 1024, torch.device("cuda", 0)
P, S = bench.make_queries(6980, d, dev)
_lib.set_profiling(True)
for N in [int(v) for v in sys.argv[1].split(",")]:
    sh = Shard(d, 0); sh.reserve(N)
    c = 0
    while c * bench.CHUNK < N:
        x = bench.corpus_chunk(c, d, dev)
        sh.add(x[: min(bench.CHUNK, N - c * bench.CHUNK)]); del x; c += 1
    for prec, flags in (("rescore", 0), ("rescore", 16), ("rescore", 32)):
        sh.set_precision(prec)
        _lib.check(_lib.lib().cmx_debug_set_tensor_flags(flags))
        for _ in range(2):
            sys.stderr.write(f"--- N={N} {prec} flags={flags}\n"); sys.stderr.flush()
            try:
                sh.search_mixed(P, S, [0.5], 1000)
            except Exception as e:
                sys.stderr.write(f"error {e}\n")
        st = sh.last_stats()
        print(N, prec, flags, st, flush=True)
    _lib.check(_lib.lib().cmx_debug_set_tensor_flags(0))
    del sh; torch.cuda.empty_cache()
