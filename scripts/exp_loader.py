"""Load path (1 GPU): index.faiss -> HBM.  Native loader (reader threads + pinned staging, cmx_index_add_from_file)
vs the round-1 path (np.fromfile chunks -> pageable -> cudaMemcpy), on /dev/shm (page-cache speed) and on the box's disk."""
import json, os, sys, time, pathlib
ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "codemix-dense-retrieval_b200")]
import numpy as np
import torch
import cmx.faiss as faiss
from cmx import io as cio

rows, d = int(os.environ.get("ROWS", 2_000_000)), 1024
x = torch.nn.functional.normalize(torch.randn((rows, d), device="cuda"), dim=1).cpu().numpy()
for base in ("/dev/shm", "/tmp"):
    path = pathlib.Path(base) / "cmx_loader_test.faiss"
    try:
        idx = faiss.IndexIDMap(faiss.IndexFlatIP(d))
        idx.add_with_ids(x, np.arange(rows, dtype=np.int64))
        t0 = time.perf_counter(); faiss.write_index(idx, path); tw = time.perf_counter() - t0
        del idx
        gb = 4e-9 * rows * d
        if base == "/tmp":  # drop the page cache if we may, so the second pass reads the disk
            os.system("sync; echo 3 > /proc/sys/vm/drop_caches 2>/dev/null")
        for label in ("native", "native_again", "numpy_chunks", "host_index_to_gpu"):
            t0 = time.perf_counter()
            if label.startswith("native"):
                g = faiss.read_index_to_gpu(str(path), 0)
                extra = dict(cio.LAST_LOAD)
            elif label == "host_index_to_gpu":  # the reference's own flow: read_index (host) + index_cpu_to_gpu
                cpu = faiss.read_index(str(path))
                t_read = time.perf_counter() - t0
                g = faiss.index_cpu_to_gpu(faiss.StandardGpuResources(), 0, cpu)
                extra = {"read_index_s": round(t_read, 3)}
                del cpu
            else:
                info = cio.inspect_index(path)
                flat = faiss.GpuIndexFlatIP(d, device=0); flat.reserveMemory(rows)
                with open(path, "rb") as fh:
                    fh.seek(info["vec_offset"])
                    r = 0
                    while r < rows:
                        n = min(1 << 16, rows - r)
                        flat.add(np.fromfile(fh, dtype="<f4", count=n * d).reshape(n, d)); r += n
                g, extra = flat, {}
            torch.cuda.synchronize(); dt = time.perf_counter() - t0
            ok = bool(np.array_equal((g.index if hasattr(g, "index") else g).reconstruct_n(rows - 5, 5), x[-5:]))
            print(json.dumps({"where": base, "path": label, "GB": round(gb, 2), "seconds": round(dt, 3), "GB_per_s": round(gb / dt, 2),
                              "ok": ok, "write_s": round(tw, 2), **{k: (round(v, 3) if isinstance(v, float) else v) for k, v in extra.items()}}), flush=True)
            del g
    finally:
        if path.exists():
            path.unlink()
