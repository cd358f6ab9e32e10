#!/usr/bin/env python
"""Benchmark of the hot path: vector-mix query step + flat inner-product top-k.

    python bench.py --gpus N --steps K --warmup W            # the CUDA path (this repo)
    python bench.py --impl reference --gpus N --steps K ...  # the CPU path (oracle port)

A "step" = one alpha-pass of the hot path: mix+normalise nq cached query pairs at
alpha=0.5 and search the whole corpus for the top k (BASELINE.json configs[1]:
8 841 823 x 1024 fp32, 6980 queries, k=1000).  At N > 1 GPUs the SAME corpus is
row-sharded over the ranks (strong scaling): global k-th cut + fused peer-memory merge.
Prints ONE JSON line on rank 0.  Synthetic data: corpus rows
normalize(N(0,I)) generated on device in 2^20-row chunks seeded 1234+chunk, queries
P seed 42, S = normalize(0.8 P + 0.6 normalize(N(0,I))) seed 43.

After the timed headline the line also carries (all untimed w.r.t. the headline):
  parity_check   64 sampled queries of the timed result against an fp32 brute force
                 (torch.matmul, TF32 off) over the corpus regenerated from its seeds;
                 at N > 1 also bit-equality with a single-GPU search of the same queries
  extra_configs  the other BASELINE configs: C4 small batches nq in {1,16,32} with an
                 HBM roofline block, the 11-alpha sweep, C1 / C5 subsets, and C3
                 (17 683 646 rows) -- sharded when N >= 2
"""
from __future__ import annotations

import argparse
import json
import os
import pathlib
import statistics
import subprocess
import sys
import time

ROOT = pathlib.Path(__file__).resolve().parent
for p in (str(ROOT), str(ROOT / "codemix-dense-retrieval_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = "queries/sec @k=1000 over 8.8M x 1024 flat-IP"
UNIT = "queries/s"
N_FULL, D_FULL, NQ_FULL, K_FULL, ALPHA = 8_841_823, 1024, 6980, 1000, 0.5
CHUNK = 1 << 20
SWEEP11 = [round(0.1 * i, 1) for i in range(11)]


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cmx", choices=["cmx", "reference"])
    # shrink knobs for development only; the default is the BASELINE configuration
    ap.add_argument("--rows", type=int, default=N_FULL)
    ap.add_argument("--dim", type=int, default=D_FULL)
    ap.add_argument("--nq", type=int, default=NQ_FULL)
    ap.add_argument("--k", type=int, default=K_FULL)
    ap.add_argument("--path", default="auto", choices=["auto", "stream", "tensor"])
    ap.add_argument("--precision", default="rescore", choices=["rescore", "split"],
                    help="tensor-path arithmetic: one fp16 MMA pass + exact fp32 rescoring (default) or 3-pass split precision")
    ap.add_argument("--exchange", default="auto", choices=["auto", "p2p", "allgather"],
                    help="multi-GPU merge: fused peer-memory kernel (p2p) or NCCL all_gather + merge")
    ap.add_argument("--data", default="iid", choices=["iid", "aniso", "shift"],
                    help="synthetic corpus: iid = normalize(N(0,I)) (default, SURVEY 8d); aniso = normalize(g + 1.5 sqrt(d) u), "
                         "one shared direction u, scores ~0.7 like real embedding cosines (SURVEY 8d variant); shift = iid "
                         "first half, second half and all queries leaning towards one direction (EN rows then ZH rows of "
                         "the bilingual index: not stationary in file order)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip extra_configs (C1/C3/C4/C5)")
    ap.add_argument("--no-parity", action="store_true", help="skip the sampled brute-force parity check")
    ap.add_argument("--stage-times", action="store_true", help="after the timed runs, print a per-stage breakdown (extra syncs)")
    ap.add_argument("--cpu-sample-queries", type=int, default=128,
                    help="queries per step of the CPU arm / cpu_baseline (full corpus rows when host RAM allows)")
    return ap.parse_args()


def workload_name(rows, dim, nq, k, data="iid") -> str:
    shape = (rows, dim, nq, k)
    if shape == (N_FULL, D_FULL, NQ_FULL, K_FULL):
        base = "C2 (BASELINE configs[1]): EN monolingual full mMARCO shape"
    elif shape == (2 * N_FULL, D_FULL, NQ_FULL, K_FULL):
        base = "C3 (BASELINE configs[2]): EN+ZH bilingual combined index shape"
    else:
        base = "other shape (not the headline configuration)"
    var = "" if data == "iid" else f", data variant {data}"
    return f"{base}: {rows} x {dim} fp32 flat-IP, {nq} queries, alpha={ALPHA}, k={k}{var}"


def workload_config(a, world: int) -> dict:
    """`config` of BOTH arms (the reference arm is timed on this arm's config)."""
    return {"workload": workload_name(a.rows, a.dim, a.nq, a.k, a.data), "rows": a.rows, "dim": a.dim, "queries": a.nq,
            "k": a.k, "alpha": ALPHA,
            "parallelism": f"corpus row shards x{world}" if world > 1 else "single GPU",
            "cache": "inputs_larger_than_L2 (corpus %.1f GB per GPU)" % (a.rows * a.dim * 4 / 1e9 / max(1, world))}


DATA = "iid"      # set from --data
ROWS_TOTAL = N_FULL


def _direction(d: int, device):
    import torch

    g = torch.Generator(device="cpu").manual_seed(7)
    u = torch.nn.functional.normalize(torch.randn((1, d), generator=g), dim=1)
    return u.to(device)


def make_queries(nq: int, d: int, device):
    import torch

    g1 = torch.Generator(device=device).manual_seed(42)
    g2 = torch.Generator(device=device).manual_seed(43)
    P = torch.randn((nq, d), generator=g1, device=device)
    if DATA == "aniso":
        P = P + 1.5 * (d ** 0.5) * _direction(d, device)
    elif DATA == "shift":
        P = P + 0.5 * (d ** 0.5) * _direction(d, device)
    P = torch.nn.functional.normalize(P, dim=1)
    G = torch.nn.functional.normalize(torch.randn((nq, d), generator=g2, device=device), dim=1)
    S = torch.nn.functional.normalize(0.8 * P + 0.6 * G, dim=1)
    return P.contiguous(), S.contiguous()


def corpus_chunk(c: int, d: int, device, rows_total=None):
    """Global chunk c (rows [c*2^20, (c+1)*2^20)), identical for every GPU count."""
    import torch

    g = torch.Generator(device=device).manual_seed(1234 + c)
    x = torch.randn((CHUNK, d), generator=g, device=device)
    if DATA == "aniso":
        x += 1.5 * (d ** 0.5) * _direction(d, device)
    elif DATA == "shift":  # rows of the second half of the corpus lean towards u
        rows = torch.arange(c * CHUNK, (c + 1) * CHUNK, device=device)
        x += (rows >= (rows_total or ROWS_TOTAL) // 2).to(x.dtype)[:, None] * (0.5 * (d ** 0.5)) * _direction(d, device)
    return torch.nn.functional.normalize(x, dim=1)


def fill_rows(add, row0: int, row1: int, d: int, device, rows_total=None) -> None:
    c = row0 // CHUNK
    while c * CHUNK < row1:
        x = corpus_chunk(c, d, device, rows_total)
        lo = max(row0, c * CHUNK) - c * CHUNK
        hi = min(row1, (c + 1) * CHUNK) - c * CHUNK
        add(x[lo:hi])
        del x
        c += 1


def fill_shard(index, d: int, device, rows_total=None) -> None:
    index.reserve_local()
    fill_rows(index.add_local, index.row0, index.row1, d, device, rows_total)
    assert index.local_complete()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50", "-i", str(self.gpu)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, smax, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.splitlines():
            f = [v.strip() for v in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        # "under load" = samples drawing more than half of the peak power seen
        loaded = [s for s, p in zip(sm, power) if p >= 0.5 * max(power)] or sm
        return {"sm_mhz": statistics.median(loaded), "sm_max_mhz": max(smax), "power_w_max": max(power),
                "samples": len(sm), "reasons": sorted(reasons)}


def ncu_traffic(kernel: str, rows: int, d: int):
    """DRAM bytes per step for the dominant kernel, from the committed ncu --set full capture:
    measured bytes per corpus row (profiles/ncu_traffic.json, captured at d=1024) x rows."""
    p = ROOT / "profiles" / "ncu_traffic.json"
    try:
        j = json.loads(p.read_text())[kernel]
        return j["dram_bytes_per_corpus_row"] * (d / 1024.0) * rows, j["dram_bytes_per_corpus_row"]
    except Exception:
        return None, None


def measured_peaks() -> dict:
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            j = json.loads(p.read_text())
            return {"bf16_tflops": j.get("bf16_tflops"), "bf16_tflops_sustained": j.get("bf16_tflops_sustained"),
                    "hbm_gbs": j.get("hbm_gbs"), "source": "measured (MEASURED_PEAKS.json)"}
        except Exception:
            pass
    return {"bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "hbm_gbs": 6650.0,
            "source": "fallback (B200_PROFILING.md)"}


def host_threads() -> int:
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return os.cpu_count() or 1


def mem_available_gb() -> float:
    try:
        for line in open("/proc/meminfo"):
            if line.startswith("MemAvailable:"):
                return int(line.split()[1]) / 1e6
    except Exception:
        pass
    return 0.0


# ------------------------------------------------------------------ CPU arm
def cpu_corpus_sample(d: int, rows_total: int):
    """Host copy of the benchmark corpus for the CPU legs: the SAME rows the CUDA arm searches
    (generated on the GPU from the same seeds and copied out) -- all of them when host RAM holds
    36 GB comfortably, else a leading block.  Returns (X numpy [rows, d], rows)."""
    import numpy as np
    import torch

    need_gb = rows_total * d * 4 / 1e9
    rows = rows_total if mem_available_gb() > 2.2 * need_gb + 16 else min(rows_total, 1 << 21)
    if torch.cuda.is_available():
        dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
        X = np.empty((rows, d), dtype=np.float32)

        def add(x):
            n0 = add.n
            X[n0:n0 + x.shape[0]] = x.cpu().numpy()
            add.n = n0 + x.shape[0]

        add.n = 0
        fill_rows(add, 0, rows, d, dev, rows_total)
        P, S = make_queries(NQ_FULL, d, dev)
        return X, rows, P.cpu().numpy(), S.cpu().numpy(), "same seeds as the CUDA arm (generated on the GPU, copied to host)"
    # no GPU at all: same formulas from the CPU generator (different random stream)
    rows = min(rows, 1 << 19)
    g = torch.Generator().manual_seed(1234)
    X = torch.nn.functional.normalize(torch.randn((rows, d), generator=g), dim=1).numpy()
    P = torch.nn.functional.normalize(torch.randn((NQ_FULL, d), generator=g), dim=1)
    G = torch.nn.functional.normalize(torch.randn((NQ_FULL, d), generator=g), dim=1)
    S = torch.nn.functional.normalize(0.8 * P + 0.6 * G, dim=1)
    return X, rows, P.numpy(), S.numpy(), "same distribution as the CUDA arm (CPU generator: no GPU visible)"


def cpu_port_step(X, P, S, k: int):
    """One step of the oracle port of the reference's CPU-FAISS path (blocked MKL SGEMM + top-k
    merge, all host threads) on the given queries; returns seconds."""
    import oracle

    t0 = time.perf_counter()
    Q, _ = oracle.mix_normalize(P, S, [ALPHA])
    oracle.flat_ip_search(X, Q[0], k, fast=True)
    return time.perf_counter() - t0


def run_reference(a) -> None:
    """--impl reference: the reference's CPU path.  faiss is not installable here (no
    network, not vendored) so the timed code is the oracle port of FAISS IndexFlatIP's
    CPU algorithm, on all host threads torch/MKL will use.  Rank 0 only.  Each step is a
    bounded sample of the workload: --cpu-sample-queries queries against the full corpus
    (all rows, when host RAM allows); throughput in queries/s needs no extrapolation then."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch

    # torchrun exports OMP_NUM_THREADS=1: the CPU arm must use every host core it may
    threads = host_threads()
    torch.set_num_threads(threads)
    d, k = a.dim, a.k
    X, rows, P, S, how = cpu_corpus_sample(d, a.rows)
    nq_s = min(a.nq, a.cpu_sample_queries)
    times = []
    for i in range(a.warmup + a.steps):
        dt = cpu_port_step(X, P[:nq_s], S[:nq_s], k)
        if i >= a.warmup:
            times.append(dt)
    dt = sum(times) / len(times)
    scale = a.rows / rows
    qps = nq_s / (dt * scale)
    sample = (f"{nq_s} of {a.nq} queries x {rows} of {a.rows} corpus rows per step"
              + ("" if rows == a.rows else ", extrapolated linearly in rows") + f"; {how}")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    line = {
        "impl": "reference", "metric": METRIC, "value": qps, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": 1e3 * dt, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic" if DATA == "iid" else f"synthetic ({DATA})",
        "config": workload_config(a, max(world, a.gpus)),
        "measured_ms_per_step": 1e3 * dt,
        "full_workload_ms_per_step_extrapolated": 1e3 * a.nq / qps,
        "cpu_baseline": {"value": qps, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                         "cpu_count": os.cpu_count()},
        "e2e": {"value": qps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------ parity (untimed)
def brute_force_topk(Q, rows_total: int, d: int, k: int, dev):
    """fp32 brute force, independent of libcmx: scores by torch.matmul (TF32 off) over the corpus
    regenerated chunk by chunk from its seeds, running top-k merge.  Q [ns, d] on dev."""
    import torch

    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        run_v = run_i = None
        c = 0
        while c * CHUNK < rows_total:
            x = corpus_chunk(c, d, dev, rows_total)
            n = min(CHUNK, rows_total - c * CHUNK)
            s = Q @ x[:n].T
            del x
            v, i = torch.topk(s, min(k, n), dim=1)
            i = i + c * CHUNK
            if run_v is None:
                run_v, run_i = v, i
            else:
                cv, ci = torch.cat([run_v, v], dim=1), torch.cat([run_i, i], dim=1)
                run_v, pos = torch.topk(cv, min(k, cv.shape[1]), dim=1)
                run_i = torch.gather(ci, 1, pos)
            c += 1
        return run_v, run_i
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev


def compare_tie_aware(D, I, Dr, Ir, rtol=1e-5, atol=1e-6) -> dict:
    """north_star parity rule: scores within rtol (+atol); ids equal rank by rank except inside a
    tie band (reference scores within tolerance of each other, or of the k-th at the boundary)."""
    import numpy as np

    D, Dr = np.asarray(D, np.float64), np.asarray(Dr, np.float64)
    I, Ir = np.asarray(I), np.asarray(Ir)
    tol = rtol * np.abs(Dr) + atol
    score_bad = int((np.abs(D - Dr) > tol).sum())
    mism = I != Ir
    hard = 0
    for r in np.nonzero(mism.any(axis=1))[0]:
        pos_ref = {int(v): p for p, v in enumerate(Ir[r])}
        kth = Dr[r, -1]
        for p in np.nonzero(mism[r])[0]:
            v, t = int(I[r, p]), 2 * tol[r, p]
            if v in pos_ref:
                hard += abs(Dr[r, pos_ref[v]] - Dr[r, p]) > t
            else:
                hard += abs(D[r, p] - kth) > t
    return {"ok": bool(score_bad == 0 and hard == 0), "id_exact_frac": float(1.0 - mism.mean()),
            "max_rel_err": float(np.max(np.abs(D - Dr) / np.maximum(np.abs(Dr), 1e-30))),
            "tie_band_swaps": int(mism.sum()), "hard_id_mismatches": int(hard), "score_violations": score_bad}


def parity_check(D, I, P, S, rows_total, d, k, dev, world, precision, n_sample=64):
    """Untimed: n_sample queries of the timed result vs the fp32 brute force; at world > 1 also
    bit-equality with a single-GPU search (full corpus on this rank's GPU) of the same queries."""
    import torch

    from cmx.engine import Shard

    nq = P.shape[0]
    sel = torch.linspace(0, nq - 1, min(n_sample, nq), device=dev).round().long().unique()
    Q = torch.nn.functional.normalize((1.0 - ALPHA) * P[sel] + ALPHA * S[sel], dim=1)
    Dr, Ir = brute_force_topk(Q, rows_total, d, k, dev)
    rep = compare_tie_aware(D[0][sel].cpu().numpy(), I[0][sel].cpu().numpy(), Dr.cpu().numpy(), Ir.cpu().numpy())
    rep["n_queries"] = int(sel.numel())
    rep["reference"] = "torch.matmul fp32 (TF32 off) brute force over the corpus regenerated from its seeds"
    if world > 1:
        free, _ = torch.cuda.mem_get_info(dev)
        need = rows_total * d * 6 + (4 << 30)
        if free > need:
            single = Shard(d, dev.index)
            single.set_precision(precision)
            single.reserve(rows_total)
            fill_rows(single.add, 0, rows_total, d, dev, rows_total)
            Ds, Is = single.search_mixed(P[sel].contiguous(), S[sel].contiguous(), [ALPHA], k)
            rep["equals_single_gpu"] = bool(torch.equal(Ds[0], D[0][sel]) and torch.equal(Is[0], I[0][sel]))
            rep["ok"] = rep["ok"] and rep["equals_single_gpu"]
            del single
        else:
            rep["equals_single_gpu"] = None
    return rep


# ------------------------------------------------------------------ CUDA arm
class Timer:
    """CUDA-event timing of K calls bracketed by barrier + synchronize, max over ranks."""

    def __init__(self, world: int, dev):
        self.world, self.dev = world, dev

    def barrier(self):
        import torch
        import torch.distributed as dist

        if self.world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(self, vals):
        import torch
        import torch.distributed as dist

        t = torch.tensor(vals, dtype=torch.float64, device=self.dev)
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(v) for v in t]

    def run(self, fn, steps: int, warmup: int, after_step=None):
        import torch

        for _ in range(warmup):
            fn()
        self.barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        out = None
        for _ in range(steps):
            out = fn()
            if after_step:
                after_step()
        ev1.record()
        self.barrier()
        return self.max_over_ranks([ev0.elapsed_time(ev1)])[0] / steps, out

    def run_wall(self, fn, steps: int, warmup: int):
        """host-clock variant for calls that end with host-visible results (e2e)."""
        import torch

        for _ in range(warmup):
            fn()
        self.barrier()
        t0 = time.perf_counter()
        out = None
        for _ in range(steps):
            out = fn()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) * 1e3
        self.barrier()
        return self.max_over_ranks([dt])[0] / steps, out


def hbm_roofline(n_local: int, d_pad: int, passes: int, score_ms: float, peaks: dict, kernel: str) -> dict:
    alg_bytes = (4.0 if passes == 3 else 2.0) * n_local * d_pad
    achieved = alg_bytes / (score_ms / 1e3) / 1e9
    return {"bound": "hbm", "kernel": kernel, "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
            "frac": achieved / peaks["hbm_gbs"], "traffic": None, "kernel_ms_per_step": score_ms,
            "algorithmic_bytes_per_launch": alg_bytes, "peak_source": peaks["source"],
            "note": "algorithmic bytes = fp16 operand plane(s) read once per sweep (%d B per corpus row); the measured peak is "
                    "a copy (read+write) figure, a read-only stream can exceed it" % int(alg_bytes / max(1, n_local))}


def small_batch_configs(index, P, S, d, peaks, timer, steps=8, warmup=3):
    """C4 regime / the reference's nq=1 calls (onepass_dense_run.py:427,460): one index.search of nq
    queries = one sweep of the resident corpus; HBM-bound.  Device-resident and host-buffer (e2e) times."""
    import torch

    from cmx.engine import mix_normalize

    out = []
    Q = mix_normalize(P[:32].contiguous(), S[:32].contiguous(), [ALPHA])[0].contiguous()
    n_local = index.row1 - index.row0
    d_pad = (d + 63) // 64 * 64
    for k in (K_FULL, 100):
        for nq in (1, 16, 32):
            q = Q[:nq].contiguous()
            q_h = q.cpu().pin_memory()
            D_h = torch.empty((nq, k), dtype=torch.float32).pin_memory()
            I_h = torch.empty((nq, k), dtype=torch.int64).pin_memory()
            acc = {"score": 0.0, "n": 0}

            def after():
                acc["score"] += index.local.last_stats()["score_ms"]; acc["n"] += 1

            ms, _ = timer.run(lambda: index.search(q, k), steps, warmup, after)
            ms_e2e, _ = timer.run_wall(lambda: index.local.search(q_h, k, id_base=index.row0, out=(D_h, I_h)), steps, 2)
            score_ms = acc["score"] / max(1, acc["n"])
            st = index.local.last_stats()
            out.append({"nq": nq, "k": k, "ms_per_search": ms, "e2e_ms_per_search": ms_e2e, "queries_per_s": nq / (ms / 1e3),
                        "slabs": st["slabs"], "path": "tensor" if st["path"] == 2 else "stream",
                        "roofline": hbm_roofline(n_local, d_pad, 1, score_ms, peaks, "tc_score_small_kernel")})
    return out


def subset_config(name, rows, d, nq, k, dev, precision, timer, steps=5, warmup=3):
    """C1 / C5: a 100k-row subset index on this GPU (every rank would hold a replica)."""
    from cmx.engine import Shard

    sh = Shard(d, dev.index)
    sh.set_precision(precision)
    sh.reserve(rows)
    fill_rows(sh.add, 0, rows, d, dev, rows)
    P, S = make_queries(nq, d, dev)
    ms, _ = timer.run(lambda: sh.search_mixed(P, S, [ALPHA], k), steps, warmup)
    st = sh.last_stats()
    del sh
    return {"workload": f"{name}: {rows} x {d}, {nq} queries, alpha={ALPHA}, k={k}", "ms_per_step": ms,
            "queries_per_s": nq / (ms / 1e3), "alg_tflops": 2.0 * nq * rows * d / (ms / 1e3) / 1e12, "slabs": st["slabs"],
            "reruns": st["reruns"]}


def run_cmx(a) -> None:
    import numpy as np  # noqa: F401
    import torch
    import torch.distributed as dist

    from cmx import _lib
    from cmx.dist import ShardedIndex

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the CUDA path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    assert world == a.gpus or world == 1, f"--gpus {a.gpus} but WORLD_SIZE={world}"
    timer = Timer(world, dev)

    d, k, nq, N = a.dim, a.k, a.nq, a.rows
    _lib.set_default_precision(a.precision)
    if os.environ.get("CMX_PAIR"):  # experiments: force the CTA-pair scorer on (1) / off (0)
        _lib.check(_lib.lib().cmx_debug_set_tensor_pair(int(os.environ["CMX_PAIR"])))

    def make_index(rows_total):
        ix = ShardedIndex(d, rows_total, device=local_rank, exchange=a.exchange)
        ix.set_precision(a.precision)
        ix.path = a.path
        fill_shard(ix, d, dev, rows_total)
        return ix

    index = make_index(N)
    P, S = make_queries(nq, d, dev)
    P_h = P.cpu().pin_memory()
    S_h = S.cpu().pin_memory()
    torch.cuda.synchronize()

    def step_device():
        return index.search_mixed(P, S, [ALPHA], k)

    out_h = index.host_output(1, nq, k)

    def step_e2e():
        # the user-facing call with HOST (pinned) buffers: H2D of P,S and D2H of (D,I) inside
        return index.search_mixed_host(P_h, S_h, [ALPHA], k, out=out_h)

    _lib.set_profiling(True)
    sampler = ClockSampler(local_rank)
    acc = {"score": 0.0, "select": 0.0, "launches": 0}

    def after_step():
        st = index.local.last_stats()
        acc["score"] += st["score_ms"]; acc["select"] += st["select_ms"]; acc["launches"] += st["score_launches"]

    for _ in range(a.warmup):
        step_device()
    timer.barrier()
    if rank == 0:
        sampler.start()
    launches0 = _lib.launch_count()
    ms_per_step, (D, I) = timer.run(step_device, a.steps, 0, after_step)
    launches = _lib.launch_count() - launches0
    stats = index.local.last_stats()
    score_ms_max, launches_sum = acc["score"], launches
    if world > 1:
        score_ms_max = timer.max_over_ranks([acc["score"]])[0]
        t = torch.tensor([float(launches)], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        launches_sum = int(t[0])

    # end-to-end through host buffers (the sampler keeps running: the line's clocks cover both loops)
    e2e_ms, (Dh, Ih) = timer.run_wall(step_e2e, a.steps, min(2, a.warmup))
    clocks = sampler.stop() if rank == 0 else None

    if a.stage_times and world > 1:
        index.profile = True
        index.timing = {}
        for _ in range(a.steps):
            step_device()
        if rank == 0:
            print("stage_ms_per_step", {k2: round(v / a.steps, 3) for k2, v in index.timing.items()}, index.local.last_stats(), file=sys.stderr)
        index.profile = False

    # light self-check of the timed result; the sampled brute-force parity follows
    ok = bool((D[0, :, 1:] <= D[0, :, :-1]).all()) and int(I.min()) >= 0 and int(I.max()) < N
    e2e_ok = True
    if rank == 0:
        # the end-to-end call (host buffers) must deliver exactly what the device-resident call computed
        e2e_ok = bool(torch.equal(Dh.reshape(-1), D.reshape(-1).cpu())) and bool(torch.equal(Ih.reshape(-1), I.reshape(-1).cpu()))
        if not (ok and e2e_ok):
            print(f"bench.py: result self-check ok={ok} e2e_matches={e2e_ok}", file=sys.stderr)

    parity = None
    if not a.no_parity:
        if rank == 0:
            parity = parity_check(D, I, P, S, N, d, k, dev, world, a.precision)
        timer.barrier()

    peaks = measured_peaks()
    n_local = index.row1 - index.row0
    d_pad = (d + 63) // 64 * 64
    used_tensor = stats["path"] == 2
    per_step_score_ms = score_ms_max / a.steps
    passes = 3 if a.precision == "split" else 1
    if used_tensor and nq <= 128:
        roofline = hbm_roofline(n_local, d_pad, passes, per_step_score_ms, peaks,
                                "tc_score_small_kernel" if nq <= (64 if passes == 3 else 32) else "tc_score_kernel")
    elif used_tensor:
        alg_flops = 2.0 * nq * n_local * d  # per step on this rank
        executed = passes * 2.0 * nq * n_local * d_pad  # fp16 MMA passes actually issued
        achieved = executed / (per_step_score_ms / 1e3) / 1e12
        peak = peaks["bf16_tflops_sustained"] or peaks["bf16_tflops"]
        traffic, per_row = ncu_traffic("tc_score_kernel" if passes == 3 else "tc_score_kernel_1pass", n_local, d)
        alg_row = 4096 if passes == 3 else 2048
        roofline = {"bound": "tensor", "kernel": "tc_score_kernel", "achieved": achieved, "peak": peak,
                    "unit": "TFLOP/s", "frac": achieved / peak, "traffic": traffic,
                    "traffic_note": f"DRAM bytes of the step's scoring launches = {per_row} B per corpus row (ncu capture, "
                                    f"profiles/ncu_traffic.json) x rows of this rank, summed over the step's launches; algorithmic = {alg_row} B per row "
                                    f"(fp16 {'hi+lo planes' if passes == 3 else 'hi plane'} read once)",
                    "passes": passes,
                    "achieved_alg_fp32_equiv": alg_flops / (per_step_score_ms / 1e3) / 1e12,
                    "kernel_ms_per_step": per_step_score_ms, "launches_per_step": acc["launches"] / a.steps,
                    "peak_source": peaks["source"] + ", sustained dense 16-bit (fp16 == bf16 rate)",
                    "note": "achieved = executed MMA flops (passes x 2*nq*N*d_pad) / CUDA-event time of the scoring launches of one step; "
                            "rescore precision: 1 fp16 pass filters, the survivors are then scored exactly in fp32 (time in select_ms_per_step)"}
    else:
        groups = (nq + 7) // 8
        alg_bytes = groups * 4.0 * n_local * d
        achieved = alg_bytes / (per_step_score_ms / 1e3) / 1e9
        traffic, per_row = ncu_traffic("stream_score_kernel", n_local, d)
        roofline = {"bound": "hbm", "kernel": "stream_score_kernel", "achieved": achieved, "peak": peaks["hbm_gbs"],
                    "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"], "traffic": None if traffic is None else traffic * groups,
                    "traffic_note": f"{per_row} DRAM B per corpus row per pass (ncu capture, profiles/ncu_traffic.json); the measured "
                                    "peak is a copy (read+write) figure, a read-only stream can exceed it",
                    "kernel_ms_per_step": per_step_score_ms, "peak_source": peaks["source"]}

    mem = index.local.memory()
    memory_gb = {k2: round(v / 1e9, 3) for k2, v in mem.items()}
    memory_gb["ratio_to_faiss_flat"] = round((mem["store"] + mem["planes"]) / max(1, mem["faiss_flat"]), 3)
    engine_info = {"path": "tensor" if used_tensor else "stream", "precision": a.precision if used_tensor else "fp32",
                   "slabs": stats["slabs"], "reruns": stats["reruns"], "score_launches_per_step": acc["launches"] / a.steps,
                   "exchange": index.exchange_used if world > 1 else None,
                   "two_phase_rescore": index.two_phase_used if world > 1 else None,
                   "fallback_steps": index.fallback_steps if world > 1 else None,
                   "fallback_chunks": index.fallback_chunks if world > 1 else None}
    # ---- the other BASELINE configs (untimed w.r.t. the headline) ----
    extras = None
    headline_shape = (N, d, nq, k) == (N_FULL, D_FULL, NQ_FULL, K_FULL) and a.path == "auto"
    if not a.no_extras and headline_shape:
        extras = {}
        if world == 1:
            extras["C4_small_batch"] = small_batch_configs(index, P, S, d, peaks, timer)
        fb0, fc0 = index.fallback_steps, index.fallback_chunks
        ms11, _ = timer.run(lambda: index.search_mixed(P, S, SWEEP11, k), 2, 1)
        extras["C4_sweep11"] = {"workload": f"11 alphas x {nq} queries over {N} x {d}, k={k}, one fused call", "n_gpus": world,
                                "ms_per_job": ms11, "queries_per_s": 11 * nq / (ms11 / 1e3),
                                "two_phase": (index.two_phase_used and index.fallback_steps == fb0) if world > 1 else None,
                                "chunks_redone": (index.fallback_chunks - fc0) if world > 1 else None}
        del index
        torch.cuda.empty_cache()
        if world == 1:
            extras["C1"] = subset_config("C1 (configs[0])", 100_000, 1024, nq, k, dev, a.precision, timer)
            extras["C1_k100"] = subset_config("C1 as the reference runs it", 100_000, 1024, nq, 100, dev, a.precision, timer)
            extras["C5_d2560"] = subset_config("C5 (configs[4]) Qwen3-Embedding-4B dim", 100_000, 2560, nq, k, dev, a.precision, timer)
            extras["C5_d4096"] = subset_config("C5 (configs[4]) Qwen3-Embedding-8B dim", 100_000, 4096, nq, k, dev, a.precision, timer)
        # C3: the bilingual combined index, twice the rows (one GPU holds it too: 109 GB)
        index3 = make_index(2 * N)
        ms3, (D3, I3) = timer.run(lambda: index3.search_mixed(P, S, [ALPHA], k), 3, 2)
        out3 = index3.host_output(1, nq, k)
        ms3_e2e, _ = timer.run_wall(lambda: index3.search_mixed_host(P_h, S_h, [ALPHA], k, out=out3), 3, 1)
        extras["C3"] = {"workload": workload_name(2 * N, d, nq, k, a.data), "n_gpus": world, "ms_per_step": ms3,
                        "queries_per_s": nq / (ms3 / 1e3), "e2e_ms_per_step": ms3_e2e, "e2e_queries_per_s": nq / (ms3_e2e / 1e3),
                        "alg_tflops_per_gpu": 2.0 * nq * 2 * N * d / world / (ms3 / 1e3) / 1e12,
                        "slabs": index3.local.last_stats()["slabs"], "reruns": index3.local.last_stats()["reruns"],
                        "two_phase": index3.two_phase_used,
                        "sorted_and_in_range": bool((D3[0, :, 1:] <= D3[0, :, :-1]).all()) and int(I3.min()) >= 0 and int(I3.max()) < 2 * N}
        del index3
        torch.cuda.empty_cache()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    value = nq / (ms_per_step / 1e3)
    e2e_value = nq / (e2e_ms / 1e3)
    cpu_baseline = None
    if world == 1 and not a.no_cpu_baseline:
        torch.set_num_threads(host_threads())
        Xs, rows_s, Pc, Sc, how = cpu_corpus_sample(d, N)
        nq_s = min(nq, a.cpu_sample_queries)
        cpu_port_step(Xs[: 1 << 16], Pc[:nq_s], Sc[:nq_s], k)  # warm the thread pool
        dt = cpu_port_step(Xs, Pc[:nq_s], Sc[:nq_s], k)
        qps = nq_s / (dt * (N / rows_s))
        cpu_baseline = {"value": qps, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port", "cpu_count": os.cpu_count(),
                        "sample": f"{nq_s} of {nq} queries x {rows_s} of {N} corpus rows ({dt:.1f} s)"
                                  + ("" if rows_s == N else ", extrapolated linearly in rows") + f"; {how}"}
        del Xs

    cfg = workload_config(a, world)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": ("f16-filter+f32-exact-rescore" if a.precision == "rescore" else "f16x3-split+f32acc") if used_tensor else "f32",
        "data": "synthetic" if DATA == "iid" else f"synthetic ({DATA})",
        "config": cfg,
        "engine": engine_info,
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_ms, "matches_device_result": e2e_ok,
                "h2d_bytes_per_step": 2 * nq * d * 4, "d2h_bytes_per_step": nq * k * 12,
                "note": "host (pinned) P,S in, host (D,I) out inside the timed region"
                        + ("; every rank uploads 1/N of the queries and writes its query slice of (D,I) straight into ONE shared "
                           "pinned host buffer" if world > 1 else "")},
        "gpu_launches": launches_sum,
        "roofline": roofline,
        "cpu_baseline": cpu_baseline,
        "select_ms_per_step": acc["select"] / a.steps,
        "memory_gb_per_gpu": memory_gb,
        "self_check": ok,
        "parity_check": parity,
        "extra_configs": extras,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    a = parse_args()
    global DATA, ROWS_TOTAL
    DATA, ROWS_TOTAL = a.data, a.rows
    if a.impl == "reference":
        run_reference(a)
    else:
        run_cmx(a)


if __name__ == "__main__":
    main()
