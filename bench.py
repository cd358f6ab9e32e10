#!/usr/bin/env python
"""Benchmark of the hot path: vector-mix query step + flat inner-product top-k.

    python bench.py --gpus N --steps K --warmup W            # the CUDA path (this repo)
    python bench.py --impl reference --gpus N --steps K ...  # the CPU path (oracle port)

A "step" = one alpha-pass of the hot path: mix+normalise nq cached query pairs at
alpha=0.5 and search the whole corpus for the top k (BASELINE.json configs[1]:
8 841 823 x 1024 fp32, 6980 queries, k=1000).  At N > 1 GPUs the SAME corpus is
row-sharded over the ranks (strong scaling) with one all_gather + merge per step.
Prints ONE JSON line on rank 0.  Synthetic data: corpus rows
normalize(N(0,I)) generated on device in 2^20-row chunks seeded 1234+chunk, queries
P seed 42, S = normalize(0.8 P + 0.6 normalize(N(0,I))) seed 43.
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
    ap.add_argument("--stage-times", action="store_true", help="after the timed runs, print a per-stage breakdown (extra syncs)")
    ap.add_argument("--cpu-sample-queries", type=int, default=512)
    return ap.parse_args()


def workload_name(a) -> str:
    shape = (a.rows, a.dim, a.nq, a.k)
    if shape == (N_FULL, D_FULL, NQ_FULL, K_FULL):
        base = "C2 (BASELINE configs[1]): EN monolingual full mMARCO shape"
    elif shape == (2 * N_FULL, D_FULL, NQ_FULL, K_FULL):
        base = "C3 (BASELINE configs[2]): EN+ZH bilingual combined index shape"
    else:
        base = "other shape (not the headline configuration)"
    var = "" if a.data == "iid" else f", data variant {a.data}"
    return f"{base}: {a.rows} x {a.dim} fp32 flat-IP, {a.nq} queries, alpha={ALPHA}, k={a.k}{var}"


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


def corpus_chunk(c: int, d: int, device):
    """Global chunk c (rows [c*2^20, (c+1)*2^20)), identical for every GPU count."""
    import torch

    g = torch.Generator(device=device).manual_seed(1234 + c)
    x = torch.randn((CHUNK, d), generator=g, device=device)
    if DATA == "aniso":
        x += 1.5 * (d ** 0.5) * _direction(d, device)
    elif DATA == "shift":  # rows of the second half of the corpus lean towards u
        rows = torch.arange(c * CHUNK, (c + 1) * CHUNK, device=device)
        x += (rows >= ROWS_TOTAL // 2).to(x.dtype)[:, None] * (0.5 * (d ** 0.5)) * _direction(d, device)
    return torch.nn.functional.normalize(x, dim=1)


def fill_shard(index, row0: int, row1: int, d: int, device) -> None:
    index.reserve_local()
    c = row0 // CHUNK
    while c * CHUNK < row1:
        x = corpus_chunk(c, d, device)
        lo = max(row0, c * CHUNK) - c * CHUNK
        hi = min(row1, (c + 1) * CHUNK) - c * CHUNK
        index.add_local(x[lo:hi])
        del x
        c += 1


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
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200", "-i", str(self.gpu)],
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
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(smax), "power_w_max": max(power),
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


# ------------------------------------------------------------------ CPU arm
def cpu_port_qps(X_sample, P, S, n_full: int, k: int, nq_sample: int):
    """The oracle port of the reference's CPU-FAISS path (blocked MKL SGEMM + top-k merge)
    timed on a bounded sample: nq_sample queries x len(X_sample) rows, extrapolated
    linearly in the row count to n_full (cost is 2*nq*N*d flops + O(nq*N) selection)."""
    import numpy as np
    import torch

    import oracle

    Ps, Ss = P[:nq_sample], S[:nq_sample]
    t0 = time.perf_counter()
    Q, _ = oracle.mix_normalize(Ps, Ss, [ALPHA])
    oracle.flat_ip_search(X_sample, Q[0], k, fast=True)
    dt = time.perf_counter() - t0
    scale = n_full / X_sample.shape[0]
    return nq_sample / (dt * scale), dt, torch.get_num_threads()


def run_reference(a) -> None:
    """--impl reference: the reference's CPU path.  faiss is not installable here (no
    network, not vendored) so the timed code is the oracle port of FAISS IndexFlatIP's
    CPU algorithm, on all host threads torch/MKL will use.  Rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import numpy as np
    import torch

    # torchrun exports OMP_NUM_THREADS=1: the CPU arm must use every host core it may
    try:
        ncores = len(os.sched_getaffinity(0))
    except Exception:
        ncores = os.cpu_count() or 1
    torch.set_num_threads(max(1, ncores))
    d, k = a.dim, a.k
    sample_rows = min(a.rows, 1 << 19)
    nq_s = min(a.nq, a.cpu_sample_queries)
    rng = np.random.default_rng(1234)
    X = rng.standard_normal((sample_rows, d), dtype=np.float32)
    X /= np.linalg.norm(X, axis=1, keepdims=True)
    P = rng.standard_normal((nq_s, d), dtype=np.float32)
    P /= np.linalg.norm(P, axis=1, keepdims=True)
    G = rng.standard_normal((nq_s, d), dtype=np.float32)
    G /= np.linalg.norm(G, axis=1, keepdims=True)
    S = 0.8 * P + 0.6 * G
    S /= np.linalg.norm(S, axis=1, keepdims=True)
    times = []
    for i in range(a.warmup + a.steps):
        qps, dt, threads = cpu_port_qps(X, P, S, a.rows, k, nq_s)
        if i >= a.warmup:
            times.append(dt)
    dt = sum(times) / len(times)
    qps = nq_s / (dt * (a.rows / sample_rows))
    sample = f"{nq_s} queries x {sample_rows} rows per step, extrapolated linearly in rows to {a.rows}"
    line = {
        "impl": "reference", "metric": METRIC, "value": qps, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": 1e3 * a.nq / qps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic" if DATA == "iid" else f"synthetic ({DATA})",
        "config": {"workload": workload_name(a), "cache": "inputs_larger_than_L2"},
        "cpu_baseline": {"value": qps, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                         "cpu_count": os.cpu_count()},
        "e2e": {"value": qps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------ CUDA arm
def run_cmx(a) -> None:
    import numpy as np
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

    d, k, nq, N = a.dim, a.k, a.nq, a.rows
    _lib.set_default_precision(a.precision)
    index = ShardedIndex(d, N, device=local_rank, exchange=a.exchange)
    index.set_precision(a.precision)
    index.path = a.path
    fill_shard(index, index.row0, index.row1, d, dev)
    assert index.local_complete()
    P, S = make_queries(nq, d, dev)
    P_h = P.cpu().pin_memory()
    S_h = S.cpu().pin_memory()
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device():
        return index.search_mixed(P, S, [ALPHA], k)

    D_h = torch.empty((1, nq, k), dtype=torch.float32).pin_memory()
    I_h = torch.empty((1, nq, k), dtype=torch.int64).pin_memory()

    def step_e2e():
        # the user-facing call with HOST (pinned) buffers: H2D of P,S and D2H of (D,I) inside
        if world == 1:
            index.local.search_mixed(P_h, S_h, [ALPHA], k, id_base=index.row0, path=a.path, out=(D_h, I_h))
        else:
            Pd = P_h.to(dev, non_blocking=True)
            Sd = S_h.to(dev, non_blocking=True)
            Dd, Id = index.search_mixed(Pd, Sd, [ALPHA], k)
            if rank == 0:  # the merged result is delivered to the host once (rank 0 writes the run file)
                D_h.copy_(Dd, non_blocking=True)
                I_h.copy_(Id, non_blocking=True)
            torch.cuda.current_stream().synchronize()
        return D_h, I_h

    _lib.set_profiling(True)
    for _ in range(a.warmup):
        step_device()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = _lib.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    score_ms = select_ms = 0.0
    score_launches = 0
    barrier()
    ev0.record()
    for _ in range(a.steps):
        D, I = step_device()
        st = index.local.last_stats()
        score_ms += st["score_ms"]; select_ms += st["select_ms"]; score_launches += st["score_launches"]
    ev1.record()
    barrier()
    launches = _lib.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    t_ms = ev0.elapsed_time(ev1)
    stats = index.local.last_stats()

    # end-to-end through host buffers
    for _ in range(min(2, a.warmup)):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        Dh, Ih = step_e2e()
    torch.cuda.synchronize()
    t_e2e = time.perf_counter() - t0

    tt = torch.tensor([t_ms, t_e2e * 1e3, float(launches), score_ms], dtype=torch.float64, device=dev)
    if world > 1:
        tmax = tt.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = tt.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        t_ms, t_e2e_ms, score_ms_max = float(tmax[0]), float(tmax[1]), float(tmax[3])
        launches = int(tsum[2])
    else:
        t_e2e_ms, score_ms_max = t_e2e * 1e3, score_ms

    if a.stage_times and world > 1:
        index.profile = True
        index.timing = {}
        for _ in range(a.steps):
            step_device()
        if rank == 0:
            print("stage_ms_per_step", {k2: round(v / a.steps, 3) for k2, v in index.timing.items()}, index.local.last_stats(), file=sys.stderr)
        index.profile = False

    # light self-check of the timed result (full parity lives in tests/)
    ok = bool((D[0, :, 1:] <= D[0, :, :-1]).all()) and int(I.min()) >= 0 and int(I.max()) < N
    if rank == 0 and not ok:
        print("bench.py: result self-check FAILED", file=sys.stderr)
    # the end-to-end call (host buffers) must deliver exactly what the device-resident call computed
    e2e_ok = True
    if rank == 0:
        e2e_ok = bool(torch.equal(Dh.reshape(-1), D.reshape(-1).cpu())) and bool(torch.equal(Ih.reshape(-1), I.reshape(-1).cpu()))
        if not e2e_ok:
            print("bench.py: e2e result differs from the device-resident result", file=sys.stderr)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    ms_per_step = t_ms / a.steps
    value = nq / (ms_per_step / 1e3)
    e2e_value = nq / (t_e2e_ms / a.steps / 1e3)
    peaks = measured_peaks()
    n_local = index.row1 - index.row0
    d_pad = (d + 63) // 64 * 64
    used_tensor = stats["path"] == 2
    per_step_score_ms = score_ms_max / a.steps
    passes = 3 if a.precision == "split" else 1
    if used_tensor and nq <= 128:
        # small batches: the tensor kernels are bound by the HBM stream of the fp16 operand plane(s)
        alg_bytes = (4.0 if passes == 3 else 2.0) * n_local * d_pad
        achieved = alg_bytes / (per_step_score_ms / 1e3) / 1e9
        roofline = {"bound": "hbm", "kernel": "tc_score_small_kernel" if nq <= (64 if passes == 3 else 32) else "tc_score_kernel",
                    "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"],
                    "traffic": None, "kernel_ms_per_step": per_step_score_ms, "peak_source": peaks["source"],
                    "note": "algorithmic bytes = fp16 operand plane(s) read once per sweep (%d B per corpus row); the measured peak is "
                            "a copy (read+write) figure, a read-only stream can exceed it" % int(alg_bytes / n_local)}
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
                    "kernel_ms_per_step": per_step_score_ms, "launches_per_step": score_launches / a.steps,
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

    cpu_baseline = None
    if world == 1 and not a.no_cpu_baseline:
        rows_s = min(n_local, 1 << 19)
        nq_s = min(nq, a.cpu_sample_queries)
        Xs = index.local.reconstruct_n(0, rows_s)
        try:
            torch.set_num_threads(max(1, len(os.sched_getaffinity(0))))
        except Exception:
            pass
        qps, dt, threads = cpu_port_qps(Xs, P_h.numpy(), S_h.numpy(), N, k, nq_s)
        cpu_baseline = {"value": qps, "unit": UNIT, "cores": threads, "kind": "port", "cpu_count": os.cpu_count(),
                        "sample": f"{nq_s} queries x {rows_s} rows ({dt:.1f} s), extrapolated linearly in rows to {N}"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": ("f16-filter+f32-exact-rescore" if a.precision == "rescore" else "f16x3-split+f32acc") if used_tensor else "f32",
        "data": "synthetic" if DATA == "iid" else f"synthetic ({DATA})",
        "config": {"workload": workload_name(a), "rows": N, "dim": d, "queries": nq, "k": k, "alpha": ALPHA,
                   "parallelism": f"corpus row shards x{world}, exchange={index.exchange_used}, two_phase_rescore={index.two_phase_used}" if world > 1 else "single GPU",
                   "cache": "inputs_larger_than_L2 (corpus %.1f GB per GPU)" % (n_local * d * 4 / 1e9),
                   "path": "tensor" if used_tensor else "stream", "precision": a.precision if used_tensor else "fp32",
                   "slabs": stats["slabs"], "reruns": stats["reruns"]},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": t_e2e_ms / a.steps, "matches_device_result": e2e_ok,
                "h2d_bytes_per_step": 2 * nq * d * 4, "d2h_bytes_per_step": nq * k * 12},
        "gpu_launches": launches,
        "roofline": roofline,
        "cpu_baseline": cpu_baseline,
        "select_ms_per_step": select_ms / a.steps,
        "self_check": ok,
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
