"""bench.py's own checker logic (no GPU): the tie-aware comparator used by `parity_check` agrees with the oracle's
comparator, the workload/config helpers are consistent between the two arms, the line's bookkeeping helpers work."""
import argparse
import pathlib
import sys

import numpy as np

import oracle

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402


def _results(rng, nq=12, k=50, n=4000, d=16):
    X = rng.standard_normal((n, d)).astype(np.float32)
    Q = rng.standard_normal((nq, d)).astype(np.float32)
    return oracle.flat_ip_search(X, Q, k)


def test_compare_tie_aware_matches_oracle_rule():
    rng = np.random.default_rng(0)
    D, I = _results(rng)
    rep = bench.compare_tie_aware(D, I, D, I)
    assert rep["ok"] and rep["id_exact_frac"] == 1.0 and rep["max_rel_err"] == 0.0
    # a swap inside a tie band is fine, a swap of clearly different scores is not
    D2, I2 = D.copy(), I.copy()
    D2[0, 3] = D2[0, 4] = D[0, 3]
    Dref = D2.copy()
    I2[0, 3], I2[0, 4] = I[0, 4], I[0, 3]
    rep = bench.compare_tie_aware(D2, I2, Dref, I)
    assert rep["ok"] and rep["tie_band_swaps"] == 2 and rep["hard_id_mismatches"] == 0
    assert oracle.compare_topk(D2, I2, Dref, I)["ok"]
    I3 = I.copy()
    I3[1, 0], I3[1, 40] = I[1, 40], I[1, 0]
    rep = bench.compare_tie_aware(D, I3, D, I)
    assert not rep["ok"] and rep["hard_id_mismatches"] == 2
    assert not oracle.compare_topk(D, I3, D, I)["ok"]
    # a boundary replacement whose score equals the k-th within tolerance passes; a worse one fails
    I4, D4 = I.copy(), D.copy()
    I4[2, -1] = 10**6
    assert bench.compare_tie_aware(D4, I4, D, I)["ok"]
    D4[2, -1] = D[2, -1] - 0.01
    rep = bench.compare_tie_aware(D4, I4, D, I)
    assert not rep["ok"] and rep["score_violations"] == 1
    # scores off by more than 1e-5 relative are violations even when ids agree
    D5 = D * np.float32(1.0001)
    assert not bench.compare_tie_aware(D5, I, D, I)["ok"]


def test_both_arms_print_the_same_config():
    a = argparse.Namespace(rows=bench.N_FULL, dim=bench.D_FULL, nq=bench.NQ_FULL, k=bench.K_FULL, data="iid")
    c1, c8 = bench.workload_config(a, 1), bench.workload_config(a, 8)
    assert c1["workload"].startswith("C2 (BASELINE configs[1])") and c1["rows"] == 8_841_823 and c1["k"] == 1000
    assert set(c1) == set(c8) == {"workload", "rows", "dim", "queries", "k", "alpha", "parallelism", "cache"}
    assert c8["parallelism"] == "corpus row shards x8" and "inputs_larger_than_L2" in c8["cache"]
    a3 = argparse.Namespace(rows=2 * bench.N_FULL, dim=1024, nq=6980, k=1000, data="iid")
    assert bench.workload_config(a3, 2)["workload"].startswith("C3 (BASELINE configs[2])")


def test_hbm_roofline_block():
    peaks = {"hbm_gbs": 6455.6, "source": "test"}
    r = bench.hbm_roofline(8_841_823, 1024, 1, 2.6, peaks, "tc_score_small_kernel")
    assert r["bound"] == "hbm" and abs(r["algorithmic_bytes_per_launch"] - 2 * 8_841_823 * 1024) < 1
    assert abs(r["achieved"] - 2 * 8_841_823 * 1024 / 2.6e-3 / 1e9) < 1e-6 and abs(r["frac"] - r["achieved"] / 6455.6) < 1e-12
