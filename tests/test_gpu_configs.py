"""Parity at the shapes BASELINE.json names (configs[0], [3], [4]); configs[1]/[2] at full size
are exercised by bench.py and by the size-independent property test in test_gpu_parity.py."""
import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu
RTOL, ATOL = 1e-5, 1e-6


def _unit_t(n, d, seed):
    import torch

    g = torch.Generator().manual_seed(seed)
    return torch.nn.functional.normalize(torch.randn((n, d), generator=g), dim=1).numpy()


def _pairs(nq, d):
    P = _unit_t(nq, d, 42)
    G = _unit_t(nq, d, 43)
    S = 0.8 * P + 0.6 * G
    S /= np.linalg.norm(S, axis=1, keepdims=True)
    return P, S.astype(np.float32)


@pytest.mark.parametrize("k", [100, 1000])
def test_config_c1_100k_subset_full_query_set(k):
    """configs[0]: EN-ZH vector mix alpha=0.5, 100k x 1024, all 6980 dev queries."""
    from cmx.engine import Shard

    N, d, nq = 100_000, 1024, 6980
    X = _unit_t(N, d, 1234)
    P, S = _pairs(nq, d)
    sh = Shard(d, 0)
    sh.add(X)
    D, I = sh.search_mixed(P, S, [0.5], k)
    Q, _ = oracle.mix_normalize(P, S, [0.5])
    Dr, Ir = oracle.flat_ip_search(X, Q[0], k, fast=True)
    rep = oracle.compare_topk(D[0], I[0], Dr, Ir, rtol=RTOL, atol=ATOL)
    assert rep["ok"], rep
    assert rep["id_exact_frac"] > 0.99


def test_config_c4_alpha_sweep_and_small_batches():
    """configs[3]: 11 alphas; per-qblock calls and tiny batches give the one-shot rows."""
    from cmx.engine import Shard, mix_normalize

    N, d, nq, k = 200_000, 1024, 512, 100
    X = _unit_t(N, d, 1235)
    P, S = _pairs(nq, d)
    alphas = [round(0.1 * i, 1) for i in range(11)]
    sh = Shard(d, 0)
    sh.add(X)
    D, I = sh.search_mixed(P, S, alphas, k)
    Qo, _ = oracle.mix_normalize(P, S, alphas)
    for ai in (0, 3, 5, 10):
        Dr, Ir = oracle.flat_ip_search(X, Qo[ai], k, fast=True)
        assert oracle.compare_topk(D[ai], I[ai], Dr, Ir, rtol=RTOL, atol=ATOL)["ok"]
    Qg = mix_normalize(P, S, alphas)
    # per-qblock search calls (the reference's loop shape) reproduce the fused result bit for bit
    for qblock in (128, 256):
        for start in range(0, nq, qblock):
            Db, Ib = sh.search(Qg[5][start:start + qblock], k, path="tensor")
            assert np.array_equal(Ib, I[5][start:start + qblock]) and np.array_equal(Db, D[5][start:start + qblock])
    # small-batch regime: the fp32 stream scorer agrees with the tensor scorer within tolerance
    for b in (1, 8, 16, 32):
        Ds, Is = sh.search(Qg[5][:b], k, path="stream")
        assert oracle.compare_topk(Ds, Is, D[5][:b], I[5][:b], rtol=RTOL, atol=ATOL)["ok"]


@pytest.mark.parametrize("d", [2560, 4096])
def test_config_c5_ablation_dims(d):
    """configs[4]: Qwen3-Embedding dims on a 100k subset."""
    from cmx.engine import Shard

    N, nq, k = 100_000, 256, 100
    X = _unit_t(N, d, 1236)
    P, S = _pairs(nq, d)
    sh = Shard(d, 0)
    sh.add(X)
    D, I = sh.search_mixed(P, S, [0.5], k)
    Q, _ = oracle.mix_normalize(P, S, [0.5])
    Dr, Ir = oracle.flat_ip_search(X, Q[0], k, fast=True)
    assert oracle.compare_topk(D[0], I[0], Dr, Ir, rtol=RTOL, atol=ATOL)["ok"]
    Ds, Is = sh.search(Q[0][:4], k, path="stream")
    assert oracle.compare_topk(Ds, Is, Dr[:4], Ir[:4], rtol=RTOL, atol=ATOL)["ok"]


def test_config_c2_full_size_sampled_parity():
    """configs[1] at FULL size (8 841 823 x 1024, 6980 queries, k = 1000, the bench's synthetic corpus):
    * 96 sampled queries against an independent brute force (cuBLAS fp32 GEMM over the regenerated
      corpus chunks + running torch.topk) with the tie-aware comparison of the oracle;
    * every list sorted, ids in range and distinct;
    * the returned score of sampled (query, id) pairs equals an fp64 recomputation."""
    import sys
    import pathlib

    import torch

    root = pathlib.Path(__file__).resolve().parent.parent
    sys.path.insert(0, str(root))
    import bench
    from cmx.engine import Shard, mix_normalize

    free, _total = torch.cuda.mem_get_info()
    if free < 70e9:
        pytest.skip("needs ~65 GB of free HBM")
    N, d, nq, k = bench.N_FULL, bench.D_FULL, bench.NQ_FULL, bench.K_FULL
    dev = torch.device("cuda", 0)
    bench.DATA, bench.ROWS_TOTAL = "iid", N
    sh = Shard(d, 0)
    sh.reserve(N)
    c = 0
    while c * bench.CHUNK < N:
        x = bench.corpus_chunk(c, d, dev)
        sh.add(x[: min(bench.CHUNK, N - c * bench.CHUNK)])
        del x
        c += 1
    P, S = bench.make_queries(nq, d, dev)
    D, I = sh.search_mixed(P, S, [0.5], k)
    assert sh.last_stats()["reruns"] == 0 and sh.last_stats()["path"] == 2
    D, I = D[0], I[0]
    assert bool((D[:, 1:] <= D[:, :-1]).all()) and int(I.min()) >= 0 and int(I.max()) < N
    srt = torch.sort(I, dim=1).values
    assert bool((srt[:, 1:] != srt[:, :-1]).all())
    # independent brute force for a sample of the queries
    Q = mix_normalize(P, S, [0.5])[0]
    sel = torch.arange(0, nq, 73, device=dev)[:96]
    Qs = Q[sel]
    best_d = torch.full((len(sel), k), -float("inf"), device=dev)
    best_i = torch.full((len(sel), k), -1, dtype=torch.int64, device=dev)
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        c = 0
        while c * bench.CHUNK < N:
            x = bench.corpus_chunk(c, d, dev)[: min(bench.CHUNK, N - c * bench.CHUNK)]
            sc = Qs @ x.T
            dd, ii = torch.topk(sc, k, dim=1)
            cat_d = torch.cat([best_d, dd], dim=1)
            cat_i = torch.cat([best_i, ii + c * bench.CHUNK], dim=1)
            best_d, pos = torch.topk(cat_d, k, dim=1)
            best_i = torch.gather(cat_i, 1, pos)
            if c == 3:  # fp64 recomputation of returned scores whose rows live in this chunk
                lo, hi = c * bench.CHUNK, c * bench.CHUNK + x.shape[0]
                Is, Ds = I[sel], D[sel]
                m = (Is >= lo) & (Is < hi)
                qi, pi = m.nonzero(as_tuple=True)
                rows = x[Is[qi, pi] - lo].double()
                exact = (rows * Qs[qi].double()).sum(dim=1)
                err = (Ds[qi, pi].double() - exact).abs().max().item()
                assert len(qi) > 1000 and err < 2e-7, (len(qi), err)
            del x, sc
            c += 1
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev
    rep = oracle.compare_topk(D[sel].cpu().numpy(), I[sel].cpu().numpy(), best_d.cpu().numpy(), best_i.cpu().numpy(),
                              rtol=RTOL, atol=ATOL)
    del sh
    torch.cuda.empty_cache()
    assert rep["ok"], rep
    assert rep["id_exact_frac"] > 0.995, rep
