"""GPU parity tests (run with -m gpu on a B200): the CUDA path, called through the C ABI
(ctypes -> libcmx.so), against the CPU oracle on the same seeded inputs.

Bars (BASELINE north_star): scores within 1e-5 relative (+1e-6 absolute floor for
near-zero scores) of the fp32 oracle; ids equal rank by rank except inside score ties
within that tolerance; the mix is bit-exact and the normalise within 2 ulp.
"""
import json

import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu

RTOL, ATOL = 1e-5, 1e-6


def _unit(rng, n, d):
    x = rng.standard_normal((n, d)).astype(np.float32)
    return x / np.linalg.norm(x, axis=1, keepdims=True)


def _aniso(rng, n, d, c=1.5):
    u = np.random.default_rng(7).standard_normal(d).astype(np.float32)
    u /= np.linalg.norm(u)
    x = rng.standard_normal((n, d)).astype(np.float32) + c * np.sqrt(d) * u
    return (x / np.linalg.norm(x, axis=1, keepdims=True)).astype(np.float32)


def _shard(X, device=0):
    from cmx.engine import Shard

    sh = Shard(X.shape[1], device)
    if X.shape[0]:
        sh.add(X)
    return sh


def _check(D, I, X, Q, k, **kw):
    Dr, Ir = oracle.flat_ip_search(X, Q, k)
    rep = oracle.compare_topk(D, I, Dr, Ir, rtol=RTOL, atol=ATOL)
    assert rep["ok"], rep
    return rep


# ---------------------------------------------------------------- prologue
def test_mix_golden_bit_exact(golden_dir):
    from cmx.engine import mix_normalize

    g = np.load(golden_dir / "mix_golden.npz")
    alphas = g["alphas"].tolist()
    out, flags = mix_normalize(g["P"], g["S"], alphas, want_flags=True)
    ref = g["Q"]
    oq, of = oracle.mix_normalize(g["P"], g["S"], alphas)
    assert np.array_equal(flags, of)
    for ai, a in enumerate(alphas):
        endpoint = abs(a) <= 1e-8 or abs(a - 1.0) <= 1e-8
        for qi in range(ref.shape[1]):
            got, want = out[ai, qi], ref[ai, qi]
            if endpoint or flags[ai, qi]:
                assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), (a, qi)
            else:
                ulp = np.abs(got.view(np.int32).astype(np.int64) - want.view(np.int32).astype(np.int64))
                assert ulp.max() <= 2, (a, qi, int(ulp.max()))


def test_mix_random_vs_oracle_and_device_io():
    import torch
    from cmx.engine import mix_normalize

    rng = np.random.default_rng(10)
    P, S = _unit(rng, 700, 1024), _unit(rng, 700, 1024)
    alphas = [0, 0.1, 0.3, 0.5, 0.7, 0.9, 1]
    out = mix_normalize(P, S, alphas)
    ref, _ = oracle.mix_normalize(P, S, alphas)
    # the reference's own reduction order is device dependent (torch CPU vs torch CUDA differ by a
    # few ulp in the norm): a few ulp against the torch-CPU oracle, and no worse than it against fp64
    ulp = np.abs(out.view(np.int32).astype(np.int64) - ref.view(np.int32).astype(np.int64))
    assert ulp.max() <= 4
    truth = oracle.mix_normalize_f64(P, S, alphas)
    big = np.abs(truth) > 1e-6
    rel_gpu = (np.abs(out - truth) / np.abs(truth))[big].max()
    rel_cpu = (np.abs(ref - truth) / np.abs(truth))[big].max()
    assert rel_gpu <= max(3e-7, 1.5 * rel_cpu), (rel_gpu, rel_cpu)
    assert np.array_equal(out[0], P) and np.array_equal(out[-1], S)
    # unnormalised mix must be bit exact: undo nothing, check through an odd dim (scalar path)
    P2, S2 = _unit(rng, 33, 70), _unit(rng, 33, 70)
    o2 = mix_normalize(P2, S2, [0.25])
    r2, _ = oracle.mix_normalize(P2, S2, [0.25])
    assert np.abs(o2.view(np.int32).astype(np.int64) - r2.view(np.int32).astype(np.int64)).max() <= 4
    # device-resident io gives the same bits as host io
    od = mix_normalize(torch.from_numpy(P).cuda(), torch.from_numpy(S).cuda(), alphas)
    assert np.array_equal(od.cpu().numpy().view(np.uint32), out.view(np.uint32))


# ---------------------------------------------------------------- stream path
@pytest.mark.parametrize("nq", [1, 2, 3, 5, 8])
@pytest.mark.parametrize("k", [1, 10, 100])
def test_stream_search_parity(nq, k):
    rng = np.random.default_rng(100 + nq + k)
    X, Q = _unit(rng, 20011, 128), _unit(rng, nq, 128)
    sh = _shard(X)
    D, I = sh.search(Q, k, path="stream")
    _check(D, I, X, Q, k)
    assert sh.last_stats()["path"] == 1


def test_stream_many_queries_and_k1000():
    rng = np.random.default_rng(11)
    X, Q = _aniso(rng, 50000, 256), _aniso(rng, 21, 256)
    sh = _shard(X)
    D, I = sh.search(Q, 1000, path="stream")
    _check(D, I, X, Q, 1000)
    assert np.all(np.diff(D, axis=1) <= 0)


# ---------------------------------------------------------------- tensor path
@pytest.mark.parametrize("d", [64, 256, 1024])
@pytest.mark.parametrize("k", [10, 100, 1000])
def test_tensor_search_parity(d, k, precision):
    rng = np.random.default_rng(200 + d + k)
    X, Q = _unit(rng, 30077, d), _unit(rng, 301, d)
    sh = _shard(X)
    D, I = sh.search(Q, k, path="tensor")
    rep = _check(D, I, X, Q, k)
    assert sh.last_stats()["path"] == 2
    assert rep["max_rel_err"] < RTOL


def test_tensor_anisotropic_realistic_scores(precision):
    rng = np.random.default_rng(12)
    X, Q = _aniso(rng, 40000, 1024), _aniso(rng, 200, 1024)
    sh = _shard(X)
    D, I = sh.search(Q, 100, path="tensor")
    rep = _check(D, I, X, Q, 100)
    assert D[:, 0].min() > 0.5  # scores in the range real BGE-M3 cosines live in
    Dt, It = oracle.flat_ip_search_f64(X, Q, 100)
    assert oracle.compare_topk(D, I, Dt, It, rtol=RTOL, atol=ATOL)["ok"]


@pytest.mark.parametrize("bn", [128, 256])
def test_tensor_tile_variants(bn):
    from cmx import _lib

    rng = np.random.default_rng(13)
    X, Q = _unit(rng, 9000, 192), _unit(rng, 130, 192)
    _lib.check(_lib.lib().cmx_debug_set_tensor_tile(bn))
    try:
        sh = _shard(X)
        D, I = sh.search(Q, 50, path="tensor")
        _check(D, I, X, Q, 50)
    finally:
        _lib.check(_lib.lib().cmx_debug_set_tensor_tile(256))


@pytest.mark.parametrize("pair", [0, 1])
@pytest.mark.parametrize("shape", [(30077, 301, 256, 100), (9000, 130, 192, 50), (70001, 700, 1024, 1000), (5003, 257, 96, 20)])
def test_tensor_single_and_pair_kernels(pair, shape, precision):
    """Both tensor-core kernels (one CTA per tile / CTA pair with cta_group::2) against the oracle."""
    from cmx import _lib

    N, nq, d, k = shape
    rng = np.random.default_rng(300 + N)
    X, Q = _unit(rng, N, d), _unit(rng, nq, d)
    _lib.check(_lib.lib().cmx_debug_set_tensor_pair(pair))
    try:
        sh = _shard(X)
        D, I = sh.search(Q, k, path="tensor")
        _check(D, I, X, Q, k)
    finally:
        _lib.check(_lib.lib().cmx_debug_set_tensor_pair(-1))


@pytest.mark.parametrize("small", [0, 1])
@pytest.mark.parametrize("nq", [5, 16, 17, 33, 64, 100, 128])
def test_tensor_small_batch_kernel(small, nq, precision):
    """nq <= 128: the corpus-as-M tensor kernel (and, for comparison, the padded 128-row one)."""
    from cmx import _lib

    rng = np.random.default_rng(400 + nq)
    d = 1024 if nq in (16, 128) else 128
    X, Q = _unit(rng, 20011, d), _unit(rng, nq, d)
    X[9000:9040] = X[100:140]  # ties
    _lib.check(_lib.lib().cmx_debug_set_tensor_small(small))
    try:
        sh = _shard(X)
        for k in (1, 100, 1000):
            D, I = sh.search(Q, k, path="tensor")
            _check(D, I, X, Q, k)
    finally:
        _lib.check(_lib.lib().cmx_debug_set_tensor_small(1))


@pytest.mark.parametrize("d", [96, 100, 768])
def test_tensor_ragged_dims(d, precision):
    """d not a multiple of 64 -> zero-padded operand planes; N, nq not tile multiples."""
    rng = np.random.default_rng(14 + d)
    X, Q = _unit(rng, 5003, d), _unit(rng, 129, d)
    sh = _shard(X)
    D, I = sh.search(Q, 20, path="tensor")
    _check(D, I, X, Q, 20)
    D2, I2 = sh.search(Q[:5], 20, path="stream")
    assert oracle.compare_topk(D2, I2, D[:5], I[:5], rtol=RTOL, atol=ATOL)["ok"]


def test_tensor_unnormalised_wide_range(precision):
    """Raw (non unit-norm) vectors with a wide dynamic range: the power-of-two operand
    scaling keeps the split exact enough."""
    rng = np.random.default_rng(15)
    X = (rng.standard_normal((20000, 128)) * np.exp(rng.uniform(-3, 3, (20000, 1)))).astype(np.float32)
    Q = (rng.standard_normal((64, 128)) * 37.0).astype(np.float32)
    sh = _shard(X)
    D, I = sh.search(Q, 30, path="tensor")
    Dt, It = oracle.flat_ip_search_f64(X, Q, 30)
    assert oracle.compare_topk(D, I, Dt, It, rtol=RTOL, atol=ATOL)["ok"]


# ---------------------------------------------------------------- edge cases
@pytest.mark.parametrize("path", ["stream", "tensor"])
def test_k_larger_than_ntotal_pads(path, precision):
    rng = np.random.default_rng(16)
    X, Q = _unit(rng, 7, 64), _unit(rng, 3, 64)
    D, I = _shard(X).search(Q, 12, path=path)
    Dr, Ir = oracle.flat_ip_search(X, Q, 12)
    assert np.array_equal(I[:, 7:], Ir[:, 7:]) and (I[:, 7:] == -1).all()
    assert (D[:, 7:] == np.finfo(np.float32).min).all()
    assert oracle.compare_topk(D, I, Dr, Ir, rtol=RTOL, atol=ATOL)["ok"]


def test_empty_index_and_empty_queries():
    from cmx.engine import Shard

    sh = Shard(32, 0)
    D, I = sh.search(np.zeros((2, 32), np.float32), 5)
    assert (I == -1).all() and (D == np.finfo(np.float32).min).all()
    sh.add(np.ones((3, 32), np.float32))
    D, I = sh.search(np.zeros((0, 32), np.float32), 5)
    assert D.shape == (0, 5) and I.shape == (0, 5)


@pytest.mark.parametrize("path", ["stream", "tensor"])
def test_duplicate_rows_tie_order(path, precision):
    """Exact ties come out in ascending row order (deterministic; one of FAISS' legal orders)."""
    rng = np.random.default_rng(17)
    X = _unit(rng, 3000, 64)
    X[1000:2000] = X[0:1000]  # every row twice
    Q = _unit(rng, 6, 64)
    D, I = _shard(X).search(Q, 40, path=path)
    Dr, Ir = oracle.flat_ip_search(X, Q, 40)
    assert oracle.compare_topk(D, I, Dr, Ir, rtol=RTOL, atol=ATOL)["ok"]
    pairs = 0
    for r in range(6):
        row = I[r].tolist()
        for pos, a in enumerate(row):
            if a < 1000 and (a + 1000) in row:
                # the copy scores bit-identically and must follow its original immediately
                assert row[pos + 1] == a + 1000 and D[r, pos] == D[r, pos + 1], (r, a)
                pairs += 1
    assert pairs > 0


@pytest.mark.parametrize("path", ["stream", "tensor"])
def test_zero_and_nan_queries(path, precision):
    rng = np.random.default_rng(18)
    X = _unit(rng, 900, 64)
    Q = _unit(rng, 4, 64)
    Q[1] = 0.0  # P = -S at alpha = 0.5: all scores tie at 0 -> first k rows
    Q[2, 5] = np.nan  # never enters a FAISS heap -> all -1
    D, I = _shard(X).search(Q, 10, path=path)
    assert I[1].tolist() == list(range(10)) and (D[1] == 0).all()
    assert (I[2] == -1).all() and (D[2] == np.finfo(np.float32).min).all()
    Dr, Ir = oracle.flat_ip_search(X, Q[[0, 3]], 10)
    assert oracle.compare_topk(D[[0, 3]], I[[0, 3]], Dr, Ir, rtol=RTOL, atol=ATOL)["ok"]


@pytest.mark.parametrize("path", ["stream", "tensor"])
def test_adversarial_order_triggers_safe_rerun(path, precision):
    """Rows sorted by ascending score for one query, visited in FILE order: every row beats the
    stale threshold, the candidate buffer overflows, the search is redone (split precision with the
    planned slabs, then worst-case-safe slabs)."""
    from cmx import _lib

    rng = np.random.default_rng(19)
    d = 64
    q = _unit(rng, 1, d)
    X = _unit(rng, 12000, d)
    X = X[np.argsort(X @ q[0])]  # ascending score for q
    Q = np.concatenate([q, _unit(rng, 9, d)], axis=0)
    sh = _shard(X)
    sh.set_cand_capacity(256)
    _lib.check(_lib.lib().cmx_debug_set_block_order(0))
    try:
        D, I = sh.search(Q, 100, path=path)
    finally:
        _lib.check(_lib.lib().cmx_debug_set_block_order(1))
    st = sh.last_stats()
    # stream: planned slabs overflow, safe slabs succeed (1); tensor: the schedule with a speculative
    # last slab fails first (2), and the rescore precision gives way to split before the safe one (3)
    assert st["reruns"] == (1 if path == "stream" else (3 if precision == "rescore" else 2)), st
    _check(D, I, X, Q, 100)


@pytest.mark.parametrize("nq", [9, 200])
def test_block_order_makes_nonstationary_corpus_benign(nq, precision):
    """The tensor path visits 256-row blocks in a golden-ratio order, so its first slabs sample the
    whole corpus.  A corpus whose second half scores higher for every query (EN rows then ZH rows
    of the bilingual index, ZH-leaning queries) overflows the buffers in file order (thresholds
    learnt on the first half admit ~11 000 rows of the second) and needs no rerun in block order."""
    from cmx import _lib

    rng = np.random.default_rng(191)
    d = 64
    u = _unit(rng, 1, d)
    X = _unit(rng, 40000, d)
    X[20000:] = X[20000:] + 0.6 * u
    X = (X / np.linalg.norm(X, axis=1, keepdims=True)).astype(np.float32)
    Q = _unit(rng, nq, d) + 1.5 * u
    Q = (Q / np.linalg.norm(Q, axis=1, keepdims=True)).astype(np.float32)
    sh = _shard(X)
    sh.set_cand_capacity(1024)
    D, I = sh.search(Q, 100, path="tensor")
    assert sh.last_stats()["reruns"] == 0, sh.last_stats()
    _check(D, I, X, Q, 100)
    _lib.check(_lib.lib().cmx_debug_set_block_order(0))
    try:
        D2, I2 = sh.search(Q, 100, path="tensor")
    finally:
        _lib.check(_lib.lib().cmx_debug_set_block_order(1))
    assert sh.last_stats()["reruns"] >= 1, sh.last_stats()
    _check(D2, I2, X, Q, 100)  # same answer either way


def test_speculative_last_slab_equals_planned_slabs(precision):
    """The tensor path scores the tail of the corpus in ONE slab under a guessed threshold (verified
    per query afterwards): fewer slabs, identical results; a guess that cannot hold (rows sorted so
    that the tail is systematically better, in file order) is detected and redone."""
    from cmx import _lib

    rng = np.random.default_rng(23)
    X, Q = _aniso(rng, 150000, 128), _aniso(rng, 300, 128)
    sh = _shard(X)
    k = 1000  # guess after the dense slab: rank 3 * 1341 * 8192 / 150000 = 220 of the 8192 rows seen
    D, I = sh.search(Q, k, path="tensor")
    st = sh.last_stats()
    _lib.check(_lib.lib().cmx_debug_set_speculate(0))
    try:
        D0, I0 = sh.search(Q, k, path="tensor")
        st0 = sh.last_stats()
    finally:
        _lib.check(_lib.lib().cmx_debug_set_speculate(1))
    assert st["reruns"] == 0 and st0["reruns"] == 0
    assert st["slabs"] < st0["slabs"], (st, st0)
    assert np.array_equal(I, I0) and np.array_equal(D, D0)
    _check(D, I, X, Q, k)
    # file order + ascending scores: the guess made on the head of the file fails its verification
    u = Q[:1]
    Xs = X[np.argsort(X @ u[0])]
    sh2 = _shard(Xs)
    _lib.check(_lib.lib().cmx_debug_set_block_order(0))
    try:
        D2, I2 = sh2.search(Q[:40], k, path="tensor")
    finally:
        _lib.check(_lib.lib().cmx_debug_set_block_order(1))
    assert sh2.last_stats()["reruns"] >= 1
    _check(D2, I2, Xs, Q[:40], k)


def test_incremental_add_and_reconstruct(precision):
    rng = np.random.default_rng(20)
    X, Q = _unit(rng, 5000, 128), _unit(rng, 40, 128)
    from cmx.engine import Shard

    sh = Shard(128, 0)
    for a, b in ((0, 1), (1, 700), (700, 701), (701, 5000)):  # forces store growth
        sh.add(X[a:b])
    assert sh.ntotal == 5000
    assert np.array_equal(sh.reconstruct_n(0, 5000), X)
    D, I = sh.search(Q, 25, path="tensor")
    _check(D, I, X, Q, 25)
    sh.add(X[:300] * np.float32(64.0))  # larger magnitude -> operand planes are re-scaled
    X2 = np.concatenate([X, X[:300] * np.float32(64.0)])
    D, I = sh.search(Q, 25, path="tensor")
    _check(D, I, X2, Q, 25)
    D, I = sh.search(Q[:3], 25, path="stream")
    _check(D, I, X2, Q[:3], 25)


def test_device_tensors_zero_copy():
    import torch

    rng = np.random.default_rng(21)
    X, Q = _unit(rng, 8000, 128), _unit(rng, 150, 128)
    from cmx.engine import Shard

    sh = Shard(128, 0)
    sh.add(torch.from_numpy(X).cuda())
    Dd, Id = sh.search(torch.from_numpy(Q).cuda(), 30)
    assert Dd.is_cuda and Id.dtype == torch.int64
    Dh, Ih = sh.search(Q, 30)
    assert np.array_equal(Dd.cpu().numpy(), Dh) and np.array_equal(Id.cpu().numpy(), Ih)
    _check(Dh, Ih, X, Q, 30)


# ---------------------------------------------------------------- fused + merge + shim
def test_search_mixed_equals_mix_then_search():
    from cmx.engine import mix_normalize

    rng = np.random.default_rng(22)
    X = _aniso(rng, 25000, 256)
    P = _aniso(rng, 140, 256)
    S = _aniso(rng, 140, 256)
    alphas = [0, 0.1, 0.5, 1]
    sh = _shard(X)
    D, I, flags = sh.search_mixed(P, S, alphas, 100, want_flags=True)
    assert D.shape == (4, 140, 100) and not flags.any()
    Qm = mix_normalize(P, S, alphas)
    for ai in range(4):
        D1, I1 = sh.search(Qm[ai], 100)
        assert np.array_equal(D[ai], D1) and np.array_equal(I[ai], I1)
    Qo, _ = oracle.mix_normalize(P, S, alphas)
    for ai in range(4):
        _check(D[ai], I[ai], X, Qo[ai], 100)


def test_merge_and_sharded_equal_single(precision):
    from cmx.engine import merge_topk
    import cmx.faiss as faiss

    rng = np.random.default_rng(23)
    X, Q = _unit(rng, 21000, 128), _unit(rng, 77, 128)
    X[15000:15500] = X[100:600]  # ties across shard boundaries
    k = 64
    D1, I1 = _shard(X).search(Q, k, path="tensor")
    bounds = faiss.IndexShardsIP.split(X.shape[0], 3)
    parts = [_shard(X[a:b]).search(Q, k, id_base=a, path="tensor") for a, b in zip(bounds[:-1], bounds[1:])]
    Dp, Ip = np.stack([p[0] for p in parts]), np.stack([p[1] for p in parts])
    Dm, Im = merge_topk(Dp, Ip)
    assert np.array_equal(Im, I1) and np.array_equal(Dm, D1)
    Do, Io = oracle.merge_topk(Dp, Ip, k)
    assert np.array_equal(Im, Io) and np.array_equal(Dm, Do)
    # the single-process multi-GPU index (here: all shards on GPU 0) gives the same answer
    cpu = faiss.IndexFlatIP(128)
    cpu.add(X)
    sharded = faiss.index_cpu_to_gpus_list(cpu, gpus=[0, 0, 0])
    sharded.path = "tensor"
    Ds, Is = sharded.search(Q, k)
    assert np.array_equal(Is, I1) and np.array_equal(Ds, D1)


def test_two_gpu_shards_if_available():
    """Real multi-GPU (single process): shards on GPU 0 and 1, merged on GPU 0."""
    import torch
    import cmx.faiss as faiss

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    rng = np.random.default_rng(26)
    X, Q = _unit(rng, 30000, 128), _unit(rng, 200, 128)
    cpu = faiss.IndexFlatIP(128)
    cpu.add(X)
    D1, I1 = faiss.index_cpu_to_gpu(faiss.StandardGpuResources(), 0, cpu).search(Q, 100)
    sharded = faiss.index_cpu_to_all_gpus(cpu, ngpu=2)
    Ds, Is = sharded.search(Q, 100)
    assert np.array_equal(Is, I1) and np.array_equal(Ds, D1)
    _check(Ds, Is, X, Q, 100)


def test_faiss_shim_dropin_flow(tmp_path):
    """The call sequence of onepass_dense_mix_run_custom_lang.py on a synthetic cached index."""
    import cmx.faiss as faiss

    rng = np.random.default_rng(24)
    X, Q = _unit(rng, 6000, 64), _unit(rng, 50, 64)
    ids = np.arange(6000, dtype=np.int64)[::-1].copy() + 10_000
    cpu = faiss.IndexIDMap(faiss.IndexFlatIP(64))
    cpu.add_with_ids(X, ids)
    faiss.write_index(cpu, str(tmp_path / "index.faiss"))
    cached = faiss.read_index(str(tmp_path / "index.faiss"))
    base = faiss.downcast_index(cached.index)
    base.reconstruct(0, np.empty((64,), np.float32))
    res = faiss.StandardGpuResources()
    gpu = faiss.index_cpu_to_gpu(res, 0, cached)
    D, I = gpu.search(Q, 100)
    Dr, Ir = oracle.flat_ip_search(X, Q, 100, ids=ids)
    assert oracle.compare_topk(D, I, Dr, Ir, rtol=RTOL, atol=ATOL)["ok"]
    # the CPU index stays valid and independent; searching it runs on the GPU as well
    cached.add_with_ids(X[:10], np.arange(10))
    assert cached.ntotal == 6010 and gpu.ntotal == 6000
    D2, I2 = cached.search(Q[:4], 5)
    assert D2.shape == (4, 5)
    with pytest.raises(AssertionError):
        gpu.search(np.zeros((1, 65), np.float32), 3)
    with pytest.raises(RuntimeError):
        gpu.search(Q, 5000)  # k above the engine cap -> RuntimeError like faiss-gpu
    back = faiss.index_gpu_to_cpu(gpu)
    assert back.ntotal == 6000 and np.array_equal(back.index.reconstruct_n(0, 6000), X)
    # streamed straight into the device index (no host-resident copy): same answers
    direct = faiss.read_index_to_gpu(str(tmp_path / "index.faiss"), 0)
    assert direct.ntotal == 6000 and direct.index.getDevice() == 0
    D3, I3 = direct.search(Q, 100)
    assert np.array_equal(I3, I) and np.array_equal(D3, D)


def test_run_alpha_sweep_files(tmp_path):
    import cmx.faiss as faiss
    from cmx import runloop

    rng = np.random.default_rng(25)
    X = _aniso(rng, 9000, 128)
    P, S = _aniso(rng, 60, 128), _aniso(rng, 60, 128)
    qids = [str(1000 + 3 * i) for i in range(60)]
    lookup = {i: f"d{i * 11}" for i in range(9000)}
    idx = faiss.IndexIDMap(faiss.GpuIndexFlatIP(128))
    idx.add_with_ids(X, np.arange(9000))
    alphas = [0.0, 0.25, 0.5, 1.0]
    files = runloop.run_alpha_sweep(idx, lookup, qids, P, S, alphas, tmp_path / "mono", k=100, alpha_batch=2)
    assert [f.name for f in files] == ["cm-alpha-0.trec", "cm-alpha-0.25.trec", "cm-alpha-0.5.trec", "cm-alpha-1.trec"]
    from cmx.engine import mix_normalize

    Qo, _ = oracle.mix_normalize(P, S, alphas)
    Qg = mix_normalize(P, S, alphas)
    for ai, f in enumerate(files):
        D, I = idx.search(Qg[ai], 100)
        _check(D, I, X, Qo[ai], 100)
        # byte-identical to what the reference loop writes for this (D, I)
        assert f.read_text() == "\n".join(oracle.mono_trec_lines(qids, D, I, lookup))
    # bilingual: two derived ids per base document
    id2doc = [f"{i // 2}#{'en' if i % 2 == 0 else 'zh'}" for i in range(9000)]
    files = runloop.run_alpha_sweep_bilingual(idx, id2doc, qids, P, S, [0.5], tmp_path / "bi", topk=500,
                                              tag="bilingual-mix-en-zh")
    D, I = idx.search(Qg[2], 500)
    raw = oracle.bilingual_raw_lines(qids, D, I, id2doc, "bilingual-mix-en-zh")
    assert (tmp_path / "bi" / "cm-alpha-0.5_raw.trec").read_text() == "".join(raw)
    assert files[0].read_text() == oracle.collapse_run_max_text(raw)
    meta = json.loads((tmp_path / "bi" / "cm-alpha-0.5_meta.json").read_text())
    assert meta["topk"] == 500 and meta["alpha"] == "0.5"


# ---------------------------------------------------------------- larger-size properties
def test_large_properties_tensor_vs_stream_and_recompute(precision):
    """1M x 1024 on device: size-independent properties + agreement of the two paths."""
    import torch
    from cmx.engine import Shard

    d, N, nq, k = 1024, 1_000_000, 512, 1000
    g = torch.Generator(device="cuda").manual_seed(1234)
    sh = Shard(d, 0)
    sh.reserve(N)
    chunks = []
    for c in range(0, N, 1 << 18):
        x = torch.randn((min(1 << 18, N - c), d), generator=g, device="cuda")
        x = torch.nn.functional.normalize(x, dim=1)
        sh.add(x)
        chunks.append(x)
    X = torch.cat(chunks)
    del chunks
    Q = torch.nn.functional.normalize(torch.randn((nq, d), generator=g, device="cuda"), dim=1)
    D, I = sh.search(Q, k, path="tensor")
    assert sh.last_stats()["reruns"] == 0
    assert bool((D[:, 1:] <= D[:, :-1]).all())  # sorted descending
    assert int(I.min()) >= 0 and int(I.max()) < N
    srt = torch.sort(I, dim=1).values
    assert bool((srt[:, 1:] != srt[:, :-1]).all())  # no duplicate ids in a row
    # recompute the returned scores in fp64 on the device
    rec = torch.einsum("qkd,qd->qk", X[I[:64]].double(), Q[:64].double())
    assert float(((rec - D[:64].double()).abs() / rec.abs().clamp_min(1e-30)).max()) < RTOL
    # completeness: nothing outside the list beats the k-th score by more than the tolerance
    full = Q[:8].double() @ X.double().T
    kth = D[:8, -1].double()
    above = (full > (kth * (1 + RTOL) + ATOL)[:, None]).sum(dim=1)
    assert bool((above <= k).all()) and bool((above >= k - 5).all())
    # the fp32 streaming path agrees with the tensor path
    Ds, Is = sh.search(Q[:8], k, path="stream")
    rep = oracle.compare_topk(Ds.cpu().numpy(), Is.cpu().numpy(), D[:8].cpu().numpy(), I[:8].cpu().numpy(),
                              rtol=RTOL, atol=ATOL)
    assert rep["ok"], rep


# ---------------------------------------------------------------- limits and workspace reuse
def test_k_max_2048_and_over_limit(precision):
    rng = np.random.default_rng(40)
    X, Q = _unit(rng, 50000, 64), _unit(rng, 40, 64)
    sh = _shard(X)
    D, I = sh.search(Q, 2048, path="tensor")
    _check(D, I, X, Q, 2048)
    D, I = sh.search(Q[:2], 2048, path="stream")
    _check(D, I, X, Q[:2], 2048)
    with pytest.raises(RuntimeError):
        sh.search(Q, 2049)
    with pytest.raises(AssertionError):
        sh.search(Q, 0)


def test_query_chunking_above_8192_queries():
    """More queries than one corpus pass handles (8192): the passes are stitched correctly."""
    rng = np.random.default_rng(41)
    X = _unit(rng, 20000, 64)
    P, S = _unit(rng, 3100, 64), _unit(rng, 3100, 64)
    sh = _shard(X)
    D, I = sh.search_mixed(P, S, [0.0, 0.5, 1.0], 10)  # 9300 queries in one call
    Qo, _ = oracle.mix_normalize(P, S, [0.0, 0.5, 1.0])
    for ai in range(3):
        for rows in (slice(0, 64), slice(2660, 2760), slice(3036, 3100)):  # around the 8192 boundary too
            Dr, Ir = oracle.flat_ip_search(X, Qo[ai][rows], 10)
            assert oracle.compare_topk(D[ai][rows], I[ai][rows], Dr, Ir, rtol=RTOL, atol=ATOL)["ok"]


def test_workspace_reuse_reset_and_two_indexes():
    from cmx.engine import Shard

    rng = np.random.default_rng(42)
    Xa, Xb = _unit(rng, 7000, 64), _unit(rng, 9000, 128)
    a, b = Shard(64, 0), Shard(128, 0)
    a.add(Xa)
    b.add(Xb)
    for nq, k in ((300, 50), (3, 7), (1000, 300), (17, 1)):  # growing and shrinking workspaces
        Qa, Qb = _unit(rng, nq, 64), _unit(rng, nq, 128)
        Da, Ia = a.search(Qa, k)
        Db, Ib = b.search(Qb, k)
        _check(Da, Ia, Xa, Qa, k)
        _check(Db, Ib, Xb, Qb, k)
    a.reset()
    assert a.ntotal == 0
    D, I = a.search(_unit(rng, 2, 64), 3)
    assert (I == -1).all()
    a.add(Xa[:100])
    Q = _unit(rng, 20, 64)
    D, I = a.search(Q, 5)
    _check(D, I, Xa[:100], Q, 5)


def test_rescore_scores_are_exact_fp32_and_near_duplicates_fall_back():
    """rescore mode: D carries fp32 FMA scores (error ~1e-7, far inside the 1e-5 bar), and a corpus
    with more near-duplicates than the candidate buffer holds ends in the split fallback, still exact."""
    from cmx.engine import Shard

    rng = np.random.default_rng(43)
    X, Q = _unit(rng, 60000, 256), _unit(rng, 150, 256)
    sh = Shard(256, 0)
    sh.set_precision("rescore")
    sh.add(X)
    D, I = sh.search(Q, 100, path="tensor")
    Dt, It = oracle.flat_ip_search_f64(X, Q, 100)
    rep = oracle.compare_topk(D, I, Dt, It, rtol=RTOL, atol=ATOL)
    assert rep["ok"] and rep["max_abs_err"] < 5e-7, rep
    assert sh.last_stats()["reruns"] == 0
    # 12 000 copies of one row, jittered far below the fp16 resolution: all inside every margin band
    Xd = X.copy()
    Xd[20000:32000] = Xd[7] + rng.standard_normal((12000, 256)).astype(np.float32) * 1e-6
    qd = Xd[7:8] + 0.01 * _unit(rng, 1, 256)
    sh2 = Shard(256, 0)
    sh2.add(Xd)
    Qd = np.concatenate([qd, Q[:130]])
    D2, I2 = sh2.search(Qd, 50, path="tensor")
    assert sh2.last_stats()["reruns"] == 1  # rescore band overflow -> split precision, planned slabs
    _check(D2, I2, Xd, Qd, 50)


def test_union_kth_matches_numpy():
    """cmx_union_kth: global k-th best of several per-shard score lists (padding = lowest float)."""
    import torch
    from cmx.engine import union_kth

    rng = np.random.default_rng(77)
    nq, k, G = 37, 50, 3
    parts = rng.standard_normal((G, nq, k)).astype(np.float32)
    parts[1, 5, 10:] = np.finfo(np.float32).min  # a short list
    parts[:, 7, :] = np.finfo(np.float32).min     # nothing anywhere
    parts[:, 9, :] = 0.25                          # all equal
    t = [torch.from_numpy(parts[g].reshape(-1).copy()).cuda() for g in range(G)]
    outs = [torch.full((nq,), 123.0, dtype=torch.float32, device="cuda") for _ in range(2)]
    union_kth([x.data_ptr() for x in t], nq, k, 3, 30, [o.data_ptr() for o in outs], 0)
    torch.cuda.synchronize()
    flat = np.sort(parts.transpose(1, 0, 2).reshape(nq, G * k), axis=1)[:, ::-1]
    want = flat[:, k - 1]
    for o in outs:
        got = o.cpu().numpy()
        assert (got[:3] == 123.0).all() and (got[30:] == 123.0).all()  # only the slice is written
        sl = slice(3, 30)
        lowest = np.finfo(np.float32).min
        ok = (got[sl] == want[sl]) | ((want[sl] == lowest) & (got[sl] <= lowest))
        assert ok.all(), (got[sl][~ok], want[sl][~ok])


@pytest.mark.parametrize("G", [2, 5])
def test_two_phase_sharded_search_on_one_gpu(G):
    """The three-step sharded rescore search (begin / union_kth / end + merge), with the shards emulated
    as G indexes on one GPU: equals the single-index result exactly, and every shard rescored only its
    part of the global band."""
    import torch
    from cmx.engine import Shard, merge_topk, union_kth

    rng = np.random.default_rng(78)
    d, N, nq, k = 128, 30000, 64, 200
    X = _aniso(rng, N, d)
    X[1000:1040] *= 3.0  # unequal row norms across shards: the margin must use the GLOBAL maxima
    P, S = _aniso(rng, nq, d), _aniso(rng, nq, d)
    alphas = [0.0, 0.35, 1.0]
    one = Shard(d, 0)
    one.set_precision("rescore")
    one.add(X)
    Pt, St = torch.from_numpy(P).cuda(), torch.from_numpy(S).cuda()
    D1, I1 = one.search_mixed(Pt, St, alphas, k, path="tensor")
    bounds = np.linspace(0, N, G + 1).astype(int)
    shards = []
    for g in range(G):
        sh = Shard(d, 0)
        sh.set_precision("rescore")
        sh.add(X[bounds[g]:bounds[g + 1]])
        shards.append(sh)
    nqt = len(alphas) * nq
    # the asynchronous building blocks of include/cmx.h, driven from one stream (stream order = the barriers)
    bnd = [torch.zeros((2,), dtype=torch.float32, device="cuda") for _ in range(G)]
    flg = [torch.full((1,), 7, dtype=torch.int32, device="cuda") for _ in range(G)]
    flag_any = torch.zeros((1,), dtype=torch.int32, device="cuda")
    asc = [torch.empty((nqt * k,), dtype=torch.float32, device="cuda") for _ in range(G)]
    for g, sh in enumerate(shards):
        sh.export_bounds(bnd[g].data_ptr())
    qp = [sh.search_prepare(Pt, St, alphas) for sh in shards]
    for g, sh in enumerate(shards):
        sh.search_begin(qp[g], nqt, k, int(bounds[g]), [b.data_ptr() for b in bnd], 1.0 / G, asc[g].data_ptr(), flg[g].data_ptr())
    kth = torch.empty((nqt,), dtype=torch.float32, device="cuda")
    union_kth([a.data_ptr() for a in asc], nqt, k, 0, nqt, [kth.data_ptr()], 0, [f.data_ptr() for f in flg], flag_any.data_ptr())
    Dp = torch.empty((G, nqt, k), dtype=torch.float32, device="cuda")
    Ip = torch.empty((G, nqt, k), dtype=torch.int64, device="cuda")
    for g, sh in enumerate(shards):
        sh.search_end([kth.data_ptr()], Dp[g].data_ptr(), Ip[g].data_ptr())
    torch.cuda.synchronize()
    assert int(flag_any.item()) == 0 and all(int(f.item()) == 0 for f in flg)
    # the margins came from the GLOBAL maxima: every shard saw the 3x rows of shard 0
    assert max(float(b[0]) for b in bnd) > 2.5
    Dm, Im = merge_topk(Dp, Ip)
    assert torch.equal(Im.view(len(alphas), nq, k), I1) and torch.equal(Dm.view(len(alphas), nq, k), D1)
    # each shard returned only rows inside the global band: far fewer than k valid entries per query
    valid = (Ip >= 0).sum(dim=2).float().mean(dim=1).cpu().numpy()
    assert valid.sum() < 1.6 * k and (valid < 0.9 * k).all(), valid


def test_pinned_host_outputs_are_written_in_place(precision):
    """Page-locked (device-mapped) host output buffers are written by the final kernels directly -- no
    staged device-to-host copy; the result is bit-identical to the staged path (pageable numpy
    buffers, or the same pinned buffers with the switch off)."""
    import torch

    from cmx import _lib

    rng = np.random.default_rng(91)
    X, P, S = _unit(rng, 30000, 128), _unit(rng, 70, 128), _unit(rng, 70, 128)
    sh = _shard(X)
    alphas, k = [0.0, 0.4], 50
    D0, I0 = sh.search_mixed(P, S, alphas, k, path="tensor")  # numpy in, numpy out: staged copies
    Pp, Sp = torch.from_numpy(P).pin_memory(), torch.from_numpy(S).pin_memory()

    def pinned_run():
        Dp = torch.full((len(alphas), 70, k), 7.0, dtype=torch.float32).pin_memory()
        Ip = torch.full((len(alphas), 70, k), 7, dtype=torch.int64).pin_memory()
        sh.search_mixed(Pp, Sp, alphas, k, path="tensor", out=(Dp, Ip))
        return Dp.numpy().copy(), Ip.numpy().copy()

    D1, I1 = pinned_run()
    assert np.array_equal(D1, np.asarray(D0)) and np.array_equal(I1, np.asarray(I0))
    _lib.check(_lib.lib().cmx_debug_set_mapped_outputs(0))
    try:
        D2, I2 = pinned_run()
    finally:
        _lib.check(_lib.lib().cmx_debug_set_mapped_outputs(1))
    assert np.array_equal(D2, D1) and np.array_equal(I2, I1)
    # plain search through the same boundary
    Q = _unit(rng, 33, 128)
    Dq0, Iq0 = sh.search(Q, 20, path="tensor")
    Dq = torch.full((33, 20), 7.0, dtype=torch.float32).pin_memory()
    Iq = torch.full((33, 20), 7, dtype=torch.int64).pin_memory()
    sh.search(torch.from_numpy(Q).pin_memory(), 20, path="tensor", out=(Dq, Iq))
    assert np.array_equal(Dq.numpy(), np.asarray(Dq0)) and np.array_equal(Iq.numpy(), np.asarray(Iq0))
