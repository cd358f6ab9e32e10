"""GPU tests of the round-2 machinery: exact scores ahead of time (prescoring), speculative mid slabs,
the asynchronous two-phase building blocks, pinned shared host results, the one-process multi-GPU index
and the index.faiss reader on a device sink."""
import ctypes

import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu
RTOL, ATOL = 1e-5, 1e-6


def _unit(rng, n, d):
    x = rng.standard_normal((n, d)).astype(np.float32)
    return x / np.linalg.norm(x, axis=1, keepdims=True)


def _unit_cuda(n, d, seed):
    import torch

    g = torch.Generator(device="cuda").manual_seed(seed)
    return torch.nn.functional.normalize(torch.randn((n, d), generator=g, device="cuda"), dim=1)


def _brute(X, Q, k):
    """fp32 brute force on the GPU (torch, TF32 off), ties by ascending row via a stable sort."""
    import torch

    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        s = Q @ X.T
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev
    v, i = torch.topk(s, k, dim=1)
    return v.cpu().numpy(), i.cpu().numpy()


def test_prescoring_is_invisible_in_the_result():
    """A last slab long enough to be cut into several scoring launches: the candidates found early get
    their exact scores beside the next launch (prescore_kernel on a second stream).  Same (D, I), bit for
    bit, as with all exact scores computed after the last tile -- and as the fp32 brute force."""
    from cmx import _lib
    from cmx.engine import Shard

    N, d, nq, k = 1_000_000, 128, 300, 1000
    X, Q = _unit_cuda(N, d, 11), _unit_cuda(nq, d, 12)
    sh = Shard(d, 0)
    sh.add(X)
    L = _lib.lib()
    _lib.check(L.cmx_debug_set_prescore(1))                # off by default (profiles/r02_prescore_experiments.md)
    _lib.check(L.cmx_debug_set_prescore_min_rows(0))       # and only beside last slabs of >= 2 M rows
    _lib.check(L.cmx_debug_set_prescore_params(1.5, 0, 8))  # deep: most of the final top-k is prescored
    try:
        D1, I1 = sh.search(Q, k, path="tensor")
        st1 = sh.last_stats()
        _lib.check(L.cmx_debug_set_prescore_params(0.75, 0, 8))
        Dd, Id = sh.search(Q, k, path="tensor")
        _lib.check(L.cmx_debug_set_prescore(0))
        D0, I0 = sh.search(Q, k, path="tensor")
        st0 = sh.last_stats()
    finally:
        _lib.check(L.cmx_debug_set_prescore(0))
        _lib.check(L.cmx_debug_set_prescore_params(0.75, 0, 8))
        _lib.check(L.cmx_debug_set_prescore_min_rows(-1))
    assert st1["reruns"] == 0 and st1["score_launches"] > st1["slabs"], st1  # the last slab ran as several launches
    assert st0["score_launches"] == st0["slabs"]
    import torch

    assert torch.equal(I1, I0) and torch.equal(D1, D0) and torch.equal(Id, I0) and torch.equal(Dd, D0)
    Dr, Ir = _brute(X, Q, k)
    rep = oracle.compare_topk(D1.cpu().numpy(), I1.cpu().numpy(), Dr, Ir, rtol=RTOL, atol=ATOL)
    assert rep["ok"], rep


def test_speculative_mid_slab_equals_planned_slabs():
    """k = 1000 over 3 M rows: dense slab, ONE speculative mid slab, speculative final slab (3 launches
    groups) -- identical to the geometric plan's result."""
    import torch

    from cmx import _lib
    from cmx.engine import Shard

    N, d, nq, k = 3_000_000, 64, 200, 1000
    X, Q = _unit_cuda(N, d, 21), _unit_cuda(nq, d, 22)
    sh = Shard(d, 0)
    sh.add(X)
    D1, I1 = sh.search(Q, k, path="tensor")
    st1 = sh.last_stats()
    assert st1["slabs"] == 3 and st1["reruns"] == 0, st1
    _lib.check(_lib.lib().cmx_debug_set_speculate(0))
    try:
        D0, I0 = sh.search(Q, k, path="tensor")
        st0 = sh.last_stats()
    finally:
        _lib.check(_lib.lib().cmx_debug_set_speculate(1))
    assert st0["slabs"] > st1["slabs"] and st0["reruns"] == 0
    assert torch.equal(I1, I0) and torch.equal(D1, D0)


def test_pending_two_phase_state_is_cancelled_by_other_calls():
    import torch

    from cmx.engine import Shard

    rng = np.random.default_rng(5)
    X, Q = _unit(rng, 20000, 64), _unit(rng, 40, 64)
    sh = Shard(64, 0)
    sh.add(X)
    Qt = torch.from_numpy(Q).cuda()
    bnd = torch.zeros((2,), dtype=torch.float32, device="cuda")
    flag = torch.zeros((1,), dtype=torch.int32, device="cuda")
    asc = torch.empty((40 * 10,), dtype=torch.float32, device="cuda")
    D = torch.empty((40, 10), dtype=torch.float32, device="cuda")
    I = torch.empty((40, 10), dtype=torch.int64, device="cuda")
    with pytest.raises(RuntimeError, match="pending"):
        sh.search_end([], D.data_ptr(), I.data_ptr())
    sh.export_bounds(bnd.data_ptr())
    sh.search_begin(Qt.data_ptr(), 40, 10, 0, [bnd.data_ptr()], 1.0, asc.data_ptr(), flag.data_ptr())
    sh.search(Q[:3], 5)  # reuses the workspace: the pending state must not survive
    with pytest.raises(RuntimeError, match="pending"):
        sh.search_end([], D.data_ptr(), I.data_ptr())
    sh.search_begin(Qt.data_ptr(), 40, 10, 0, [bnd.data_ptr()], 1.0, asc.data_ptr(), flag.data_ptr())
    sh.add(X[:10])
    with pytest.raises(RuntimeError, match="pending"):
        sh.search_end([], D.data_ptr(), I.data_ptr())
    # a complete begin / end with no global cut == the plain search
    sh.search_begin(Qt.data_ptr(), 40, 10, 0, [bnd.data_ptr()], 1.0, asc.data_ptr(), flag.data_ptr())
    sh.search_end([], D.data_ptr(), I.data_ptr())
    torch.cuda.synchronize()
    D1, I1 = sh.search(Qt, 10, path="tensor")
    assert int(flag.item()) == 0 and torch.equal(I, I1) and torch.equal(D, D1)


def test_empty_and_unusable_shards_in_two_phase():
    """An empty shard takes part in a two-phase step (all-padding lists); a shard that cannot run the one-pass
    arithmetic raises no error but sets its status word, so that all shards fall back together."""
    import torch

    from cmx.engine import Shard, union_kth

    rng = np.random.default_rng(6)
    X, Q = _unit(rng, 9000, 64), _unit(rng, 17, 64)
    Qt = torch.from_numpy(Q).cuda()
    k = 20
    full, empty, split = Shard(64, 0), Shard(64, 0), Shard(64, 0)
    full.add(X)
    split.add(X[:100])
    split.set_precision("split")
    shards = [full, empty, split]
    bnd = [torch.zeros((2,), dtype=torch.float32, device="cuda") for _ in shards]
    flg = [torch.zeros((1,), dtype=torch.int32, device="cuda") for _ in shards]
    asc = [torch.empty((17 * k,), dtype=torch.float32, device="cuda") for _ in shards]
    for sh, b in zip(shards, bnd):
        sh.export_bounds(b.data_ptr())
    for g, sh in enumerate(shards):
        sh.search_begin(Qt.data_ptr(), 17, k, 0, [b.data_ptr() for b in bnd], 1.0 / 3, asc[g].data_ptr(), flg[g].data_ptr())
    flag_any = torch.zeros((1,), dtype=torch.int32, device="cuda")
    kth = torch.empty((17,), dtype=torch.float32, device="cuda")
    union_kth([a.data_ptr() for a in asc[:2]], 17, k, 0, 17, [kth.data_ptr()], 0, [f.data_ptr() for f in flg], flag_any.data_ptr())
    D = torch.empty((3, 17, k), dtype=torch.float32, device="cuda")
    I = torch.empty((3, 17, k), dtype=torch.int64, device="cuda")
    for g, sh in enumerate(shards):
        sh.search_end([kth.data_ptr()], D[g].data_ptr(), I[g].data_ptr())
    torch.cuda.synchronize()
    assert int(flg[0].item()) == 0 and int(flg[1].item()) == 0 and int(flg[2].item()) == 0x100
    assert int(flag_any.item()) == 0x100
    assert bool((I[1] == -1).all()) and bool((asc[1] == np.finfo(np.float32).min).all())
    D1, I1 = full.search(Qt, k, path="tensor")
    assert torch.equal(I[0], I1) and torch.equal(D[0], D1)  # shards 0 + 1 alone: the cut is shard 0's own k-th


def test_finite_rows_with_overflowing_norm_use_split_precision():
    """||x|| > 1.8e19 overflows the fp32 sum of squares of the error bound: the index reports an infinite
    bound and searches in split precision instead of silently underestimating the margin."""
    from cmx.engine import Shard

    rng = np.random.default_rng(8)
    X = _unit(rng, 5000, 64) * np.float32(3e19)
    Q = _unit(rng, 40, 64)
    sh = Shard(64, 0)
    sh.set_precision("rescore")
    sh.add(X)
    nb, _ = sh.error_bounds()
    assert np.isinf(nb)
    D, I = sh.search(Q, 10, path="tensor")
    Dr, Ir = oracle.flat_ip_search(X, Q, 10)
    assert oracle.compare_topk(D, I, Dr, Ir, rtol=RTOL, atol=ATOL)["ok"]


def test_shards_index_one_process(tmp_path):
    """cmx.faiss.IndexShardsIP runs the SAME two-phase step as the torchrun path, one thread per shard over a
    LocalFabric -- here three shards on one GPU (any GPU count >= 1): equals the single-index result bit for
    bit, for a batch larger than one two-phase chunk too, and for host results in alternating pinned buffers."""
    import cmx.faiss as faiss
    from cmx.engine import Shard

    rng = np.random.default_rng(9)
    N, d, nq, k = 60000, 128, 150, 100
    X, P, S = _unit(rng, N, d), _unit(rng, nq, d), _unit(rng, nq, d)
    X[40000:40300] = X[200:500]  # ties across shards
    one = Shard(d, 0)
    one.add(X)
    alphas = [0.0, 0.3, 1.0]
    D1, I1 = one.search_mixed(P, S, alphas, k, path="tensor")
    cpu = faiss.IndexFlatIP(d)
    cpu.add(X)
    sharded = faiss.index_cpu_to_gpus_list(cpu, gpus=[0, 0, 0])
    Ds, Is = sharded.search_mixed(P, S, alphas, k)
    assert sharded.two_phase_used
    assert np.array_equal(Is, I1) and np.array_equal(Ds, D1)
    Ds2, Is2 = sharded.search_mixed(P, S, alphas, k)  # the other pinned buffer; the first result is still intact
    assert Ds2.ctypes.data != Ds.ctypes.data and np.array_equal(Is, I1) and np.array_equal(Is2, I1)
    Dq, Iq = sharded.search(P, k)
    assert np.array_equal(Iq, I1[0]) and np.array_equal(Dq, D1[0])
    De, Ie = sharded.search(P[:0], k)  # no queries: nothing to do, no collective
    assert De.shape == (0, k) and Ie.shape == (0, k)
    D2q, I2q = sharded.search(P[:2], k)  # fewer queries than shards: some merge slices are empty
    assert np.array_equal(I2q, I1[0][:2]) and np.array_equal(D2q, D1[0][:2])
    # 60 alphas x 150 queries = 9000 > 8192: two chunks of the two-phase pass
    many = [i / 59.0 for i in range(60)]
    Dm, Im = sharded.search_mixed(P, S, many, 10)
    D0, I0 = one.search_mixed(P, S, many, 10, path="tensor")
    assert np.array_equal(Im, I0) and np.array_equal(Dm, D0)
    assert all(v.fallback_steps == 0 and v.fallback_chunks == 0 for v in sharded._views)
    # one shard reports a buffer overflow in one chunk: every shard redoes that chunk only, same answer
    from cmx import _lib

    _lib.check(_lib.lib().cmx_debug_inject_begin_status(4, 1))  # the 5th begin of the 6 (3 shards x 2 chunks)
    Dm2, Im2 = sharded.search_mixed(P, S, many, 10)
    assert np.array_equal(Im2, I0) and np.array_equal(Dm2, D0)
    assert all(v.fallback_chunks == 1 and v.fallback_steps == 0 and v.last_status == 1 for v in sharded._views)


def test_two_gpu_shards_one_process_if_available():
    """Real multi-GPU, one process: peer access + event barriers between the devices' streams."""
    import torch
    import cmx.faiss as faiss

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    rng = np.random.default_rng(26)
    N, d, nq, k = 400_000, 128, 500, 1000
    X, P, S = _unit(rng, N, d), _unit(rng, nq, d), _unit(rng, nq, d)
    cpu = faiss.IndexFlatIP(d)
    cpu.add(X)
    one = faiss.index_cpu_to_gpu(faiss.StandardGpuResources(), 0, cpu)
    D1, I1 = one.search_mixed(P, S, [0.5, 0.9], k)
    sharded = faiss.index_cpu_to_all_gpus(cpu, ngpu=min(4, torch.cuda.device_count()))
    Ds, Is = sharded.search_mixed(P, S, [0.5, 0.9], k)
    assert sharded.two_phase_used
    assert np.array_equal(Is, np.asarray(I1)) and np.array_equal(Ds, np.asarray(D1))


def test_repeated_runs_are_bit_identical_across_kernel_variants():
    """compute-sanitizer (racecheck / synccheck) is closed on this GPU pool, so the hand-rolled mbarrier / TMEM
    pipelines are checked the way a race would show up: the same search repeated under every scheduling variant --
    static and dynamic tile order, single-CTA and CTA-pair scorer, tile widths 256 / 128, prescoring on / off, both
    small-batch widths -- must give bit-identical (D, I) every time."""
    import torch

    from cmx import _lib
    from cmx.engine import Shard

    L = _lib.lib()
    N, d, k = 600_000, 256, 500
    X = _unit_cuda(N, d, 51)
    sh = Shard(d, 0)
    sh.add(X)
    for nq in (700, 24):
        Q = _unit_cuda(nq, d, 52 + nq)
        ref = None
        variants = [dict(), dict(flags=512), dict(pair=1), dict(tile=128), dict(prescore=1), dict(flags=512, prescore=1)]
        try:
            for rep in range(3):
                for v in variants:
                    _lib.check(L.cmx_debug_set_tensor_flags(v.get("flags", 0)))
                    _lib.check(L.cmx_debug_set_tensor_pair(v.get("pair", -1)))
                    _lib.check(L.cmx_debug_set_tensor_tile(v.get("tile", 256)))
                    _lib.check(L.cmx_debug_set_prescore(v.get("prescore", 0)))
                    _lib.check(L.cmx_debug_set_prescore_min_rows(0 if v.get("prescore") else -1))
                    D, I = sh.search(Q, k, path="tensor")
                    assert sh.last_stats()["reruns"] == 0
                    if ref is None:
                        ref = (D.clone(), I.clone())
                    else:
                        assert torch.equal(I, ref[1]) and torch.equal(D, ref[0]), (nq, rep, v)
        finally:
            _lib.check(L.cmx_debug_set_tensor_flags(0))
            _lib.check(L.cmx_debug_set_tensor_pair(-1))
            _lib.check(L.cmx_debug_set_tensor_tile(256))
            _lib.check(L.cmx_debug_set_prescore(0))
            _lib.check(L.cmx_debug_set_prescore_min_rows(-1))
        Dr, Ir = _brute(X, Q, k)
        assert oracle.compare_topk(ref[0].cpu().numpy(), ref[1].cpu().numpy(), Dr, Ir, rtol=RTOL, atol=ATOL)["ok"]


def test_early_slab_epilogue_paths_agree():
    """The pool epilogue of the early slabs (csrc/tc_score.cu) has three regimes: everything fits the 512-entry warp
    pool (a few percent pass: the sample-slab plan), the pool fills up inside a tile and is emptied in place (a third
    of all scores pass: a 4096-entry candidate buffer), and a single 32-column chunk overflows the empty pool (half of
    all scores pass: split precision with a buffer of 2k entries).  Every regime, with and without speculation and
    with either first-slab rule, must return the (D, I) of the default plan of its precision bit for bit, and that
    must be the exact top-k.  (scripts/exp_pool_paths.py counts the paths with the -DCMX_TC_TIMERS build on exactly
    these inputs -- profiles/r02_pool_epilogue_paths.jsonl: 0 / 1295 / 1549 pools emptied in place, 58 chunks appended
    lane by lane in the split / 2048 case.)"""
    import torch

    from cmx import _lib
    from cmx.engine import Shard

    L = _lib.lib()
    N, d, k, nq = 400_000, 256, 1000, 300
    X = _unit_cuda(N, d, 71)
    Q = _unit_cuda(nq, d, 72)
    Dr, Ir = _brute(X, Q, k)
    for precision, caps in (("rescore", (0, 4096)), ("split", (0, 2048))):
        sh = Shard(d, 0)
        sh.set_precision(precision)
        sh.add(X)
        ref = None
        plans = set()
        try:
            for cap in caps:  # 0: default capacity (8192)
                for small_first in (1, 0):
                    for spec in (1, 0):
                        _lib.check(L.cmx_debug_set_small_first(small_first))
                        _lib.check(L.cmx_debug_set_speculate(spec))
                        sh.set_cand_capacity(cap)
                        D, I = sh.search(Q, k, path="tensor")
                        st = sh.last_stats()
                        assert st["reruns"] == 0, (precision, cap, small_first, spec, st)
                        plans.add(st["slabs"])
                        if ref is None:
                            ref = (D.clone(), I.clone())
                        else:
                            assert torch.equal(I, ref[1]) and torch.equal(D, ref[0]), (precision, cap, small_first, spec, st)
        finally:
            _lib.check(L.cmx_debug_set_small_first(1))
            _lib.check(L.cmx_debug_set_speculate(1))
        assert len(plans) >= 3, plans  # the settings really produced different slab plans
        assert oracle.compare_topk(ref[0].cpu().numpy(), ref[1].cpu().numpy(), Dr, Ir, rtol=RTOL, atol=ATOL)["ok"], precision


def test_large_pageable_add_is_staged_and_exact():
    """add() of a large pageable host array (index_cpu_to_gpu of a read_index'ed corpus) goes through the page-locked
    staging pipeline (parallel memcpy + overlapped H2D); the stored rows and the corpus statistics are the same as for
    the direct copy of a device tensor."""
    import torch

    from cmx.engine import Shard

    rng = np.random.default_rng(61)
    n, d = 300_000, 256  # 307 MB > the 256 MB switch, several 64 MB chunks with a ragged last one
    X = rng.standard_normal((n, d)).astype(np.float32)
    X[12345] *= 7.0
    a, b = Shard(d, 0), Shard(d, 0)
    a.add(X)                          # pageable numpy: staged
    b.add(torch.from_numpy(X).cuda())  # device tensor: one D2D copy
    assert a.ntotal == b.ntotal == n
    assert a.error_bounds() == b.error_bounds()
    for i0 in (0, 65535, 131072, n - 1000):
        assert np.array_equal(a.reconstruct_n(i0, 1000), X[i0:i0 + 1000])
    Q = _unit(rng, 50, d)
    Da, Ia = a.search(Q, 20)
    Db, Ib = b.search(Q, 20)
    assert np.array_equal(Ia, Ib) and np.array_equal(Da, Db)
    a.add(X[:10])                     # small adds keep the direct path and append after the staged rows
    assert a.ntotal == n + 10 and np.array_equal(a.reconstruct_n(n, 10), X[:10])


def test_mixed_query_norms_in_one_batch_stay_exact():
    """The query planes of a batch share ONE power-of-two scale: a large-norm query makes the fp16 residual -- and so
    the margin -- of a small-norm query in the same batch relatively wide.  That may cost candidates (or, past half a
    buffer, the fall-back to split precision) but never exactness: every query gets the same (D, I) as when it is
    searched alone."""
    from cmx.engine import Shard

    rng = np.random.default_rng(71)
    X = _unit(rng, 120_000, 128)
    Q = _unit(rng, 96, 128)
    Q *= np.float32(10.0) ** rng.uniform(-3, 3, size=(96, 1)).astype(np.float32)  # norms over six decades
    sh = Shard(128, 0)
    sh.add(X)
    k = 100
    D, I = sh.search(Q, k, path="tensor")
    st = sh.last_stats()
    Dr, Ir = oracle.flat_ip_search(X, Q, k)
    rep = oracle.compare_topk(D, I, Dr, Ir, rtol=RTOL, atol=0.0)  # purely relative: scores span 1e-4 .. 1e2
    assert rep["ok"], (rep, st)
    for i in (int(np.argmin(np.linalg.norm(Q, axis=1))), int(np.argmax(np.linalg.norm(Q, axis=1))), 17):
        Di, Ii = sh.search(Q[i:i + 1], k, path="tensor")
        if st["reruns"] == 0:  # both ran the one-pass arithmetic + exact rescoring: bit-identical
            assert np.array_equal(Ii[0], I[i]) and np.array_equal(Di[0], D[i])
        else:                  # the batch fell back to split precision: equal within the parity tolerance
            assert oracle.compare_topk(Di, Ii, D[i:i + 1], I[i:i + 1], rtol=RTOL, atol=0.0)["ok"]


def test_device_collapse_by_base_id(golden_dir, tmp_path):
    """cmx_collapse_max: grouping by base id with fuse = max on the GPU -- equal to its numpy restatement,
    and the run files written from it byte-identical to the reference's text round trip (oracle + the golden fixture
    produced by the reference's own collapse_run_max)."""
    import json

    import torch
    from conftest import collapse_groups_ref

    from cmx import runloop

    rng = np.random.default_rng(81)
    nq, k, nrows = 300, 500, 40_000
    D = -np.sort(-rng.uniform(0.2, 0.9, (nq, k)).astype(np.float32), axis=1)
    I = rng.integers(0, nrows, (nq, k)).astype(np.int64)
    I[:, 1::3] = (I[:, 0::3][:, : I[:, 1::3].shape[1]] + nrows // 2) % nrows  # the other language of the hit before: same base
    D[:, 10] = D[:, 9]                                   # exact ties
    D[:, 21] = D[:, 20] - np.float32(3e-8)               # ties after 6-decimal rounding
    D[7, 300:] = -np.abs(D[7, 300:])                     # negative scores sort last
    I[2, 5], I[4, 8] = -1, nrows + 3                     # skipped hits
    id2doc = [f"{i % (nrows // 2)}#{'en' if i < nrows // 2 else 'zh'}" for i in range(nrows)]
    bt = runloop.BaseTable(id2doc)
    codes_dev = torch.from_numpy(bt.codes).cuda()
    want = collapse_groups_ref(D, I, bt.codes, nrows)
    for Dt, It in ((torch.from_numpy(D).cuda(), torch.from_numpy(I).cuda()),           # device results (one GPU)
                   (torch.from_numpy(D).pin_memory(), torch.from_numpy(I).pin_memory())):  # pinned host results (shards)
        got = runloop.collapse_max_device(Dt, It, codes_dev, nrows)
        assert got is not None
        assert np.array_equal(got[2], want[2])
        for r in range(nq):
            n = int(want[2][r])
            assert np.array_equal(got[0][r, :n], want[0][r, :n]) and np.array_equal(got[1][r, :n], want[1][r, :n])
    qids = [str(100 + i) for i in range(nq)]
    tag = "bilingual-mix-en-zh"
    runloop.write_bilingual_trec(tmp_path / "raw.trec", tmp_path / "col.trec", qids, D, I, bt, tag, groups=got)
    raw_lines = oracle.bilingual_raw_lines(qids, D, I, id2doc, tag)
    assert (tmp_path / "raw.trec").read_text() == "".join(raw_lines)
    assert (tmp_path / "col.trec").read_text() == oracle.collapse_run_max_text(raw_lines)
    # scores the 64-bit keys cannot carry exactly are reported, not mangled
    Dbad = D.copy()
    Dbad[0, 0] = np.float32(3e7)
    assert runloop.collapse_max_device(torch.from_numpy(Dbad).cuda(), torch.from_numpy(I).cuda(), codes_dev, nrows) is None
    Dbad[0, 0] = np.float32(-0.0)
    assert runloop.collapse_max_device(torch.from_numpy(Dbad).cuda(), torch.from_numpy(I).cuda(), codes_dev, nrows) is None
    # the golden fixture: raw lines in, the reference's own collapsed text out
    g = json.loads((golden_dir / "text_golden.json").read_text())["collapse"]
    rows = [l.split() for l in g["raw"]]
    gq = sorted({r[0] for r in rows}, key=lambda x: int(x[1:]))
    dids = sorted({r[2] for r in rows})
    bt2 = runloop.BaseTable(dids)
    kk = max(int(r[3]) for r in rows)
    Dg = np.zeros((len(gq), kk), np.float32)
    Ig = np.full((len(gq), kk), -1, np.int64)
    for qid, _, did, rank, sc, _ in rows:
        Dg[gq.index(qid), int(rank) - 1] = np.float32(sc)
        Ig[gq.index(qid), int(rank) - 1] = dids.index(did)
    grp = runloop.collapse_max_device(torch.from_numpy(Dg).cuda(), torch.from_numpy(Ig).cuda(), torch.from_numpy(bt2.codes).cuda(), len(dids))
    runloop.write_bilingual_trec(tmp_path / "graw.trec", tmp_path / "gcol.trec", gq, Dg, Ig, bt2, "bilingual-mix-en-zh", groups=grp)
    assert (tmp_path / "gcol.trec").read_text() == g["out"]


def test_memory_accounting_is_bounded():
    """Default precision = fp32 store + ONE fp16 plane = 1.5x a FAISS flat index (+ a workspace that does not grow
    with the corpus); the split precision adds the second plane (2x)."""
    from cmx.engine import Shard

    N, d = 200_000, 1024
    X = _unit_cuda(N, d, 31)
    Q = _unit_cuda(700, d, 32)
    sh = Shard(d, 0)
    sh.reserve(N)
    sh.add(X)
    m0 = sh.memory()
    assert m0["store"] == m0["faiss_flat"] == N * d * 4 and m0["planes"] == 0  # planes are built by the first tensor-path search
    sh.search(Q, 1000, path="tensor")
    m1 = sh.memory()
    assert m1["planes"] == N * d * 2 and (m1["store"] + m1["planes"]) == 1.5 * m1["faiss_flat"]
    ws = m1["workspace"]
    assert 700 * 8192 * 8 <= ws <= 700 * 8192 * 8 + 64 * (1 << 20)  # candidate buffers dominate: [nq, 8192] keys
    sh.set_precision("split")
    sh.search(Q, 1000, path="tensor")
    m2 = sh.memory()
    assert m2["planes"] == 2 * N * d * 2 and m2["workspace"] <= ws + 16 * (1 << 20)


def test_index_fixture_streams_into_a_gpu_index(tmp_path):
    """read_index_to_gpu on an index.faiss assembled with bare struct.pack (tests/test_host_logic.py)."""
    import cmx.faiss as faiss
    from test_host_logic import _faiss18_fixture

    rng = np.random.default_rng(41)
    n, d = 3000, 64
    x = _unit(rng, n, d)
    ids = (rng.permutation(100_000)[:n] + 5).astype(np.int64)
    blob, _ = _faiss18_fixture(x, ids)
    path = tmp_path / "index.faiss"
    path.write_bytes(blob)
    gidx = faiss.read_index_to_gpu(str(path), device=0)
    assert gidx.ntotal == n and np.array_equal(gidx.id_map, ids)
    assert np.array_equal(gidx.index.reconstruct_n(0, n), x)
    Q = _unit(rng, 9, d)
    D, I = gidx.search(Q, 7)
    Dr, Ir = oracle.flat_ip_search(x, Q, 7, ids=ids)
    assert oracle.compare_topk(D, I, Dr, Ir, rtol=RTOL, atol=ATOL)["ok"]
