"""The oracle against golden vectors produced by the reference's own code
(tests/golden/make_golden.py) and against an fp64 truth for the unpinned part."""
import json

import numpy as np

import oracle


def test_mix_matches_reference_safe_mix(golden_dir):
    g = np.load(golden_dir / "mix_golden.npz")
    Q, flags = oracle.mix_normalize(g["P"], g["S"], g["alphas"].tolist())
    ref = g["Q"]
    assert Q.shape == ref.shape
    # bit-exact, NaN-aware (fallback rows carry the caller's NaN/inf through)
    assert np.array_equal(Q.view(np.uint32), ref.view(np.uint32))
    assert flags.any(), "fixture must exercise the non-finite fallback"


def test_mix_f64_truth_close(golden_dir):
    g = np.load(golden_dir / "mix_golden.npz")
    a = g["alphas"].tolist()
    Q, flags = oracle.mix_normalize(g["P"], g["S"], a)
    T = oracle.mix_normalize_f64(g["P"], g["S"], a)
    ok = (flags == 0)[:, :, None] & np.isfinite(T) & np.isfinite(Q)
    ok[:, 10, :] = False  # the 1e20 row: fp32 norm overflows by design, fp64 does not
    ok[:, 3, :] = False  # P = -S: catastrophic cancellation, fp32 products round differently
    with np.errstate(all="ignore"):
        err = np.abs(np.where(ok, Q - T, 0.0))
    assert err.max() < 3e-7


def test_format_and_parse_alpha(golden_dir):
    g = json.loads((golden_dir / "text_golden.json").read_text())
    assert [oracle.format_alpha(a) for a in g["format_alpha"]["in"]] == g["format_alpha"]["out"]
    for s, want in zip(g["parse_alpha_list"]["in"], g["parse_alpha_list"]["out"]):
        try:
            got = oracle.parse_alpha_list(s)
        except SystemExit as exc:
            got = {"SystemExit": str(exc)}
        assert got == want


def test_trec_lines(golden_dir):
    g = json.loads((golden_dir / "text_golden.json").read_text())
    scores = np.array([np.frombuffer(bytes.fromhex(h), dtype=np.float32)[0] for h in g["trec_scores_f32_hex"]])
    n = len(scores)
    for i in range(n):
        qid, doc = str(1000 + i), 7000000 + 13 * i
        # one query whose i-th ranked hit is (doc, score): build rows that put it at rank i+1
        D = np.zeros((1, i + 1), np.float32)
        I = np.zeros((1, i + 1), np.int64)
        D[0, i], I[0, i] = scores[i], doc
        lines = oracle.mono_trec_lines([qid], D, I, {doc: str(doc)})
        assert lines[i] == g["mono_lines"][i]
        id2doc = [f"{doc}#en"]
        Ib = np.full((1, i + 1), -1, np.int64)
        Ib[0, i] = 0
        raw = oracle.bilingual_raw_lines([qid], D, Ib, id2doc, "bilingual-mix-en-zh")
        assert raw == [g["raw_lines"][i]]


def test_id_lookup_default():
    lines = oracle.mono_trec_lines(["q"], np.array([[0.5, -3.0]], np.float32), np.array([[4, -1]]), {4: "doc4"})
    assert lines == ["q\tQ0\tdoc4\t1\t0.5000\tonepass-cm", "q\tQ0\t-1\t2\t-3.0000\tonepass-cm"]


def test_collapse_run_max(golden_dir, tmp_path):
    g = json.loads((golden_dir / "text_golden.json").read_text())
    assert oracle.collapse_run_max_text(g["collapse"]["raw"]) == g["collapse"]["out"]
    pin, pout = tmp_path / "a_raw.trec", tmp_path / "a.trec"
    pin.write_text("".join(g["collapse"]["raw"]))
    oracle.collapse_run_max(pin, pout)
    assert pout.read_text() == g["collapse"]["out"]


def _unit(rng, n, d):
    x = rng.standard_normal((n, d)).astype(np.float32)
    return x / np.linalg.norm(x, axis=1, keepdims=True)


def test_search_against_f64_truth():
    rng = np.random.default_rng(0)
    X, Q = _unit(rng, 5000, 96), _unit(rng, 37, 96)
    D, I = oracle.flat_ip_search(X, Q, 50, block=1024)
    Dt, It = oracle.flat_ip_search_f64(X, Q, 50)
    rep = oracle.compare_topk(D, I, Dt, It)
    assert rep["ok"], rep
    assert np.all(np.diff(D, axis=1) <= 0)
    Df, If = oracle.flat_ip_search(X, Q, 50, block=1024, fast=True)
    assert oracle.compare_topk(Df, If, Dt, It)["ok"]


def test_search_padding_ids_and_ties():
    rng = np.random.default_rng(1)
    X = _unit(rng, 7, 16)
    X[5] = X[2]  # duplicate rows -> exact tie, lower row first
    Q = _unit(rng, 3, 16)
    ids = np.arange(100, 107, dtype=np.int64) * 3
    D, I = oracle.flat_ip_search(X, Q, 10, ids=ids)
    assert (I[:, 7:] == -1).all() and (D[:, 7:] == np.finfo(np.float32).min).all()
    for r in range(3):
        row = I[r, :7].tolist()
        assert row.index(ids[2]) + 1 == row.index(ids[5])
    D0, I0 = oracle.flat_ip_search(np.zeros((0, 16), np.float32), Q, 4)
    assert (I0 == -1).all()


def test_merge_topk_equals_single():
    rng = np.random.default_rng(2)
    X, Q = _unit(rng, 3000, 32), _unit(rng, 11, 32)
    D, I = oracle.flat_ip_search(X, Q, 20)
    parts = [oracle.flat_ip_search(X[a:b], Q, 20, ids=np.arange(a, b)) for a, b in ((0, 1000), (1000, 1700), (1700, 3000))]
    Dm, Im = oracle.merge_topk(np.stack([p[0] for p in parts]), np.stack([p[1] for p in parts]), 20)
    assert np.array_equal(Im, I) and np.array_equal(Dm, D)


def test_compare_topk_flags_real_errors():
    rng = np.random.default_rng(3)
    X, Q = _unit(rng, 500, 32), _unit(rng, 5, 32)
    D, I = oracle.flat_ip_search(X, Q, 10)
    I2 = I.copy()
    I2[0, 0] = I[0, 9]
    assert not oracle.compare_topk(D, I2, D, I)["ok"]
    D2 = D.copy()
    D2[1, 3] *= 1.001
    assert not oracle.compare_topk(D2, I, D, I)["ok"]
