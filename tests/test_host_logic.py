"""Host-side logic (no GPU): labels, vectorised TREC text, collapse, file formats,
and that libcmx.so loads and exports every symbol include/cmx.h declares."""
import ctypes
import json
import pathlib
import re

import numpy as np
import pytest

import oracle
from cmx import runloop
from cmx import io as cio

ROOT = pathlib.Path(__file__).resolve().parent.parent


def test_labels_match_golden(golden_dir):
    g = json.loads((golden_dir / "text_golden.json").read_text())
    assert [runloop.format_alpha(a) for a in g["format_alpha"]["in"]] == g["format_alpha"]["out"]
    for s, want in zip(g["parse_alpha_list"]["in"], g["parse_alpha_list"]["out"]):
        try:
            got = runloop.parse_alpha_list(s)
        except SystemExit as exc:
            got = {"SystemExit": str(exc)}
        assert got == want


def test_cli_helpers_match_golden(golden_dir, monkeypatch):
    """sanitize_tag / default_query_cache_root / parse_query_specs (rewritten here) against outputs of the
    reference's own functions (tests/golden/make_golden.py --cli): cache directory names and parser messages
    are what a user of the reference sees."""
    import pathlib as pl

    from cmx import cli

    g = json.loads((golden_dir / "cli_golden.json").read_text())
    assert [cli.sanitize_tag(t) for t in g["sanitize_tag"]["in"]] == g["sanitize_tag"]["out"]
    for case in g["default_query_cache_root"]:
        for k in ("QUERY_CACHE_ROOT", "QUERY_CACHE_ROOT_BASE"):
            monkeypatch.delenv(k, raising=False)
        for k, v in case["env"].items():
            monkeypatch.setenv(k, v)
        got = cli.default_query_cache_root(case["repo"], case["encoder"])
        if "leaf" in case:
            assert got.name == case["leaf"] and got.parent == pl.Path.cwd() / "data"
        else:
            assert str(got) == case["path"]
    for args, want in zip(g["parse_query_specs"]["in"], g["parse_query_specs"]["out"]):
        try:
            got = [[lang, str(path)] for lang, path in cli.parse_query_specs(*args)]
        except SystemExit as exc:
            got = {"SystemExit": str(exc)}
        assert got == want, (args, got, want)


def test_batch_sorted_order_is_the_reference_row_order():
    """combine_cached_indexes: map lines in batches of 20 000, each batch sorted by int id (stable)."""
    from cmx import cli

    rng = np.random.default_rng(3)
    lid = rng.permutation(50_000).astype(np.int64)
    lid[100] = lid[7]  # a repeated id keeps file order inside its batch
    order = cli._batch_sorted_order(lid)
    want = []
    for b0 in range(0, len(lid), cli.RECON_BATCH):
        batch = sorted(range(b0, min(b0 + cli.RECON_BATCH, len(lid))), key=lambda i: lid[i])
        want.extend(batch)
    assert order.tolist() == want
    assert cli._batch_sorted_order(np.empty((0,), np.int64)).shape == (0,)


def _fake_results(rng, nq, k, nrows):
    D = -np.sort(-rng.uniform(-0.2, 0.9, (nq, k)).astype(np.float32), axis=1)
    I = rng.integers(0, nrows, (nq, k)).astype(np.int64)
    return D, I


def test_mono_text_equals_oracle_lines():
    rng = np.random.default_rng(6)
    D, I = _fake_results(rng, 13, 50, 400)
    I[3, 40:] = -1
    D[3, 40:] = np.finfo(np.float32).min
    I[5, 7] = 399  # not in lookup -> str(id)
    lookup = {i: f"doc{i * 7}" for i in range(399)}
    qids = [str(100 + i) for i in range(13)]
    want = "\n".join(oracle.mono_trec_lines(qids, D, I, lookup))
    assert runloop.mono_trec_bytes(qids, D, I, runloop.DocTable(lookup)).decode("utf-8") == want


def test_c_formatter_matches_python_on_arbitrary_float32():
    """cmx_trec_mono / cmx_trec_bilingual print scores exactly like f"{x:.4f}" / f"{x:.6f}" for ANY float32:
    random bit patterns over all exponents, halfway cases, -0.0, subnormals, huge values, inf and nan."""
    rng = np.random.default_rng(99)
    bits = rng.integers(0, 2 ** 32, 40000, dtype=np.uint64).astype(np.uint32)
    vals = bits.view(np.float32).copy()
    halfway = (rng.integers(-20000, 20000, 4000) / 20000.0 + 0.00005).astype(np.float32)  # x.xxxx5 at 4 decimals
    halfway6 = (rng.integers(-2000000, 2000000, 4000) / 2000000.0 + 0.0000005).astype(np.float32)
    special = np.array([0.0, -0.0, 1e-45, -1e-45, 1.17549435e-38, 3.4028235e38, -3.4028235e38, 9.9999997e10, 1.0e11,
                        -1.0e11, 123456.789, np.inf, -np.inf, np.nan, 0.99995, 0.999995, -0.00005, 2.5e-5, 3.5e-5], np.float32)
    allv = np.concatenate([vals, halfway, halfway6, special])
    k = 97
    pad = (-len(allv)) % k
    allv = np.concatenate([allv, np.zeros(pad, np.float32)])
    D = allv.reshape(-1, k)
    nq = D.shape[0]
    I = np.tile(np.arange(k, dtype=np.int64), (nq, 1))
    qids = [str(i) for i in range(nq)]
    lookup = {i: f"d{i}" for i in range(k)}
    want = "\n".join(oracle.mono_trec_lines(qids, D, I, lookup))
    assert runloop.mono_trec_bytes(qids, D, I, lookup).decode("utf-8") == want
    id2doc = [f"{i}#en" for i in range(k)]  # every hit its own base: the collapsed file carries each rounded score
    raw_lines = oracle.bilingual_raw_lines(qids, D, I, id2doc, "t")
    raw, col = runloop.bilingual_bytes(qids, D, I, id2doc, "t")
    assert raw.decode("utf-8") == "".join(raw_lines)
    finite_rows = np.isfinite(D).all(axis=1)  # the reference's collapse re-parses text: float("nan") ordering is its own story
    sub = np.nonzero(finite_rows)[0][:150]
    qs = [qids[i] for i in sub]
    raw_sub = oracle.bilingual_raw_lines(qs, D[sub], I[sub], id2doc, "t")
    assert runloop.bilingual_bytes(qs, D[sub], I[sub], id2doc, "t")[1].decode("utf-8") == oracle.collapse_run_max_text(raw_sub)


def test_bilingual_raw_and_collapse_equal_oracle():
    rng = np.random.default_rng(7)
    nq, k, nrows = 9, 60, 80
    D, I = _fake_results(rng, nq, k, nrows)
    D[:, 10] = D[:, 9]  # exact ties
    D[:, 21] = D[:, 20] - np.float32(3e-8)  # ties after 6-decimal rounding
    I[2, 5] = -1
    I[4, 8] = nrows + 3  # out of range -> skipped, rank kept
    id2doc = [f"{i // 2}#{'en' if i % 2 == 0 else 'zh'}" for i in range(nrows)]
    qids = [f"q{i}" for i in range(nq)]
    tag = "bilingual-mix-en-zh"
    raw_lines = oracle.bilingual_raw_lines(qids, D, I, id2doc, tag)
    assert runloop.bilingual_bytes(qids, D, I, id2doc, tag)[0].decode("utf-8") == "".join(raw_lines)
    want = oracle.collapse_run_max_text(raw_lines)
    assert runloop.bilingual_bytes(qids, D, I, id2doc, "x")[1].decode("utf-8") == want


def test_bilingual_writer_with_precomputed_groups(tmp_path):
    """cmx_trec_bilingual_file_pre: the collapsed run written from group lists (what cmx_collapse_max produces on
    the device; here its numpy restatement) is byte-identical to the host grouping and to the oracle's text round trip."""
    from conftest import collapse_groups_ref

    rng = np.random.default_rng(8)
    nq, k, nrows = 11, 70, 90
    D, I = _fake_results(rng, nq, k, nrows)
    D[:, 10] = D[:, 9]
    D[:, 21] = D[:, 20] - np.float32(3e-8)
    D[3, 40:] = -np.abs(D[3, 40:])  # negative scores
    I[2, 5] = -1
    I[4, 8] = nrows + 3
    id2doc = [f"{i // 2}#{'en' if i % 2 == 0 else 'zh'}" for i in range(nrows)]
    qids = [f"q{i}" for i in range(nq)]
    tag = "bilingual-mix-en-zh"
    bt = runloop.BaseTable(id2doc)
    groups = collapse_groups_ref(D, I, bt.codes, nrows)
    runloop.write_bilingual_trec(tmp_path / "raw_a.trec", tmp_path / "col_a.trec", qids, D, I, bt, tag)
    runloop.write_bilingual_trec(tmp_path / "raw_b.trec", tmp_path / "col_b.trec", qids, D, I, bt, tag, groups=groups)
    assert (tmp_path / "raw_a.trec").read_bytes() == (tmp_path / "raw_b.trec").read_bytes()
    assert (tmp_path / "col_a.trec").read_bytes() == (tmp_path / "col_b.trec").read_bytes()
    raw_lines = oracle.bilingual_raw_lines(qids, D, I, id2doc, tag)
    assert (tmp_path / "col_b.trec").read_text() == oracle.collapse_run_max_text(raw_lines)
    bad = (groups[0].copy(), groups[1], groups[2])
    bad[0][0, 0] = 10**6  # a base code outside the table is refused, not dereferenced
    with pytest.raises(RuntimeError):
        runloop.write_bilingual_trec(tmp_path / "r.trec", tmp_path / "c.trec", qids, D, I, bt, tag, groups=bad)


class _OracleIndex:
    """Stands in for a cmx.faiss index in the sweep loops (CPU): search_mixed through the oracle."""

    def __init__(self, X):
        self.X, self.ntotal, self.d = X, X.shape[0], X.shape[1]

    def search_mixed(self, P, S, alphas, k):
        Q, _ = oracle.mix_normalize(P, S, list(alphas))
        out = [oracle.flat_ip_search(self.X, Q[a], k) for a in range(len(alphas))]
        return np.stack([o[0] for o in out]), np.stack([o[1] for o in out])


def test_file_writers_and_pipelined_sweeps_equal_oracle_text(tmp_path):
    """cmx_trec_mono_file / cmx_trec_bilingual_file (parallel pwrite, no assembled text) and the
    pipelined sweep loops write exactly the bytes the reference loops would."""
    rng = np.random.default_rng(61)
    n, d, nq, k = 3000, 32, 157, 40  # more queries than formatter threads, ragged thread ranges
    X = rng.standard_normal((n, d)).astype(np.float32)
    X /= np.linalg.norm(X, axis=1, keepdims=True)
    P = rng.standard_normal((nq, d)).astype(np.float32)
    S = rng.standard_normal((nq, d)).astype(np.float32)
    qids = [str(9000 + 3 * i) for i in range(nq)]
    lookup = {i: f"D{i:07d}" for i in range(n - 5)}  # the last ids print as numbers
    alphas = [0.0, 0.25, 0.5, 1.0]
    idx = _OracleIndex(X)
    files = runloop.run_alpha_sweep(idx, lookup, qids, P, S, alphas, tmp_path / "mono", k=k)
    assert [f.name for f in files] == ["cm-alpha-0.trec", "cm-alpha-0.25.trec", "cm-alpha-0.5.trec", "cm-alpha-1.trec"]
    for a, f in zip(alphas, files):
        D, I = idx.search_mixed(P, S, [a], k)
        want = "\n".join(oracle.mono_trec_lines(qids, D[0], I[0], lookup))
        assert f.read_bytes() == want.encode("utf-8")
        assert runloop.mono_trec_bytes(qids, D[0], I[0], lookup) == want.encode("utf-8")
    assert not list((tmp_path / "mono").glob("*.tmp*"))
    # one thread / many threads give the same file
    sz = runloop.write_mono_trec(tmp_path / "one.trec", qids, D[0], I[0], lookup, nthreads=1)
    assert (tmp_path / "one.trec").read_bytes() == files[-1].read_bytes() and sz == files[-1].stat().st_size
    runloop.write_mono_trec(tmp_path / "many.trec", qids, D[0], I[0], lookup, nthreads=64)
    assert (tmp_path / "many.trec").read_bytes() == files[-1].read_bytes()
    # bilingual: raw + collapsed
    id2doc = [f"{i // 2}#{'en' if i % 2 == 0 else 'zh'}" for i in range(n)]
    bfiles = runloop.run_alpha_sweep_bilingual(idx, id2doc, qids, P, S, [0.5, 0.75], tmp_path / "bi", topk=k, tag="bilingual-mix-en-zh")
    for a, f in zip([0.5, 0.75], bfiles):
        D, I = idx.search_mixed(P, S, [a], k)
        raw_lines = oracle.bilingual_raw_lines(qids, D[0], I[0], id2doc, "bilingual-mix-en-zh")
        raw = f.with_name(f.stem + "_raw.trec")
        assert raw.read_text() == "".join(raw_lines)
        assert f.read_text() == oracle.collapse_run_max_text(raw_lines)
        meta = json.loads(f.with_name(f.stem + "_meta.json").read_text())
        assert meta["topk"] == k and meta["index"]["size"] == n
    # empty result set
    assert runloop.write_mono_trec(tmp_path / "empty.trec", [], np.zeros((0, 5), np.float32), np.zeros((0, 5), np.int64), lookup) == 0
    assert (tmp_path / "empty.trec").read_bytes() == b""


def test_base_table_arrow_path_equals_python_path():
    rng = np.random.default_rng(8)
    names = [f"{int(v)}#{'en' if i % 2 else 'zh'}" for i, v in enumerate(rng.integers(0, 3000, 9000))]
    names += ["plain", "a#b#c", "#x", "", "\u6587\u6863#zh", "a#b#c"]

    class PyOnly(runloop.BaseTable):
        def _init_arrow(self, names):
            return False

    fast, slow = runloop.BaseTable(names), PyOnly(names)
    assert fast._init_arrow(names)  # the arrow path is available here and was taken
    for a, b in ((fast.docs, slow.docs), (fast.bases, slow.bases)):
        assert a.n == b.n and a.buf == b.buf and np.array_equal(a.off, b.off)
    assert fast.codes.dtype == np.int32 and np.array_equal(fast.codes, slow.codes)


def test_collapse_text_form_matches_golden(golden_dir, tmp_path):
    g = json.loads((golden_dir / "text_golden.json").read_text())
    pin, pout = tmp_path / "x_raw.trec", tmp_path / "x.trec"
    pin.write_text("".join(g["collapse"]["raw"]))
    runloop.collapse_run_max(pin, pout)
    assert pout.read_text() == g["collapse"]["out"]


def test_query_cache_roundtrip(tmp_path):
    rng = np.random.default_rng(8)
    qids = [str(i) for i in (5, 3, 11)]
    vecs = rng.standard_normal((3, 8)).astype(np.float32)
    cio.save_query_cache(tmp_path, "en", qids, vecs)
    assert np.array_equal(cio.load_query_cache(tmp_path, "en", qids), vecs)
    assert cio.load_query_cache(tmp_path, "en", qids[::-1]) is None  # order matters
    assert cio.load_query_cache(tmp_path, "zh", qids) is None
    cio.save_query_cache(tmp_path, "zh", qids, {q: v for q, v in zip(qids, vecs)})
    assert np.array_equal(cio.load_query_cache(tmp_path, "zh", qids), vecs)


def test_docid_map_roundtrip(tmp_path):
    p = tmp_path / "docid_map.tsv"
    cio.write_docid_map(p, [0, 1, 2], ["7#en", "9#en", "4#en"], ["7", "9", "4"], "en")
    with open(p, "a") as fh:
        fh.write("bad\tline\n")
        fh.write("x\ty\tz\tw\n")
    lookup, kept, derived = cio.read_docid_map(p)
    assert lookup == {0: "7", 1: "9", 2: "4"} and kept == ["7", "9", "4"] and derived == ["7#en", "9#en", "4#en"]


def test_docid_table_fast_path_equals_literal_reader(tmp_path):
    """read_docid_table (pyarrow fast path with strict preconditions) == read_docid_map (the reference's
    per-line reader) on clean maps, and falls back to it on anything unusual."""
    rng = np.random.default_rng(31)

    def fast(f):
        try:
            return cio._docid_table_arrow(f, runloop.DocTable, runloop.StrTable) is not None
        except Exception:
            return False

    def check(name, lines, expect_keys_none=None, expect_fast=None):
        f = tmp_path / f"{name}.tsv"
        f.write_bytes("".join(lines).encode("utf-8"))
        if expect_fast is not None:
            assert fast(f) == expect_fast, name
        id_lookup, kept, _ = cio.read_docid_map(f)
        docs, text, n = cio.read_docid_table(f)
        assert n == len(kept) and text == "\n".join(sorted(set(kept)))
        probe = list(id_lookup.keys())[:50] + [-1, 10 ** 9] + [int(v) for v in rng.integers(0, 3000, 50)]
        assert docs.lookup(probe) == [id_lookup.get(i, str(i)) for i in probe]
        if expect_keys_none is not None:
            assert (docs.keys is None) == expect_keys_none
        return docs

    hdr = "int_id\tderived_id\tbase_id\tlang\n"
    n = 2000
    clean = [hdr] + [f"{i}\t{i * 3}#en\t{i * 3}\ten\n" for i in range(n)]
    check("clean", clean, True, expect_fast=True)
    check("no_trailing_newline", clean[:-1] + [clean[-1].rstrip("\n")], True)
    perm = rng.permutation(n)
    check("shuffled", [hdr] + [f"{i}\t{i}#zh\tdoc-{i % 700}\tzh\n" for i in perm], True, expect_fast=True)  # repeated base ids
    check("sparse", [hdr] + [f"{i * 5 + 2}\tx\tb{i}\ten\n" for i in range(500)], False, expect_fast=True)
    check("unicode", [hdr] + [f"{i}\td\t\u6587\u6863{i}\u00e9\tzh\n" for i in range(300)], True)
    check("three_columns", ["int_id\tderived_id\tbase_id\n"] + [f"{i}\td{i}\tb{i}\n" for i in range(100)], True)
    # fallbacks: each of these must take the literal reader (and still agree with it)
    check("dup_ids", expect_fast=False, lines=[hdr, "1\ta\tfirst\ten\n", "2\tb\tsecond\ten\n", "1\tc\tlast\ten\n"])
    check("short_line", expect_fast=False, lines=[hdr, "0\ta\tb0\ten\n", "garbage\n", "1\ta\tb1\ten\n"])
    check("bad_int", expect_fast=False, lines=[hdr, "0\ta\tb0\ten\n", "x7\ta\tb7\ten\n", " 2 \ta\tb2\ten\n", "+3\ta\tb3\ten\n"])
    check("crlf", expect_fast=False, lines=[hdr.replace("\n", "\r\n"), "0\ta\tb0\ten\r\n", "1\ta\tb1\ten\r\n"])
    check("quotes", [hdr, '0\ta\t"b0\ten\n', "1\ta\tb'1\ten\n"], True)
    check("empty", expect_fast=False, lines=[hdr])


def test_index_file_roundtrip(tmp_path):
    import cmx.faiss as faiss

    rng = np.random.default_rng(9)
    x = rng.standard_normal((37, 12)).astype(np.float32)
    ids = rng.permutation(1000)[:37].astype(np.int64)
    idx = faiss.IndexIDMap(faiss.IndexFlatIP(12))
    idx.add_with_ids(x, ids)
    faiss.write_index(idx, tmp_path / "index.faiss")
    raw = (tmp_path / "index.faiss").read_bytes()
    assert raw[:4] == b"IxMp" and raw[4 + 33: 4 + 37] == b"IxFI"
    assert len(raw) == 2 * (4 + 33) + 8 + 37 * 12 * 4 + 8 + 37 * 8
    back = faiss.read_index(str(tmp_path / "index.faiss"))
    assert back.d == 12 and back.ntotal == 37
    assert np.array_equal(back.id_map, ids)
    base = faiss.downcast_index(back.index)
    tmp = np.empty((12,), np.float32)
    base.reconstruct(0, tmp)
    assert np.array_equal(tmp, x[0])
    assert np.array_equal(base.reconstruct_n(0, 37), x)
    with pytest.raises(RuntimeError):
        back.reconstruct(0)
    flat = faiss.IndexFlatIP(12)
    flat.add(x)
    faiss.write_index(flat, tmp_path / "flat.faiss")
    assert np.array_equal(faiss.read_index(tmp_path / "flat.faiss").reconstruct_n(0, 37), x)
    (tmp_path / "bad.faiss").write_bytes(b"IxXX" + raw[4:])
    with pytest.raises(RuntimeError):
        faiss.read_index(tmp_path / "bad.faiss")
    (tmp_path / "trunc.faiss").write_bytes(raw[:200])
    with pytest.raises(RuntimeError):
        faiss.read_index(tmp_path / "trunc.faiss")


def _faiss18_fixture(x, ids):
    """An ``index.faiss`` assembled byte by byte with bare ``struct.pack`` in the field order of faiss 1.8's
    index_write.cpp (write_index_header + the IxFI / IxMp branches) -- independent of cmx.io's writer:
        fourcc "IxMp" | int32 d | int64 ntotal | int64 dummy | int64 dummy | uint8 is_trained | int32 metric_type
        fourcc "IxFI" | <same header> | uint64 n_floats | n_floats x float32 (WRITEVECTOR(codes) in float units)
        uint64 n_ids | n_ids x int64 (WRITEVECTOR(id_map))"""
    import struct

    n, d = x.shape

    def header():
        return struct.pack("<i", d) + struct.pack("<q", n) + struct.pack("<q", 1 << 20) + struct.pack("<q", 1 << 20) \
            + struct.pack("<B", 1) + struct.pack("<i", 0)

    flat = b"IxFI" + header() + struct.pack("<Q", n * d) + b"".join(struct.pack("<f", float(v)) for v in x.reshape(-1))
    return b"IxMp" + header() + flat + struct.pack("<Q", n) + b"".join(struct.pack("<q", int(v)) for v in ids), flat


def test_index_reader_against_independent_byte_fixture(tmp_path):
    """read_index / inspect_index / stream_index_rows on a file this repository's writer never touched
    (faiss is not installable here, so this cannot be a faiss-written file: parity for the format stays
    'unpinned', but reader and writer no longer vouch for each other)."""
    import cmx.faiss as faiss

    rng = np.random.default_rng(31)
    n, d = 53, 20
    x = rng.standard_normal((n, d)).astype(np.float32)
    ids = (rng.permutation(10_000)[:n] * 7 + 3).astype(np.int64)  # a non-identity, non-monotonic id_map
    blob, flat_blob = _faiss18_fixture(x, ids)
    path = tmp_path / "index.faiss"
    path.write_bytes(blob)
    info = cio.inspect_index(path)
    assert info == {"kind": "IxMp", "d": d, "ntotal": n, "vec_offset": 4 + 33 + 4 + 33 + 8,
                    "ids_offset": 4 + 33 + 4 + 33 + 8 + 4 * n * d + 8}
    idx = faiss.read_index(str(path))
    assert isinstance(idx, faiss.IndexIDMap) and idx.d == d and idx.ntotal == n
    assert np.array_equal(idx.id_map, ids)
    base = faiss.downcast_index(idx.index)
    assert np.array_equal(base.reconstruct_n(0, n), x)
    sink = faiss.IndexFlatIP(d)
    assert cio.stream_index_rows(path, sink, 5, 40, chunk_rows=8).tolist() == ids[5:40].tolist()
    assert np.array_equal(sink.reconstruct_n(0, 35), x[5:40])
    # the nested flat index alone is a valid IxFI file
    (tmp_path / "flat.faiss").write_bytes(flat_blob)
    flat = faiss.read_index(tmp_path / "flat.faiss")
    assert isinstance(flat, faiss.IndexFlatIP) and np.array_equal(flat.reconstruct_n(0, n), x)
    # and the writer produces exactly these bytes
    faiss.write_index(idx, tmp_path / "rewritten.faiss")
    assert (tmp_path / "rewritten.faiss").read_bytes() == blob


def test_index_file_validator_and_streaming_reader(tmp_path):
    """inspect_index checks the whole layout without loading vectors; stream_index_rows feeds any row
    range to a sink (a GPU index or one rank's shard in production, a host index here)."""
    import cmx.faiss as faiss

    rng = np.random.default_rng(12)
    n, d = 1000, 24
    X = rng.standard_normal((n, d)).astype(np.float32)
    ids = rng.permutation(5000)[:n].astype(np.int64)
    idx = faiss.IndexIDMap(faiss.IndexFlatIP(d))
    idx.add_with_ids(X, ids)
    path = tmp_path / "index.faiss"
    faiss.write_index(idx, str(path))
    info = cio.inspect_index(path)
    assert info["kind"] == "IxMp" and (info["d"], info["ntotal"]) == (d, n)
    for r0, r1, chunk in ((0, n, 64), (0, n, 4096), (137, 611, 100), (999, 1000, 7), (500, 500, 8)):
        sink = faiss.IndexFlatIP(d)
        got = cio.stream_index_rows(path, sink, r0, r1, chunk_rows=chunk)
        assert got.tolist() == ids[r0:r1].tolist() and sink.ntotal == r1 - r0
        if r1 > r0:
            assert np.array_equal(sink.reconstruct_n(0, r1 - r0), X[r0:r1])
    with pytest.raises(RuntimeError):
        cio.stream_index_rows(path, faiss.IndexFlatIP(d), 10, n + 1)
    flat_path = tmp_path / "flat.faiss"
    faiss.write_index(idx.index, str(flat_path))
    sink = faiss.IndexFlatIP(d)
    assert cio.inspect_index(flat_path)["kind"] == "IxFI"
    assert cio.stream_index_rows(flat_path, sink, 3, 9).tolist() == [3, 4, 5, 6, 7, 8]
    # corrupted files are rejected before a single vector is read
    raw = path.read_bytes()
    bad = {"truncated": raw[:-5], "trailing": raw + b"\0" * 8, "fourcc": b"IxF2" + raw[4:],
           "count": raw[:4 + 33 + 4 + 33] + (n * d + 1).to_bytes(8, "little") + raw[4 + 33 + 4 + 33 + 8:],
           "outer_ntotal": raw[:8] + (n + 1).to_bytes(8, "little") + raw[16:],
           "metric": raw[:4 + 29] + (1).to_bytes(4, "little") + raw[4 + 33:]}
    for name, blob in bad.items():
        f = tmp_path / f"bad_{name}.faiss"
        f.write_bytes(blob)
        with pytest.raises(RuntimeError):
            cio.inspect_index(f)


def test_faiss_shim_host_semantics():
    import cmx.faiss as faiss

    assert hasattr(faiss, "StandardGpuResources")  # the reference's feature probe
    idx = faiss.IndexIDMap(faiss.IndexFlatIP(4))
    with pytest.raises(AssertionError):
        idx.add_with_ids(np.zeros((2, 5), np.float32), np.arange(2))
    with pytest.raises(AssertionError):
        idx.add_with_ids(np.zeros((2, 4), np.float32), np.arange(3))
    x = np.arange(8, dtype=np.float64).reshape(2, 4)  # coerced to float32, copied
    idx.add_with_ids(x, [10, 20])
    x[:] = -1
    assert idx.ntotal == 2 and idx.index.reconstruct(1).tolist() == [4.0, 5.0, 6.0, 7.0]
    with pytest.raises(RuntimeError):
        idx.index.reconstruct(2)
    with pytest.raises(RuntimeError):
        faiss.IndexIDMap(idx.index)
    assert faiss.IndexShardsIP.split(10, 4) == [0, 2, 5, 7, 10]
    # IndexIDMap id translation: -1 stays -1; the identity map (ids 0..n-1 in row order) is recognised
    I = np.array([[1, 0, -1], [0, -1, -1]], dtype=np.int64)
    assert idx._translate(I).tolist() == [[20, 10, -1], [10, -1, -1]]
    idx.add_with_ids(np.ones((1, 4), np.float32), [7])  # the cached map follows later adds
    assert idx._translate(np.array([[2, 1, -1]])).tolist() == [[7, 20, -1]]
    ident = faiss.IndexIDMap(faiss.IndexFlatIP(4))
    ident.add_with_ids(np.zeros((3, 4), np.float32), np.arange(3))
    assert ident._translate(I) is I
    ident.add_with_ids(np.zeros((2, 4), np.float32), [3, 9])  # no longer the identity
    assert ident._translate(np.array([[4, 3, 0, -1]])).tolist() == [[9, 3, 0, -1]]
    ident.reset()
    ident.add_with_ids(np.zeros((2, 4), np.float32), [5, 6])
    assert ident._translate(np.array([[1, 0]])).tolist() == [[6, 5]]


def test_library_exports_every_declared_symbol():
    from cmx import _lib

    header = (ROOT / "include" / "cmx.h").read_text()
    declared = sorted(set(re.findall(r"CMX_API\s+[\w\s\*]+?\b(cmx_\w+)\s*\(", header)))
    assert len(declared) >= 20, declared
    assert _lib.LIB_PATH.exists(), "libcmx.so not built (python codemix-dense-retrieval_b200/build.py)"
    L = ctypes.CDLL(str(_lib.LIB_PATH))
    missing = [s for s in declared if not hasattr(L, s)]
    assert not missing, missing
    _lib.lib()  # prototypes bind
    assert L.cmx_version() >= 100


def _plan(n, k, cap, rescore=1, safe=0, spec=1):
    from cmx import _lib

    rows = (ctypes.c_int64 * 4096)()
    ns, ss, sr = ctypes.c_int(0), ctypes.c_int(0), ctypes.c_int(0)
    _lib.check(_lib.lib().cmx_debug_plan_slabs(n, k, cap, rescore, safe, spec, rows, 4096, ctypes.byref(ns), ctypes.byref(ss),
                                               ctypes.byref(sr)))
    return [int(rows[i]) for i in range(min(ns.value, 4096))], ns.value, ss.value, sr.value


def _plan_ranks(n, k, cap, rescore=1, safe=0, spec=1, nslabs=16):
    from cmx import _lib

    ranks = (ctypes.c_int * nslabs)()
    _lib.check(_lib.lib().cmx_debug_plan_ranks(n, k, cap, rescore, safe, spec, ranks, nslabs))
    return [int(v) for v in ranks]


def test_slab_schedule_invariants():
    """Host logic of the tensor path's slab schedule: whole 256-row blocks, dense first slab,
    geometric growth, speculative slabs only when their rank estimate is trustworthy, safe slabs bounded."""
    from cmx import _lib

    n = 8_841_823
    npad = (n + 255) // 256 * 256
    # C2: a dense SAMPLE slab (the smallest first slab that still gives a 3-launch plan), ONE speculative mid slab up
    # to the row count its rank-80 guess can serve (k' * f / 32), then the rest of the corpus under a second guess
    rows, ns, ss, sr = _plan(n, 1000, 8192)
    f = rows[0]
    assert ns == 3 and ss == 2 and 2048 <= f < 8192 and f % 256 == 0
    mid = 1341 * f // 32 - f
    assert rows == [f, mid // 256 * 256, npad - f - mid // 256 * 256]
    assert _plan_ranks(n, 1000, 8192)[:3] == [0, 80, sr]
    seen = sum(rows[:2])
    r0 = 1341 * seen / npad
    assert r0 >= 32 and sr == int(np.ceil(2.5 * r0)) and sr < 750
    # ... and f - 256 would not do: the final guess would rest on a rank below 32
    f2 = f - 256
    assert 1341 * (1341 * f2 // 32 // 256 * 256) / npad < 32
    # the first slab fills the buffers when it is asked to (experiments) or when nothing is speculative
    _lib.check(_lib.lib().cmx_debug_set_small_first(0))
    try:
        rows_b, ns_b, ss_b, sr_b = _plan(n, 1000, 8192)
        mid_b = 1341 * 8192 // 32 - 8192
        assert rows_b == [8192, mid_b // 256 * 256, npad - 8192 - mid_b // 256 * 256] and ns_b == 3 and ss_b == 2
        rows8b, ns8b, ss8b, _ = _plan(1_105_228, 1000, 8192)
        assert ns8b == 3 and ss8b == 2 and rows8b[:2] == [8192, 20736]
    finally:
        _lib.check(_lib.lib().cmx_debug_set_small_first(1))
    rows_g, ns_g, ss_g, _ = _plan(n, 1000, 8192, spec=0)
    assert ss_g == -1 and ns_g == 7 and rows_g[:4] == [8192, 20736, 73728, 262144] and sum(rows_g) == npad
    for a, b in zip(rows_g[1:-1], rows_g[2:-1]):
        assert 3.0 < b / a < 3.7  # x(1 + (C - k') / 2k') per slab
    # shards of a 2-, 4- and 8-GPU search: 3 launches as well, the sample slab shrinks with the shard
    firsts = [f]
    for g in (2, 4, 8):
        rows_s, ns_s, ss_s, sr_s = _plan(n // g, 1000, 8192)
        assert ns_s == 3 and ss_s == 2 and 80 < sr_s < 750 and rows_s[1] == (1341 * rows_s[0] // 32 - rows_s[0]) // 256 * 256
        assert _plan_ranks(n // g, 1000, 8192)[:2] == [0, 80]
        firsts.append(rows_s[0])
    assert firsts == sorted(firsts, reverse=True) and firsts[-1] == 2048
    rows8, ns8, ss8, sr8 = _plan(1_105_228, 1000, 8192)
    # small k: the geometric plan is already 3-4 slabs and the rank estimate never qualifies
    assert _plan(n, 100, 8192)[2] == -1 and _plan(n, 10, 8192)[2] == -1
    # split precision plans on k itself
    rows_s, _, ss_s, sr_s = _plan(n, 1000, 8192, rescore=0)
    assert sum(rows_s) == npad and ss_s >= 1 and sr_s < 1000
    for r in (rows, rows_g, rows8, rows_s):
        assert all(v % 256 == 0 and v > 0 for v in r) and 2048 <= r[0] <= 8192 and sum(r) == ((sum(r) + 255) // 256) * 256
    # worst-case-safe schedule: no slab larger than the free room; tiny buffers get pieces of one block
    rows_safe, ns_safe, ss_safe, _ = _plan(100_000, 1000, 8192, rescore=0, safe=1)
    assert ss_safe == -1 and max(rows_safe[1:]) <= 8192 - 1000 and sum(rows_safe) == (100_000 + 255) // 256 * 256
    rows_tiny, _, _, _ = _plan(3000, 100, 256, rescore=0, safe=1)
    assert rows_tiny[0] == 256 and max(rows_tiny[1:]) <= 156 and sum(rows_tiny) == 3072
    pos = 256
    for v in rows_tiny[1:]:
        assert pos // 256 == (pos + v - 1) // 256  # never straddles a block
        pos += v


def test_block_order_is_a_well_spread_permutation():
    """pick_perm: position j -> block (j * P) mod nblk is a permutation, and every prefix is spread evenly
    over the corpus (what makes thresholds learnt on the first slabs valid for all rows)."""
    from math import gcd

    from cmx import _lib

    L = _lib.lib()
    for nblk in (1, 2, 3, 7, 20, 36, 157, 4096, 34539, 69077):
        P = int(L.cmx_debug_block_perm(nblk))
        assert 1 <= P < max(nblk, 2) and gcd(P, nblk) == 1
        if nblk > 4096:
            continue
        order = [(j * P) % nblk for j in range(nblk)]
        assert sorted(order) == list(range(nblk))
    nblk = 34539  # C2
    P = int(L.cmx_debug_block_perm(nblk))
    assert abs(P / nblk - 0.6180339887) < 1e-3
    for m in (32, 113, 401, 1425):  # prefixes = the first slabs
        pts = np.sort((np.arange(m, dtype=np.int64) * P) % nblk)
        gaps = np.diff(np.concatenate([pts, [pts[0] + nblk]]))
        assert gaps.max() <= 3.3 * nblk / m, (m, gaps.max(), nblk / m)  # three-distance theorem: no big holes
        half = (pts < nblk // 2).sum()
        assert abs(half - m / 2) <= 2  # both halves of the file (EN rows, ZH rows) equally sampled


def test_no_gpu_fails_loudly():
    """Without a CUDA device the product path must raise, never fall back to the CPU."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import cmx.faiss as faiss
    from cmx import engine

    with pytest.raises(RuntimeError):
        faiss.StandardGpuResources()
    with pytest.raises(RuntimeError):
        engine.Shard(8, 0)
    idx = faiss.IndexFlatIP(8)
    idx.add(np.zeros((3, 8), np.float32))
    with pytest.raises(RuntimeError):
        idx.search(np.zeros((1, 8), np.float32), 2)
    with pytest.raises(RuntimeError):
        engine.mix_normalize(np.zeros((2, 8), np.float32), np.zeros((2, 8), np.float32), [0.5])


def test_no_undefined_globals_in_host_modules():
    """Every global name a function of the host package (and of bench.py / __graft_entry__.py) loads is
    defined at module level or is a builtin -- GPU-only code paths cannot be executed here, but a
    misspelt module alias in one of them must not wait for a GPU box to be found."""
    import builtins
    import dis
    import types

    files = sorted((ROOT / "codemix-dense-retrieval_b200" / "cmx").glob("*.py")) + [ROOT / "bench.py", ROOT / "__graft_entry__.py"]
    assert len(files) >= 8
    problems = []
    for f in files:
        code = compile(f.read_text(), str(f), "exec")
        defined = set(dir(builtins)) | {"__file__", "__name__", "__doc__", "__package__", "__spec__", "__builtins__", "__annotations__"}
        loads = []

        def walk(co):
            for ins in dis.get_instructions(co):
                if ins.opname in ("STORE_NAME", "STORE_GLOBAL") and co is code:
                    defined.add(ins.argval)
                if ins.opname in ("STORE_GLOBAL",):
                    defined.add(ins.argval)
                if ins.opname == "LOAD_GLOBAL" or (ins.opname == "LOAD_NAME" and co is code):
                    loads.append((ins.argval, co.co_name))
            for c in co.co_consts:
                if isinstance(c, types.CodeType):
                    walk(c)

        walk(code)
        # names bound at module level by `import a.b as c`, `from x import y`, def and class are STORE_NAMEs
        for name, where in loads:
            if name not in defined:
                problems.append(f"{f.name}: {name} (in {where})")
    assert not problems, problems
