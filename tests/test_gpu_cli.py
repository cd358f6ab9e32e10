"""End-to-end CLI runs on a synthetic cached index + query cache (needs a GPU): the files the
run scripts write equal what the reference loop would write for the same (D, I), and (D, I)
matches the oracle."""
import json

import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu


def _unit(rng, n, d):
    x = rng.standard_normal((n, d)).astype(np.float32)
    return x / np.linalg.norm(x, axis=1, keepdims=True)


def _make_lang_index(root, lang, X, base_ids):
    import cmx.faiss as faiss
    from cmx import io as cio

    d = root / lang
    d.mkdir(parents=True)
    idx = faiss.IndexIDMap(faiss.IndexFlatIP(X.shape[1]))
    idx.add_with_ids(X, np.arange(X.shape[0], dtype=np.int64))
    faiss.write_index(idx, str(d / "index.faiss"))
    cio.write_docid_map(d / "docid_map.tsv", range(X.shape[0]), [f"{b}#{lang}" for b in base_ids], base_ids, lang)
    (d / "docids.txt").write_text("\n".join(base_ids))


def _make_queries(tmp, rng, nq, d):
    from cmx import io as cio

    qids = [str(2000 + 7 * i) for i in range(nq)]
    P, S = _unit(rng, nq, d), _unit(rng, nq, d)
    (tmp / "queries.en.tsv").write_text("id\ttext\n" + "".join(f"{q}\tquery {q}\n" for q in qids))
    # the zh file has one extra qid and a different order: only the intersection in EN order is used
    zh = qids[::-1] + ["999999"]
    (tmp / "queries.zh.tsv").write_text("".join(f"{q}\t查询 {q}\n" for q in zh))
    cio.save_query_cache(tmp / "qcache", "en", qids, P)
    cio.save_query_cache(tmp / "qcache", "zh", qids, S)
    return qids, P, S


def test_mono_cli(tmp_path):
    from cmx import cli
    from cmx.engine import mix_normalize

    rng = np.random.default_rng(31)
    d, N, nq = 64, 4000, 37
    X = _unit(rng, N, d)
    base = [str(100000 + 3 * i) for i in range(N)]
    _make_lang_index(tmp_path / "idx", "english", X, base)
    qids, P, S = _make_queries(tmp_path, rng, nq, d)
    rc = cli.main(["mono", "--config", "collection-english", "--index_root", str(tmp_path / "idx"),
                   "--query_tsv", f"en={tmp_path / 'queries.en.tsv'}", "--query_tsv", f"zh={tmp_path / 'queries.zh.tsv'}",
                   "--cm_alphas", "0,0.3,1", "--cache_queries", "--query_cache_dir", str(tmp_path / "qcache"),
                   "--run_out", str(tmp_path / "runs"), "--docids_out", str(tmp_path / "docids.txt"),
                   "--gpu_faiss", "--qblock", "1024", "--encoder", "BAAI/bge-m3", "--batch", "1024", "--dtype", "fp16"])
    assert rc == 0
    assert (tmp_path / "docids.txt").read_text() == "\n".join(sorted(set(base)))
    Qg = mix_normalize(P, S, [0.0, 0.3, 1.0])
    lookup = {i: b for i, b in enumerate(base)}
    for ai, label in enumerate(["0", "0.3", "1"]):
        Dr, Ir = oracle.flat_ip_search(X, Qg[ai], 100)
        text = (tmp_path / "runs" / f"cm-alpha-{label}.trec").read_text()
        lines = text.split("\n")
        assert len(lines) == nq * 100
        # parse back and compare with the oracle under the tie-aware rule
        inv = {b: i for i, b in lookup.items()}
        I = np.array([inv[l.split("\t")[2]] for l in lines]).reshape(nq, 100)
        D = np.array([float(l.split("\t")[4]) for l in lines], dtype=np.float32).reshape(nq, 100)
        rep = oracle.compare_topk(D, I, Dr, Ir, rtol=1e-5, atol=6e-5)  # scores carry 4 decimals in the file
        assert rep["ok"], rep
        assert lines[0].split("\t")[:2] == [qids[0], "Q0"] and lines[0].endswith("\tonepass-cm")


def test_bilingual_cli(tmp_path):
    from cmx import cli
    from cmx.engine import mix_normalize
    import cmx.faiss as faiss

    rng = np.random.default_rng(32)
    d, N, nq = 64, 3000, 21
    Xen, Xzh = _unit(rng, N, d), _unit(rng, N, d)
    base = [str(500 + i) for i in range(N)]
    _make_lang_index(tmp_path / "idx", "english", Xen, base)
    _make_lang_index(tmp_path / "idx", "chinese", Xzh, base)
    qids, P, S = _make_queries(tmp_path, rng, nq, d)
    rc = cli.main(["bilingual", "--langs", "english,chinese", "--index_root", str(tmp_path / "idx"),
                   "--query_tsv", f"en={tmp_path / 'queries.en.tsv'}", "--query_tsv", f"zh={tmp_path / 'queries.zh.tsv'}",
                   "--cm_alphas", "0.5", "--cache_queries", "--query_cache_dir", str(tmp_path / "qcache"),
                   "--outdir", str(tmp_path / "out"), "--docids_out", str(tmp_path / "docids.txt"), "--topk", "200",
                   "--gpu_faiss", "--repo", "unicamp-dl/mmarco", "--encoder", "BAAI/bge-m3"])
    assert rc == 0
    out = tmp_path / "out"
    map_lines = (out / "docid_map.tsv").read_text().splitlines()
    assert map_lines[0] == "derived_id\tbase_id\tlang" and len(map_lines) == 2 * N + 1
    assert map_lines[1] == f"{base[0]}#english\t{base[0]}\tenglish" and map_lines[N + 1] == f"{base[0]}#chinese\t{base[0]}\tchinese"
    id2doc = [l.split("\t")[0] for l in map_lines[1:]]
    X = np.concatenate([Xen, Xzh])
    Qg = mix_normalize(P, S, [0.5])[0]
    gpu = faiss.GpuIndexFlatIP(d)
    gpu.add(X)
    D, I = gpu.search(Qg, 200)
    raw = oracle.bilingual_raw_lines(qids, D, I, id2doc, "bilingual-mix-en-zh")
    assert (out / "cm-alpha-0.5_raw.trec").read_text() == "".join(raw)
    assert (out / "cm-alpha-0.5.trec").read_text() == oracle.collapse_run_max_text(raw)
    Dr, Ir = oracle.flat_ip_search(X, Qg, 200)
    assert oracle.compare_topk(D, I, Dr, Ir, rtol=1e-5, atol=1e-6)["ok"]
    meta = json.loads((out / "cm-alpha-0.5_meta.json").read_text())
    assert meta["index"]["size"] == 2 * N and meta["topk"] == 200
    from cmx import runloop

    assert runloop.LAST_SWEEP.get("device_collapse") == 1, runloop.LAST_SWEEP  # the collapsed run came from cmx_collapse_max


def test_bilingual_cli_sharded_and_unordered_map(tmp_path, monkeypatch):
    """The combined index on three shards (one GPU here) gives byte-identical run files; a docid_map.tsv that
    lists the rows in another order than they are stored takes the device-side gather and reproduces the
    reference's row order (map lines in batches, each batch sorted by int id)."""
    from cmx import cli
    from cmx import io as cio

    rng = np.random.default_rng(33)
    d, N, nq = 64, 2500, 17
    Xen, Xzh = _unit(rng, N, d), _unit(rng, N, d)
    base = [str(900 + i) for i in range(N)]
    _make_lang_index(tmp_path / "idx", "english", Xen, base)
    _make_lang_index(tmp_path / "idx", "chinese", Xzh, base)
    qids, P, S = _make_queries(tmp_path, rng, nq, d)

    def run(outname, index_root):
        rc = cli.main(["bilingual", "--langs", "english,chinese", "--index_root", str(index_root),
                       "--query_tsv", f"en={tmp_path / 'queries.en.tsv'}", "--query_tsv", f"zh={tmp_path / 'queries.zh.tsv'}",
                       "--cm_alphas", "0.25,1", "--query_cache_dir", str(tmp_path / "qcache"), "--outdir", str(tmp_path / outname),
                       "--docids_out", str(tmp_path / f"{outname}.docids"), "--topk", "100", "--gpu_faiss"])
        assert rc == 0
        return {p.name: p.read_bytes() for p in sorted((tmp_path / outname).iterdir()) if p.suffix in (".trec", ".tsv")}

    from cmx import runloop

    one = run("one", tmp_path / "idx")
    assert runloop.LAST_SWEEP.get("device_collapse") == 2, runloop.LAST_SWEEP
    monkeypatch.setenv("CMX_DEVICES", "0,0,0")
    three = run("three", tmp_path / "idx")
    monkeypatch.delenv("CMX_DEVICES")
    assert runloop.LAST_SWEEP.get("device_collapse") == 2, runloop.LAST_SWEEP  # pinned host results of the shards, read in place
    assert one.keys() == three.keys() and all(one[k] == three[k] for k in one)
    # a map whose lines come in reversed order: the reference sorts each batch of 20 000 lines by int id, so with
    # fewer lines than one batch the combined row order is the storage order again -- same files
    idx2 = tmp_path / "idx2"
    for lang, X in (("english", Xen), ("chinese", Xzh)):
        _make_lang_index(idx2, lang, X, base)
        lines = (idx2 / lang / "docid_map.tsv").read_text().splitlines()
        (idx2 / lang / "docid_map.tsv").write_text("\n".join([lines[0]] + lines[:0:-1]) + "\n")
    rev = run("rev", idx2)
    assert all(rev[k] == one[k] for k in one)
    # a map that keeps only every second row: the device-side gather builds the smaller combined index
    idx3 = tmp_path / "idx3"
    for lang, X in (("english", Xen), ("chinese", Xzh)):
        _make_lang_index(idx3, lang, X, base)
        lines = (idx3 / lang / "docid_map.tsv").read_text().splitlines()
        (idx3 / lang / "docid_map.tsv").write_text("\n".join([lines[0]] + lines[1::2]) + "\n")
    half = run("half", idx3)
    ml = half["docid_map.tsv"].decode().splitlines()
    assert len(ml) == 1 + 2 * (N // 2) and ml[1] == f"{base[0]}#english\t{base[0]}\tenglish" and ml[2].startswith(f"{base[2]}#english")
    raw = half["cm-alpha-1_raw.trec"].decode().splitlines()
    assert len(raw) == nq * 100 and all(int(l.split()[2].split("#")[0]) % 2 == 0 for l in raw[:200])  # bases 900, 902, ...


def test_missing_caches_stop_with_message(tmp_path):
    from cmx import cli

    with pytest.raises(SystemExit) as e:
        cli.main(["mono", "--config", "collection-english", "--index_root", str(tmp_path / "nope"),
                  "--q_en", "a", "--q_zh", "b", "--run_out", str(tmp_path / "r"), "--docids_out", str(tmp_path / "d.txt")])
    assert "out of scope" in str(e.value)
