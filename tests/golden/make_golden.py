#!/usr/bin/env python
"""Generate golden vectors by EXECUTING the reference's own code for the hot path.

Run in the build container only (needs /root/reference, read-only):

    python tests/golden/make_golden.py

The reference scripts cannot be imported as modules (they import faiss and
sentence_transformers at module top, neither is installed), so the function
definitions / f-strings on the hot path are pulled out of the reference files
with ``ast`` at generation time and executed here.  No reference source is
written into this repository -- only the inputs and the outputs it produced.

``normalize_embeddings`` (sentence-transformers 5.0.0 ``util.normalize_embeddings``)
is the one external symbol ``safe_mix`` needs; its published definition is
``torch.nn.functional.normalize(embeddings, p=2, dim=1)`` and is bound as such.

Outputs (committed): tests/golden/mix_golden.npz, tests/golden/text_golden.json and -- ``--cli`` --
tests/golden/cli_golden.json (the user-visible helpers of the run scripts: cache directory names, the
--query_tsv parser and its messages).
"""

from __future__ import annotations

import ast
import json
import logging
import pathlib
import tempfile

import numpy as np
import torch

REF = pathlib.Path("/root/reference")
MONO = REF / "onepass_dense_mix_run_custom_lang.py"
BILI = REF / "onepass_bilingual_mix_hub_custom_lang.py"
OUT = pathlib.Path(__file__).resolve().parent


def _load_functions(path: pathlib.Path, names):
    tree = ast.parse(path.read_text(encoding="utf-8"))
    ns = {
        "np": np,
        "torch": torch,
        "logging": logging,
        "normalize_embeddings": lambda t: torch.nn.functional.normalize(t, p=2, dim=1),
        "Path": pathlib.Path,
        "pathlib": pathlib,
    }
    exec("from typing import *", ns)
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            mod = ast.Module(body=[node], type_ignores=[])
            exec(compile(mod, str(path), "exec"), ns)
    missing = [n for n in names if n not in ns]
    assert not missing, missing
    return ns


def _find_fstring(path: pathlib.Path, must_contain):
    tree = ast.parse(path.read_text(encoding="utf-8"))
    hits = []
    for node in ast.walk(tree):
        if isinstance(node, ast.JoinedStr):
            text = ast.unparse(node)
            if all(m in text for m in must_contain):
                hits.append(node)
    assert len(hits) >= 1, (path, must_contain)
    return compile(ast.Expression(body=hits[0]), str(path), "eval")


def make_inputs():
    rng = np.random.default_rng(20260118)
    nq, d = 24, 64
    P = rng.standard_normal((nq, d)).astype(np.float32)
    P /= np.linalg.norm(P, axis=1, keepdims=True)
    G = rng.standard_normal((nq, d)).astype(np.float32)
    G /= np.linalg.norm(G, axis=1, keepdims=True)
    S = (0.8 * P + 0.6 * G).astype(np.float32)
    S /= np.linalg.norm(S, axis=1, keepdims=True)
    # edge rows
    S[3] = -P[3]  # alpha=0.5 -> exact zero vector -> 0/eps = 0 (finite, no fallback)
    P[5, 7] = np.nan  # non-finite -> fallback
    S[6, 0] = np.inf  # non-finite -> fallback
    P[8] = 0.0
    S[8] = 0.0  # zero vectors
    P[9] *= 1e-30
    S[9] *= 1e-30  # tiny norms: clamp at eps=1e-12
    P[10] *= 1e20
    S[10] *= 1e20  # sum of squares overflows fp32 -> inf norm -> zeros
    alphas = [0.0, 1e-9, 0.1, 0.25, 0.3, 0.5, 0.5000001, 0.7, 0.75, 0.9, 1.0 - 1e-9, 1.0, 1.5, -0.25]
    return P, S, alphas


def main():
    mono = _load_functions(MONO, ["safe_mix", "format_alpha", "parse_alpha_list"])
    bili = _load_functions(BILI, ["safe_mix", "collapse_run_max", "format_alpha"])
    logging.disable(logging.WARNING)

    # ---- safe_mix ---------------------------------------------------------
    P, S, alphas = make_inputs()
    nq, d = P.shape
    Q = np.empty((len(alphas), nq, d), dtype=np.float32)
    Qb = np.empty_like(Q)
    for ai, a in enumerate(alphas):
        for qi in range(nq):
            Q[ai, qi] = mono["safe_mix"](P[qi], S[qi], a, str(qi), "cpu", ("en", "zh"))
            Qb[ai, qi] = bili["safe_mix"](P[qi], S[qi], a, str(qi), "cpu", ("en", "zh"))
    assert np.array_equal(Q, Qb, equal_nan=True), "mono and bilingual safe_mix disagree"
    np.savez_compressed(OUT / "mix_golden.npz", P=P, S=S, alphas=np.array(alphas, dtype=np.float64), Q=Q)

    # ---- labels / parsing -------------------------------------------------
    label_in = [0.0, 1.0, 0.1, 0.25, 0.3, 0.5, 0.7, 0.75, 0.9, 0.12345, 0.99999, 1e-9, 1 - 1e-9,
                0.00004, 0.00005, 0.00006, 2.0, -1.0, -0.5, 0.3333333, 1.25, 10.0]
    labels = [mono["format_alpha"](a) for a in label_in]
    assert labels == [bili["format_alpha"](a) for a in label_in]
    parse_in = ["0,0.1,0.3,0.5,0.7,0.9,1", "0.0,0.25,0.5,0.75,1.0", " 0.5 , ,1 ", "", ",", "0.5,abc", "1e-1,.5"]
    parse_out = []
    for s in parse_in:
        try:
            parse_out.append(mono["parse_alpha_list"](s))
        except SystemExit as exc:
            parse_out.append({"SystemExit": str(exc)})

    # ---- TREC f-strings ---------------------------------------------------
    mono_expr = _find_fstring(MONO, ["onepass-cm", "Q0"])
    raw_expr = _find_fstring(BILI, ["Q0", "{tag}", "{sc:.6f}"])
    rng = np.random.default_rng(7)
    scores = np.concatenate([
        rng.uniform(-1, 1, 40).astype(np.float32),
        np.array([0.0, -0.0, 0.99995, 0.00005, 0.12345, 0.123449, 0.5, 1.0, -1.0, 0.7071068,
                  1e-7, 0.99999994, 3.4028235e38, -3.4028235e38, 0.00015, 0.00025, 0.00035], dtype=np.float32),
    ])
    mono_lines, raw_lines = [], []
    for i, sc in enumerate(scores):
        qid, doc, rank = str(1000 + i), str(7000000 + 13 * i), i + 1
        mono_lines.append(eval(mono_expr, {}, {"qid": qid, "doc": doc, "rank": rank, "score": sc}))
        raw_lines.append(eval(raw_expr, {}, {"qid": qid, "did": f"{doc}#en", "rank": rank,
                                             "sc": np.float32(sc).tolist(), "tag": "bilingual-mix-en-zh"}))

    # ---- collapse_run_max -------------------------------------------------
    rng = np.random.default_rng(11)
    raw = []
    for q in range(6):
        base_ids = rng.integers(0, 40, size=30)
        sc = np.sort(rng.uniform(0.2, 0.9, 30).astype(np.float32))[::-1]
        sc[5] = sc[4]  # exact tie
        sc[11] = np.float32(sc[10] - 2e-7)  # ties after 6-decimal rounding
        for rank, (b, s) in enumerate(zip(base_ids, sc), 1):
            lang = "en" if rng.random() < 0.5 else "zh"
            if rank in (7, 19):
                continue  # skipped slots keep their rank numbers in the raw file
            raw.append(f"q{q} Q0 {b}#{lang} {rank} {float(s):.6f} bilingual-mix-en-zh\n")
    with tempfile.TemporaryDirectory() as td:
        pin = pathlib.Path(td) / "in_raw.trec"
        pout = pathlib.Path(td) / "out.trec"
        pin.write_text("".join(raw), encoding="utf-8")
        bili["collapse_run_max"](pin, pout)
        collapsed = pout.read_text(encoding="utf-8")

    (OUT / "text_golden.json").write_text(json.dumps({
        "format_alpha": {"in": label_in, "out": labels},
        "parse_alpha_list": {"in": parse_in, "out": parse_out},
        "trec_scores_f32_hex": [np.float32(s).tobytes().hex() for s in scores],
        "mono_lines": mono_lines,
        "raw_lines": raw_lines,
        "collapse": {"raw": raw, "out": collapsed},
    }, indent=1), encoding="utf-8")
    print("wrote", OUT / "mix_golden.npz", OUT / "text_golden.json")


def main_cli():
    import os
    import re
    import sys

    ns = _load_functions(MONO, ["sanitize_tag", "default_query_cache_root", "parse_query_specs"])
    ns.update({"re": re, "os": os, "__file__": "/nonexistent/script.py"})
    tags = ["unicamp-dl/mmarco", "BAAI/bge-m3", "/a b/c//", "---", "", "Qwen/Qwen3-Embedding-8B", "é–x y.z_", "//", "a/b/"]
    out = {"sanitize_tag": {"in": tags, "out": [ns["sanitize_tag"](t) for t in tags]}}
    roots = []
    for env in ({}, {"QUERY_CACHE_ROOT": "/data/qc"}, {"QUERY_CACHE_ROOT_BASE": "/mnt/base"}, {"QUERY_CACHE_ROOT_BASE": ""}):
        for k in ("QUERY_CACHE_ROOT", "QUERY_CACHE_ROOT_BASE"):
            os.environ.pop(k, None)
        os.environ.update(env)
        for repo, enc in (("unicamp-dl/mmarco", "BAAI/bge-m3"), ("x/y z", "intfloat/multilingual-e5-large")):
            if not env:  # the reference's default base is its own script directory: only the leaf is comparable
                roots.append({"env": env, "repo": repo, "encoder": enc, "leaf": ns["default_query_cache_root"](repo, enc).name})
            else:
                roots.append({"env": env, "repo": repo, "encoder": enc, "path": str(ns["default_query_cache_root"](repo, enc))})
    for k in ("QUERY_CACHE_ROOT", "QUERY_CACHE_ROOT_BASE"):
        os.environ.pop(k, None)
    out["default_query_cache_root"] = roots
    specs_in = [(["en=a.tsv", "zh=b.tsv"], None, None), ([" hi = /x/y.tsv ", "en=/z"], None, None), (None, "q.en", "q.zh"),
                (None, "q.en", None), (None, None, None), (["en"], None, None), (["=x", "zh=y"], None, None), (["en=", "zh=y"], None, None),
                (["en=a"], None, None), (["en=a", "zh=b", "de=c"], None, None), (["en=a", "en=b"], None, None), ([], "a", "b"),
                (["a=b=c", "d=e"], None, None)]
    res = []
    for args in specs_in:
        try:
            res.append([[lang, str(path)] for lang, path in ns["parse_query_specs"](*args)])
        except SystemExit as exc:
            res.append({"SystemExit": str(exc)})
    out["parse_query_specs"] = {"in": [list(a) for a in specs_in], "out": res}
    (OUT / "cli_golden.json").write_text(json.dumps(out, indent=1), encoding="utf-8")
    print("wrote", OUT / "cli_golden.json")


if __name__ == "__main__":
    import sys

    if "--cli" in sys.argv:
        main_cli()
    else:
        main()
