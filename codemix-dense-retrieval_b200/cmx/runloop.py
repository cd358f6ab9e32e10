"""The per-alpha run loop of the onepass vector-mix scripts, array-at-a-time.

Mirrors (same names, argument meaning, file names and byte-for-byte file content):
  * mono:      onepass_dense_mix_run_custom_lang.py:844-890  -> ``run_alpha_sweep``
  * bilingual: onepass_bilingual_mix_hub_custom_lang.py:901-962 -> ``run_alpha_sweep_bilingual``
  * ``format_alpha`` :304-308, ``parse_alpha_list`` :287-301,
    ``collapse_run_max`` onepass_bilingual_mix_hub_custom_lang.py:165-181

What changed relative to the reference loop: the ~7000 per-query ``safe_mix`` GPU
round-trips per alpha become one fused kernel (``search_mixed``), and the ~700k
f-strings per alpha become vectorised numpy string ops.  Output files are
identical to what the reference loop would write for the same (D, I).
"""
from __future__ import annotations

import json
import os
import pathlib
import time
from typing import Dict, List, Optional, Sequence

import numpy as np

_SDT = np.dtypes.StringDType()


# ---- alpha labels -------------------------------------------------------------
def format_alpha(alpha: float) -> str:
    if abs(alpha - round(alpha)) < 1e-8:
        return str(int(round(alpha)))
    text = f"{alpha:.4f}".rstrip("0").rstrip(".")
    return text if text else "0"


def parse_alpha_list(alpha_str: str) -> List[float]:
    if not alpha_str:
        raise SystemExit("--cm_alphas must contain at least one value.")
    alphas: List[float] = []
    for tok in alpha_str.split(","):
        tok = tok.strip()
        if not tok:
            continue
        try:
            alphas.append(float(tok))
        except ValueError as exc:
            raise SystemExit(f"[ERROR] Could not parse alpha '{tok}': {exc}") from exc
    if not alphas:
        raise SystemExit("No valid alpha values parsed from --cm_alphas.")
    return alphas


# ---- vectorised fixed-point score text ------------------------------------------
def format_scores(D: np.ndarray, decimals: int) -> np.ndarray:
    """Elementwise equivalent of ``f"{float(x):.{decimals}f}"`` for float32 input.

    A float32 times 10**decimals (decimals <= 6) is exact in float64, so rounding
    half-to-even on it reproduces Python's correctly rounded decimal formatting.
    Values too large for that (|x| >= 1e11), NaN and inf take the slow exact path.
    """
    x = np.asarray(D, dtype=np.float32).astype(np.float64)
    scale = 10 ** decimals
    flat = x.reshape(-1)
    slow = ~np.isfinite(flat) | (np.abs(flat) >= 1e11)
    safe = np.where(slow, 0.0, flat)
    n = np.rint(np.abs(safe) * scale).astype(np.int64)
    ip = (n // scale).astype(_SDT)
    fp = np.strings.zfill((n % scale).astype(_SDT), decimals)
    sign = np.where(np.signbit(safe), "-", "").astype(_SDT)
    txt = np.strings.add(np.strings.add(np.strings.add(sign, ip), "."), fp)
    if slow.any():
        idx = np.nonzero(slow)[0]
        fmt = f"{{:.{decimals}f}}"
        for i in idx:
            txt[i] = fmt.format(float(flat[i]))
    return txt.reshape(x.shape)


class DocTable:
    """int id -> document-id string, vectorised; ids missing from the table print as
    ``str(id)`` (the reference's ``id_lookup.get(int(doc), str(doc))``)."""

    def __init__(self, id_lookup):
        if isinstance(id_lookup, DocTable):
            self.keys, self.names = id_lookup.keys, id_lookup.names
            return
        if isinstance(id_lookup, dict):
            keys = np.fromiter(id_lookup.keys(), dtype=np.int64, count=len(id_lookup))
            names = np.array(list(id_lookup.values()), dtype=_SDT)
        else:  # sequence: position = id
            names = np.array(list(id_lookup), dtype=_SDT)
            keys = np.arange(len(names), dtype=np.int64)
        order = np.argsort(keys, kind="stable")
        self.keys, self.names = keys[order], names[order]

    def lookup(self, ids: np.ndarray) -> np.ndarray:
        ids = np.asarray(ids, dtype=np.int64)
        flat = ids.reshape(-1)
        out = flat.astype(_SDT)
        if self.keys.size:
            pos = np.clip(np.searchsorted(self.keys, flat), 0, self.keys.size - 1)
            hit = self.keys[pos] == flat
            out = np.where(hit, self.names[pos], out)
        return out.reshape(ids.shape)


def mono_trec_text(qids: Sequence[str], D: np.ndarray, I: np.ndarray, docs: DocTable, tag: str = "onepass-cm") -> str:
    """Text of one mono run file: lines ``qid\\tQ0\\tdoc\\trank\\tscore(.4f)\\ttag`` joined
    with newlines, no trailing newline (onepass_dense_mix_run_custom_lang.py:879-888)."""
    D = np.asarray(D)
    I = np.asarray(I)
    nq, k = D.shape
    if nq == 0:
        return ""
    q = np.repeat(np.array(list(qids), dtype=_SDT), k).reshape(nq, k)
    rank = np.tile(np.arange(1, k + 1, dtype=np.int64).astype(_SDT), (nq, 1))
    line = np.strings.add(q, "\tQ0\t")
    line = np.strings.add(line, docs.lookup(I))
    line = np.strings.add(np.strings.add(line, "\t"), rank)
    line = np.strings.add(np.strings.add(line, "\t"), format_scores(D, 4))
    line = np.strings.add(line, "\t" + tag)
    return "\n".join(line.reshape(-1).tolist())


def bilingual_raw_text(qids: Sequence[str], D: np.ndarray, I: np.ndarray, id2doc: Sequence[str], tag: str) -> str:
    """Raw bilingual run: ``qid Q0 did rank score(.6f) tag\\n``; rows with an id outside
    [0, len(id2doc)) are skipped and keep their rank number
    (onepass_bilingual_mix_hub_custom_lang.py:950-958)."""
    D = np.asarray(D)
    I = np.asarray(I)
    nq, k = D.shape
    if nq == 0:
        return ""
    names = id2doc if isinstance(id2doc, np.ndarray) and id2doc.dtype == _SDT else np.array(list(id2doc), dtype=_SDT)
    valid = (I >= 0) & (I < len(names))
    did = names[np.clip(I, 0, max(len(names) - 1, 0))] if len(names) else np.full(I.shape, "", dtype=_SDT)
    q = np.repeat(np.array(list(qids), dtype=_SDT), k).reshape(nq, k)
    rank = np.tile(np.arange(1, k + 1, dtype=np.int64).astype(_SDT), (nq, 1))
    line = np.strings.add(np.strings.add(q, " Q0 "), did)
    line = np.strings.add(np.strings.add(line, " "), rank)
    line = np.strings.add(np.strings.add(line, " "), format_scores(D, 6))
    line = np.strings.add(line, " " + tag + "\n")
    return "".join(line[valid].tolist())


def collapse_by_base(qids: Sequence[str], D: np.ndarray, I: np.ndarray, id2doc: Sequence[str]) -> str:
    """``collapse_run_max`` computed from (D, I) instead of re-parsing the raw text:
    group hits of a query by ``base = did.split('#',1)[0]``, score = max of the
    6-decimal-ROUNDED scores, stable descending sort (ties keep first-seen order),
    re-rank, tag ``bilingual-mix`` (onepass_bilingual_mix_hub_custom_lang.py:165-181)."""
    D = np.asarray(D)
    I = np.asarray(I)
    nq, k = D.shape
    names = list(id2doc)
    bases_all = np.array([n.split("#", 1)[0] for n in names], dtype=_SDT) if names else np.empty((0,), dtype=_SDT)
    # integer code per base string (first-seen order is restored per query below)
    uniq, base_code = (np.unique(bases_all, return_inverse=True) if len(names) else (bases_all, np.empty((0,), np.int64)))
    valid = (I >= 0) & (I < len(names))
    # scores as the raw file would carry them: rounded to 6 decimals, then float()
    x = np.asarray(D, dtype=np.float32).astype(np.float64)
    rounded = np.copysign(np.rint(np.abs(x) * 1e6), x) / 1e6
    out: List[str] = []
    # a qid occurring in several rows is one group in the reference (dict keyed by qid)
    first_row: Dict[str, int] = {}
    rows_of: Dict[str, List[int]] = {}
    for r, qid in enumerate(qids):
        rows_of.setdefault(qid, []).append(r)
        first_row.setdefault(qid, r)
    for qid, rows in rows_of.items():
        sel_codes = np.concatenate([base_code[I[r][valid[r]]] for r in rows]) if rows else np.empty((0,), np.int64)
        sel_scores = np.concatenate([rounded[r][valid[r]] for r in rows])
        if sel_codes.size == 0:
            continue
        # first-seen order of bases + max score per base
        u, first_idx, inv = np.unique(sel_codes, return_index=True, return_inverse=True)
        mx = np.full(u.shape, -np.inf)
        np.maximum.at(mx, inv, sel_scores)
        seen_order = np.argsort(first_idx, kind="stable")
        u, mx = u[seen_order], mx[seen_order]
        order = np.argsort(-mx, kind="stable")
        u, mx = u[order], mx[order]
        base_txt = uniq[u]
        score_txt = _format_f64(mx, 6)
        ranks = np.arange(1, u.size + 1, dtype=np.int64).astype(_SDT)
        line = np.strings.add(qid + " Q0 ", base_txt)
        line = np.strings.add(np.strings.add(line, " "), ranks)
        line = np.strings.add(np.strings.add(line, " "), score_txt)
        line = np.strings.add(line, " bilingual-mix\n")
        out.append("".join(line.tolist()))
    return "".join(out)


def _format_f64(x: np.ndarray, decimals: int) -> np.ndarray:
    """f"{v:.6f}" for doubles that are already multiples of 1e-6 (up to rounding)."""
    scale = 10 ** decimals
    n = np.rint(np.abs(x) * scale).astype(np.int64)
    ip = (n // scale).astype(_SDT)
    fp = np.strings.zfill((n % scale).astype(_SDT), decimals)
    sign = np.where(np.signbit(x), "-", "").astype(_SDT)
    return np.strings.add(np.strings.add(np.strings.add(sign, ip), "."), fp)


def collapse_run_max(in_run, out_run) -> None:
    """Text-to-text form with the reference's signature (used when only the raw file exists)."""
    by_q: Dict[str, Dict[str, float]] = {}
    with open(in_run, "r", encoding="utf-8") as f:
        for line in f:
            line = line.strip()
            if not line:
                continue
            qid, _, did, _rk, sc, _tag = line.split()
            base = did.split("#", 1)[0]
            score = float(sc)
            g = by_q.setdefault(qid, {})
            if base not in g or score > g[base]:
                g[base] = score
    with open(out_run, "w", encoding="utf-8") as out:
        for qid, groups in by_q.items():
            items = sorted(groups.items(), key=lambda kv: kv[1], reverse=True)
            for rank, (base, val) in enumerate(items, 1):
                out.write(f"{qid} Q0 {base} {rank} {val:.6f} bilingual-mix\n")


def _atomic_write(path: pathlib.Path, text: str) -> None:
    tmp = path.with_name(path.name + f".tmp{os.getpid()}")
    tmp.write_text(text, encoding="utf-8")
    os.replace(tmp, path)


def _to_host(a):
    return a.detach().cpu().numpy() if hasattr(a, "detach") else np.asarray(a)


# ---- the sweeps -------------------------------------------------------------------
def run_alpha_sweep(index, id_lookup, qids: Sequence[str], P, S, alphas: Sequence[float], outdir,
                    k: int = 100, qblock: int = 256, tag: str = "onepass-cm", alpha_batch: int = 1,
                    log=None) -> List[pathlib.Path]:
    """Mono vector-mix sweep: for every alpha write ``cm-alpha-<label>.trec`` with the top-k
    of normalise((1-alpha) P + alpha S) against ``index``.

    ``index`` is a cmx.faiss index (anything with ``search_mixed``); ``id_lookup`` maps the
    ints returned by the index to document ids (dict, sequence or DocTable).  ``qblock`` is
    accepted for CLI compatibility: searches are independent per query, so the whole alpha
    batch is issued in one call.  ``alpha_batch`` alphas share one fused launch sequence.
    """
    outdir = pathlib.Path(outdir)
    outdir.mkdir(parents=True, exist_ok=True)
    docs = DocTable(id_lookup)
    qids = list(qids)
    written: List[pathlib.Path] = []
    alphas = [float(a) for a in alphas]
    for a0 in range(0, len(alphas), max(1, alpha_batch)):
        group = alphas[a0 : a0 + max(1, alpha_batch)]
        t0 = time.perf_counter()
        D, I = index.search_mixed(P, S, group, k)
        D, I = _to_host(D), _to_host(I)
        t1 = time.perf_counter()
        for gi, alpha in enumerate(group):
            label = format_alpha(alpha)
            run_path = outdir / f"cm-alpha-{label}.trec"
            _atomic_write(run_path, mono_trec_text(qids, D[gi], I[gi], docs, tag))
            written.append(run_path)
            if log:
                log(f"Run saved: {run_path}  ({len(qids)} queries, alpha={label}, search {t1 - t0:.3f}s)")
    return written


def run_alpha_sweep_bilingual(index, id2doc: Sequence[str], qids: Sequence[str], P, S, alphas: Sequence[float],
                              outdir, topk: int = 500, qblock: int = 128, tag: str = "bilingual-mix",
                              meta: Optional[dict] = None, alpha_batch: int = 1, log=None) -> List[pathlib.Path]:
    """Bilingual sweep: per alpha write ``cm-alpha-<label>_raw.trec`` (derived ``base#lang``
    ids), the collapsed ``cm-alpha-<label>.trec`` (max over languages) and
    ``cm-alpha-<label>_meta.json``."""
    outdir = pathlib.Path(outdir)
    outdir.mkdir(parents=True, exist_ok=True)
    qids = list(qids)
    names = np.array(list(id2doc), dtype=_SDT)
    written: List[pathlib.Path] = []
    alphas = [float(a) for a in alphas]
    for a0 in range(0, len(alphas), max(1, alpha_batch)):
        group = alphas[a0 : a0 + max(1, alpha_batch)]
        D, I = index.search_mixed(P, S, group, topk)
        D, I = _to_host(D), _to_host(I)
        for gi, alpha in enumerate(group):
            label = format_alpha(alpha)
            set_name = f"cm-alpha-{label}"
            run_raw = outdir / f"{set_name}_raw.trec"
            run_base = outdir / f"{set_name}.trec"
            _atomic_write(run_raw, bilingual_raw_text(qids, D[gi], I[gi], names, tag))
            _atomic_write(run_base, collapse_by_base(qids, D[gi], I[gi], id2doc))
            m = dict(meta or {})
            m.update({"alpha": label, "runs": {"raw": str(run_raw), "base": str(run_base)},
                      "index": {"type": "IndexIDMap(IndexFlatIP)", "size": int(index.ntotal), "dim": int(index.d)},
                      "topk": int(topk), "qblock": int(qblock)})
            _atomic_write(outdir / f"{set_name}_meta.json", json.dumps(m, indent=2))
            written.append(run_base)
            if log:
                log(f"Completed set '{set_name}' -> {run_raw.name} , {run_base.name}")
    return written
