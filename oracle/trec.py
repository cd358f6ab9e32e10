"""Oracle for alpha labels and TREC run text (test infrastructure only).

Pure-Python restatements, row at a time, of:
  * ``format_alpha``       onepass_dense_mix_run_custom_lang.py:304-308
  * ``parse_alpha_list``   onepass_dense_mix_run_custom_lang.py:287-301
  * mono TREC lines        onepass_dense_mix_run_custom_lang.py:879-888
  * bilingual raw lines    onepass_bilingual_mix_hub_custom_lang.py:950-958
  * ``collapse_run_max``   onepass_bilingual_mix_hub_custom_lang.py:165-181
"""

from __future__ import annotations

from typing import Dict, Iterable, List, Sequence


def format_alpha(alpha: float) -> str:
    # integer-valued -> "0"/"1"; else 4 decimals with trailing zeros stripped
    if abs(alpha - round(alpha)) < 1e-8:
        return str(int(round(alpha)))
    text = f"{alpha:.4f}".rstrip("0").rstrip(".")
    return text if text else "0"


def parse_alpha_list(alpha_str: str) -> List[float]:
    if not alpha_str:
        raise SystemExit("--cm_alphas must contain at least one value.")
    out: List[float] = []
    for tok in alpha_str.split(","):
        tok = tok.strip()
        if not tok:
            continue
        try:
            out.append(float(tok))
        except ValueError as exc:
            raise SystemExit(f"[ERROR] Could not parse alpha '{tok}': {exc}") from exc
    if not out:
        raise SystemExit("No valid alpha values parsed from --cm_alphas.")
    return out


def mono_trec_lines(qids: Sequence[str], D, I, id_lookup: Dict[int, str]) -> List[str]:
    """Lines of one mono run; the file is ``"\\n".join(lines)`` (no trailing newline)."""
    lines: List[str] = []
    for row, qid in enumerate(qids):
        doc_ids = [id_lookup.get(int(doc), str(doc)) for doc in I[row]]
        for rank, (doc, score) in enumerate(zip(doc_ids, D[row]), 1):
            lines.append(f"{qid}\tQ0\t{doc}\t{rank}\t{float(score):.4f}\tonepass-cm")
    return lines


def bilingual_raw_lines(qids: Sequence[str], D, I, id2doc: Sequence[str], tag: str) -> List[str]:
    """Raw bilingual lines (each ends with a newline); out-of-range ids are skipped
    and the rank of the skipped slot is NOT reused."""
    lines: List[str] = []
    for r, qid in enumerate(qids):
        for rank, (sc, ix) in enumerate(zip(list(map(float, D[r])), list(map(int, I[r]))), 1):
            if ix < 0 or ix >= len(id2doc):
                continue
            lines.append(f"{qid} Q0 {id2doc[ix]} {rank} {sc:.6f} {tag}\n")
    return lines


def collapse_run_max_text(raw_lines: Iterable[str]) -> str:
    """Collapse ``base#lang`` ids to ``base`` keeping the max (6-decimal) score."""
    by_q: Dict[str, Dict[str, List[float]]] = {}
    for line in raw_lines:
        line = line.strip()
        if not line:
            continue
        qid, _, did, _rk, sc, _tag = line.split()
        base = did.split("#", 1)[0]
        by_q.setdefault(qid, {}).setdefault(base, []).append(float(sc))
    out: List[str] = []
    for qid, groups in by_q.items():
        items = [(b, max(scores)) for b, scores in groups.items()]
        items.sort(key=lambda x: x[1], reverse=True)
        for rank, (base, val) in enumerate(items, 1):
            out.append(f"{qid} Q0 {base} {rank} {val:.6f} bilingual-mix\n")
    return "".join(out)


def collapse_run_max(in_run, out_run) -> None:
    with open(in_run, "r", encoding="utf-8") as f:
        text = collapse_run_max_text(f)
    with open(out_run, "w", encoding="utf-8") as out:
        out.write(text)
