"""Command-line run scripts for the cached case of the reference's onepass vector-mix jobs.

    python -m cmx.cli mono       <flags of onepass_dense_mix_run_custom_lang.py>
    python -m cmx.cli bilingual  <flags of onepass_bilingual_mix_hub_custom_lang.py>

Same flag names and output files as the reference scripts
(onepass_dense_mix_run_custom_lang.py:388-477, onepass_bilingual_mix_hub_custom_lang.py:429-500) for the
path this repository covers: cached per-language index (``<index_root>/<lang>/{index.faiss,docid_map.tsv,
docids.txt}``) + cached query vectors (``<cache>/<lang>/queries.npz``) -> per-alpha TREC files.  Encoding
corpora or queries (HF encoder + network) is out of scope: when a cache is missing the script stops and says
so instead of loading an encoder.  Encoder-side flags are accepted and ignored.
"""
from __future__ import annotations

import argparse
import logging
import os
import pathlib
import re
import sys
import time
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import faiss
from . import io as cio
from . import runloop

RECON_BATCH = 20000  # the reference adds cached rows in batches of 20 000 map lines, each sorted by id


def sanitize_tag(text: str) -> str:
    clean = re.sub(r"[^A-Za-z0-9_.-]+", "-", text.strip("/"))
    return clean.strip("-") or "run"


def default_query_cache_root(repo: str, encoder: str) -> pathlib.Path:
    env_root = os.environ.get("QUERY_CACHE_ROOT")
    if env_root:
        return pathlib.Path(env_root)
    base = os.environ.get("QUERY_CACHE_ROOT_BASE", str(pathlib.Path.cwd() / "data"))
    return pathlib.Path(base) / f"enc-query-{sanitize_tag(repo.split('/')[-1])}-{sanitize_tag(encoder.split('/')[-1])}"


def parse_query_specs(query_tsv_args, q_primary, q_secondary) -> List[Tuple[str, pathlib.Path]]:
    specs: List[Tuple[str, pathlib.Path]] = []
    if query_tsv_args:
        for entry in query_tsv_args:
            if "=" not in entry:
                raise SystemExit(f"--query_tsv expects LANG=PATH, got '{entry}'.")
            lang, path = (v.strip() for v in entry.split("=", 1))
            if not lang or not path:
                raise SystemExit(f"[ERROR] Bad --query_tsv entry '{entry}'.")
            specs.append((lang, pathlib.Path(path)))
    else:
        if not q_primary or not q_secondary:
            raise SystemExit("Provide either --query_tsv twice or both --q_en and --q_zh.")
        specs = [("en", pathlib.Path(q_primary)), ("zh", pathlib.Path(q_secondary))]
    if len(specs) != 2:
        raise SystemExit(f"Exactly two query TSV specs required, got {len(specs)}.")
    if specs[0][0] == specs[1][0]:
        raise SystemExit(f"Duplicate language '{specs[0][0]}' in query specs.")
    return specs


def load_cached_index(index_root: pathlib.Path, lang: str, expected_dim: Optional[int] = None):
    lang_dir = index_root / lang
    paths = [lang_dir / "index.faiss", lang_dir / "docid_map.tsv", lang_dir / "docids.txt"]
    if not all(p.exists() for p in paths):
        return None
    index = faiss.read_index(str(paths[0]))
    if expected_dim and index.d != expected_dim:
        logging.warning("Cached index dim mismatch for %s: expected %s, found %s.", lang, expected_dim, index.d)
        return None
    base = faiss.downcast_index(index.index if hasattr(index, "index") else index)
    base.reconstruct(0, np.empty((base.d,), dtype=np.float32))
    return {"lang": lang, "index": index, "base_index": base, "map_path": paths[1]}


def _common(ap: argparse.ArgumentParser) -> None:
    ap.add_argument("--repo", default="unicamp-dl/mmarco")
    ap.add_argument("--encoder", default="BAAI/bge-m3")
    ap.add_argument("--query_tsv", action="append", metavar="LANG=PATH")
    ap.add_argument("--q_en")
    ap.add_argument("--q_zh")
    ap.add_argument("--qid_field", default="id")
    ap.add_argument("--qtext_field", default="text")
    ap.add_argument("--cm_alphas", default="0.0,0.25,0.5,0.75,1.0")
    ap.add_argument("--gpu_faiss", action="store_true")
    ap.add_argument("--faiss_gpu_id", type=int, default=0)
    ap.add_argument("--gpus", type=int, default=1, help="row-shard the index over this many GPUs (new)")
    ap.add_argument("--max_queries", type=int)
    ap.add_argument("--docids_out", required=True)
    ap.add_argument("--index_root", default=os.environ.get("INDEX_ROOT", "indexes/idx-mmarco-bge-m3"))
    ap.add_argument("--cache_queries", action="store_true")
    ap.add_argument("--query_cache_dir")
    ap.add_argument("--alpha_batch", type=int, default=1, help="alphas fused into one launch sequence (new)")


def _load_queries(args):
    (l1, p1), (l2, p2) = parse_query_specs(args.query_tsv, args.q_en, args.q_zh)
    rows1 = cio.read_queries_tsv(p1, args.qid_field, args.qtext_field)
    rows2 = cio.read_queries_tsv(p2, args.qid_field, args.qtext_field)
    second = {q for q, _ in rows2}
    common = [q for q, _ in rows1 if q in second]
    if not common:
        raise SystemExit(f"No overlapping qids between query files for {l1} and {l2}.")
    if args.max_queries:
        common = common[: args.max_queries]
    root = pathlib.Path(args.query_cache_dir) if args.query_cache_dir else default_query_cache_root(args.repo, args.encoder)
    P = cio.load_query_cache(root, l1, common)
    S = cio.load_query_cache(root, l2, common)
    for lang, v in ((l1, P), (l2, S)):
        if v is None:
            raise SystemExit(f"No usable query cache for '{lang}' under {root} (qids must match in order). "
                             "Encoding queries needs the HF encoder and is out of scope of this path; "
                             "produce the cache with the reference's cache_queries_for_mix.py.")
    return (l1, l2), common, P, S


def _place(index_cpu, args):
    if args.gpus and args.gpus > 1:
        return faiss.index_cpu_to_gpus_list(index_cpu, gpus=list(range(args.gpus)))
    return faiss.index_cpu_to_gpu(faiss.StandardGpuResources(), args.faiss_gpu_id, index_cpu)


def main_mono(argv: Optional[Sequence[str]] = None) -> int:
    ap = argparse.ArgumentParser(prog="cmx.cli mono", formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    _common(ap)
    ap.add_argument("--config", required=True, help="collection-<lang>")
    ap.add_argument("--qblock", type=int, default=256)
    ap.add_argument("--topk", type=int, default=100, help="the reference hard-codes 100")
    ap.add_argument("--run_out", required=True)
    args, ignored = ap.parse_known_args(argv)
    logging.basicConfig(format="%(asctime)s | %(levelname)s | %(message)s", datefmt="%H:%M:%S", level=logging.INFO)
    if ignored:
        logging.info("Ignoring encoder-side flags: %s", " ".join(ignored))
    doc_lang = args.config.replace("collection-", "")
    cached = load_cached_index(pathlib.Path(args.index_root), doc_lang)
    if cached is None:
        raise SystemExit(f"No cached index for '{doc_lang}' under {args.index_root}; encoding the corpus is out of scope "
                         "(build it with the reference's encode_multilingual_corpus.py).")
    id_lookup, docids_text, n_kept = cio.read_docid_table(cached["map_path"])
    logging.info("Cached index ready for %s: %d vectors", doc_lang, n_kept)
    pathlib.Path(args.docids_out).parent.mkdir(parents=True, exist_ok=True)
    pathlib.Path(args.docids_out).write_text(docids_text)
    index = _place(cached["index"], args)
    _, qids, P, S = _load_queries(args)
    alphas = runloop.parse_alpha_list(args.cm_alphas)
    t0 = time.perf_counter()
    files = runloop.run_alpha_sweep(index, id_lookup, qids, P, S, alphas, args.run_out, k=args.topk, qblock=args.qblock,
                                    alpha_batch=args.alpha_batch, log=logging.info)
    logging.info("Completed %d alpha settings in %.2fs.", len(files), time.perf_counter() - t0)
    return 0


def combine_cached_indexes(cached_list, map_out_path: pathlib.Path, place):
    """Bilingual combined index (onepass_bilingual_mix_hub_custom_lang.py:606-702) without the 17.7 M
    Python-level reconstruct calls: map lines are taken in file order in batches of 20 000, each batch
    sorted by int id (as the reference does), the rows gathered array-at-a-time and appended to ONE flat
    index with ids 0..n-1.  Returns (index, id2doc, base_ids_in_first_seen_order)."""
    dim = cached_list[0]["base_index"].d
    combined = faiss.IndexIDMap(faiss.IndexFlatIP(dim))
    id2doc: List[str] = []
    bases_seen: List[str] = []
    seen = set()
    next_id = 0
    with open(map_out_path, "w", encoding="utf-8") as map_fh:
        print("derived_id\tbase_id\tlang", file=map_fh)
        for cached in cached_list:
            lang, base_index = cached["lang"], cached["base_index"]
            local_ids, derived, bases = [], [], []
            with open(cached["map_path"], "r", encoding="utf-8") as fh:
                next(fh, None)
                for line in fh:
                    parts = line.rstrip("\n").split("\t")
                    if len(parts) < 3:
                        continue
                    try:
                        local_ids.append(int(parts[0]))
                    except ValueError:
                        continue
                    derived.append(parts[1])
                    bases.append(parts[2])
            lid = np.asarray(local_ids, dtype=np.int64)
            order = np.concatenate([b0 + np.argsort(lid[b0 : b0 + RECON_BATCH], kind="stable")
                                    for b0 in range(0, len(lid), RECON_BATCH)]) if len(lid) else np.empty((0,), np.int64)
            src = lid[order]
            rows_all = base_index.reconstruct_n(0, base_index.ntotal)
            identity = len(src) == base_index.ntotal and np.array_equal(src, np.arange(len(src)))
            step = 1 << 18
            for c0 in range(0, len(src), step):
                rows = rows_all[c0 : c0 + step] if identity else rows_all[src[c0 : c0 + step]]
                combined.add_with_ids(rows, np.arange(next_id + c0, next_id + c0 + len(rows), dtype=np.int64))
            for j in order:
                id2doc.append(derived[j])
                map_fh.write(f"{derived[j]}\t{bases[j]}\t{lang}\n")
                if bases[j] not in seen:
                    seen.add(bases[j])
                    bases_seen.append(bases[j])
            next_id += len(src)
    return place(combined), id2doc, bases_seen


def main_bilingual(argv: Optional[Sequence[str]] = None) -> int:
    ap = argparse.ArgumentParser(prog="cmx.cli bilingual", formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    _common(ap)
    ap.add_argument("--langs", required=True, help="Comma-separated document languages, e.g. 'en,zh'")
    ap.add_argument("--outdir", required=True)
    ap.add_argument("--topk", type=int, default=500)
    ap.add_argument("--qblock", type=int, default=128)
    args, ignored = ap.parse_known_args(argv)
    logging.basicConfig(format="%(asctime)s | %(levelname)s | %(message)s", datefmt="%H:%M:%S", level=logging.INFO)
    if ignored:
        logging.info("Ignoring encoder-side flags: %s", " ".join(ignored))
    langs = [v.strip() for v in args.langs.split(",") if v.strip()]
    outdir = pathlib.Path(args.outdir)
    outdir.mkdir(parents=True, exist_ok=True)
    cached_list = []
    for lang in langs:
        c = load_cached_index(pathlib.Path(args.index_root), lang)
        if c is None:
            raise SystemExit(f"No cached index for '{lang}' under {args.index_root}; encoding the corpus is out of scope.")
        cached_list.append(c)
    index, id2doc, bases = combine_cached_indexes(cached_list, outdir / "docid_map.tsv", lambda ix: _place(ix, args))
    pathlib.Path(args.docids_out).parent.mkdir(parents=True, exist_ok=True)
    pathlib.Path(args.docids_out).write_text("\n".join(sorted(set(bases))))
    if index.ntotal == 0:
        raise SystemExit("No documents indexed. Check corpus fields and filters.")
    (l1, l2), qids, P, S = _load_queries(args)
    alphas = runloop.parse_alpha_list(args.cm_alphas)
    tag = "bilingual-mix-" + re.sub(r"\s+", "_", l1.strip()) + "-" + re.sub(r"\s+", "_", l2.strip())
    meta = {"encoder": args.encoder, "repo": args.repo, "langs": langs, "cm_alphas": args.cm_alphas,
            "docids_out": str(args.docids_out), "docid_map": str(outdir / "docid_map.tsv"),
            "max_queries": args.max_queries, "kept_total": int(index.ntotal)}
    files = runloop.run_alpha_sweep_bilingual(index, id2doc, qids, P, S, alphas, outdir, topk=args.topk,
                                              qblock=args.qblock, tag=tag, meta=meta,
                                              alpha_batch=args.alpha_batch, log=logging.info)
    logging.info("All alpha settings completed (%d). Outputs in: %s", len(files), outdir)
    return 0


def main(argv: Optional[Sequence[str]] = None) -> int:
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv or argv[0] not in ("mono", "bilingual"):
        print(__doc__)
        return 2
    return main_mono(argv[1:]) if argv[0] == "mono" else main_bilingual(argv[1:])


if __name__ == "__main__":
    raise SystemExit(main())
