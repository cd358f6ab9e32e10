"""Command-line run scripts for the cached case of the reference's onepass vector-mix jobs.

    python -m cmx.cli mono       <flags of onepass_dense_mix_run_custom_lang.py>
    python -m cmx.cli bilingual  <flags of onepass_bilingual_mix_hub_custom_lang.py>

Same flag names and output files as the reference scripts
(onepass_dense_mix_run_custom_lang.py:388-477, onepass_bilingual_mix_hub_custom_lang.py:429-500) for the
path this repository covers: cached per-language index (``<index_root>/<lang>/{index.faiss,docid_map.tsv,
docids.txt}``) + cached query vectors (``<cache>/<lang>/queries.npz``) -> per-alpha TREC files.  Encoding
corpora or queries (HF encoder + network) is out of scope: when a cache is missing the script stops and says
so instead of loading an encoder.  Encoder-side flags are accepted and ignored.

What differs from the reference's data flow: the vectors of an ``index.faiss`` never become a host-resident
index.  ``index.faiss`` is validated (``inspect_index``) and its float block is streamed by the native loader
(reader threads -> page-locked staging buffers -> HBM) into the GPU index -- or, with ``--gpus N``, each
GPU's row range into its shard.  The bilingual combined index (reference :606-702: 17.7 M Python-level
``reconstruct`` calls into a CPU index that is then cloned to the GPU) is built the same way: both language
files stream into ONE device index, ``docid_map.tsv`` is emitted from the map files' columns, and no row is
materialised in host memory.
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


# The three helpers below reproduce reference behaviour that users see (directory names of the query cache,
# error messages of the --query_tsv parser): onepass_dense_mix_run_custom_lang.py:150-168,311-339.
def sanitize_tag(text: str) -> str:
    """Anything outside [A-Za-z0-9_.-] collapses to '-'; an empty result becomes 'run'."""
    cleaned = re.sub(r"[^A-Za-z0-9_.-]+", "-", text.strip("/")).strip("-")
    return cleaned if cleaned else "run"


def default_query_cache_root(repo: str, encoder: str) -> pathlib.Path:
    """$QUERY_CACHE_ROOT, else <$QUERY_CACHE_ROOT_BASE or ./data>/enc-query-<repo>-<encoder>."""
    explicit = os.environ.get("QUERY_CACHE_ROOT")
    if explicit:
        return pathlib.Path(explicit)
    # (the reference's default base is the directory of its script; a drop-in has no such directory: cwd/data)
    parent = pathlib.Path(os.environ.get("QUERY_CACHE_ROOT_BASE", str(pathlib.Path.cwd() / "data")))
    leaf = "enc-query-{}-{}".format(*(sanitize_tag(name.split("/")[-1]) for name in (repo, encoder)))
    return parent / leaf


def parse_query_specs(query_tsv_args, q_primary, q_secondary) -> List[Tuple[str, pathlib.Path]]:
    """[(lang, path), (lang, path)] from two ``--query_tsv LANG=PATH`` flags, or from --q_en / --q_zh."""
    if query_tsv_args:
        specs = []
        for item in query_tsv_args:
            lang, sep, path = item.partition("=")
            if not sep:
                raise SystemExit(f"--query_tsv expects LANG=PATH, got '{item}'.")
            lang, path = lang.strip(), path.strip()
            if not (lang and path):
                raise SystemExit(f"[ERROR] Bad --query_tsv entry '{item}'.")
            specs.append((lang, pathlib.Path(path)))
    elif q_primary and q_secondary:
        specs = [("en", pathlib.Path(q_primary)), ("zh", pathlib.Path(q_secondary))]
    else:
        raise SystemExit("Provide either --query_tsv twice or both --q_en and --q_zh.")
    if len(specs) != 2:
        raise SystemExit(f"Exactly two query TSV specs required, got {len(specs)}.")
    if specs[0][0] == specs[1][0]:
        raise SystemExit(f"Duplicate language '{specs[0][0]}' in query specs.")
    return specs


def locate_cached_index(index_root: pathlib.Path, lang: str, expected_dim: Optional[int] = None):
    """The cached index of one language WITHOUT loading its vectors: the three files must exist and
    ``index.faiss`` must pass the whole-layout validation (reference: read_index + a probe reconstruct,
    onepass_dense_mix_run_custom_lang.py:238-284)."""
    lang_dir = index_root / lang
    paths = [lang_dir / "index.faiss", lang_dir / "docid_map.tsv", lang_dir / "docids.txt"]
    if not all(p.exists() for p in paths):
        return None
    try:
        info = cio.inspect_index(paths[0])
    except Exception as exc:  # noqa: BLE001 -- the reference logs and treats the cache as unusable
        logging.warning("Failed to read cached index for %s: %s", lang, exc)
        return None
    if expected_dim and info["d"] != expected_dim:
        logging.warning("Cached index dim mismatch for %s: expected %s, found %s.", lang, expected_dim, info["d"])
        return None
    if info["ntotal"] == 0:
        return None
    return {"lang": lang, "index_path": paths[0], "map_path": paths[1], "info": info}


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


def _devices(args) -> List[int]:
    if os.environ.get("CMX_DEVICES"):  # explicit device list, e.g. "0,0,0" = three shards on one GPU (tests)
        return [int(v) for v in os.environ["CMX_DEVICES"].split(",")]
    if args.gpus and args.gpus > 1:
        return list(range(args.gpus))
    return [int(args.faiss_gpu_id)]


def build_device_index(segments, d: int, devices: Sequence[int], ids: Optional[np.ndarray] = None, log=None):
    """``IndexIDMap`` over a GPU-resident flat index holding the concatenation of ``segments`` -- a list of
    (path, byte offset of the float block, rows) -- streamed from the files by the native loader.  One device:
    ``GpuIndexFlatIP``; several: ``IndexShardsIP`` (row i -> shard floor(i*G/n)), every shard loading its own row
    range, all shards at once.  ``ids`` = the user ids (default 0..n-1)."""
    total = sum(int(r) for _, _, r in segments)
    t0 = time.perf_counter()
    if len(devices) == 1:
        flat = faiss.GpuIndexFlatIP(d, device=devices[0])
        flat.reserveMemory(total)
        for path, off, rows in segments:
            flat.add_from_file(path, off, rows)
    else:
        flat = faiss.IndexShardsIP(d, devices)
        flat.add_from_files(segments)
    dt = time.perf_counter() - t0
    if log:
        log(f"Streamed {total} vectors ({4e-9 * d * total:.1f} GB) into {len(devices)} GPU(s) in {dt:.2f}s "
            f"({4e-9 * d * total / max(dt, 1e-9):.2f} GB/s)")
    out = faiss.IndexIDMap(faiss.IndexFlatIP(d))  # the wrapper wants an empty index at construction
    out.index = flat
    out._ids = [np.arange(total, dtype=np.int64) if ids is None else np.ascontiguousarray(ids, dtype=np.int64)]
    return out


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
    cached = locate_cached_index(pathlib.Path(args.index_root), doc_lang)
    if cached is None:
        raise SystemExit(f"No cached index for '{doc_lang}' under {args.index_root}; encoding the corpus is out of scope "
                         "(build it with the reference's encode_multilingual_corpus.py).")
    id_lookup, docids_text, n_kept = cio.read_docid_table(cached["map_path"])
    logging.info("Cached index ready for %s: %d vectors", doc_lang, n_kept)
    pathlib.Path(args.docids_out).parent.mkdir(parents=True, exist_ok=True)
    pathlib.Path(args.docids_out).write_text(docids_text)
    info = cached["info"]
    index = build_device_index([(cached["index_path"], info["vec_offset"], info["ntotal"])], info["d"], _devices(args),
                               ids=cio.read_index_ids(cached["index_path"], info), log=logging.info)
    _, qids, P, S = _load_queries(args)
    alphas = runloop.parse_alpha_list(args.cm_alphas)
    t0 = time.perf_counter()
    files = runloop.run_alpha_sweep(index, id_lookup, qids, P, S, alphas, args.run_out, k=args.topk, qblock=args.qblock,
                                    alpha_batch=args.alpha_batch, log=logging.info)
    logging.info("Completed %d alpha settings in %.2fs.", len(files), time.perf_counter() - t0)
    return 0


def _batch_sorted_order(local_ids: np.ndarray) -> np.ndarray:
    """The reference takes map lines in file order in batches of RECON_BATCH and sorts each batch by int id
    (stable) before adding it (:639-641): the permutation of the map lines that gives the combined row order."""
    if not len(local_ids):
        return np.empty((0,), np.int64)
    return np.concatenate([b0 + np.argsort(local_ids[b0 : b0 + RECON_BATCH], kind="stable")
                           for b0 in range(0, len(local_ids), RECON_BATCH)])


def combine_cached_indexes(cached_list, map_out_path: pathlib.Path, devices: Sequence[int], log=None):
    """Bilingual combined index (onepass_bilingual_mix_hub_custom_lang.py:606-702) on the device: same row order
    (map lines in batches of 20 000, each sorted by int id), ids 0..n-1, same ``docid_map.tsv`` -- but the rows go
    from the language files straight into HBM (native loader), or, when the map does not list the rows in storage
    order, through a device-side gather; the map columns are handled as arrays (pyarrow) when the files are clean.
    Returns (index, id2doc, text of sorted unique base ids for --docids_out)."""
    import pyarrow as pa
    import pyarrow.compute as pc

    d = cached_list[0]["info"]["d"]
    segments, gathers = [], []  # (path, offset, rows) in combined order; per language: None or the row permutation
    derived_parts, base_parts = [], []
    with open(map_out_path, "wb") as map_fh:
        map_fh.write(b"derived_id\tbase_id\tlang\n")
        for cached in cached_list:
            lang, info = cached["lang"], cached["info"]
            cols = cio.read_docid_columns(cached["map_path"])
            if cols is None:  # irregular file: the literal per-line reader
                lid, derived_l, base_l = [], [], []
                with open(cached["map_path"], "r", encoding="utf-8") as fh:
                    next(fh, None)
                    for line in fh:
                        parts = line.rstrip("\n").split("\t")
                        if len(parts) < 3:
                            continue
                        try:
                            lid.append(int(parts[0]))
                        except ValueError:
                            continue
                        derived_l.append(parts[1])
                        base_l.append(parts[2])
                lid = np.asarray(lid, dtype=np.int64)
                derived, base = pa.array(derived_l, type=pa.string()), pa.array(base_l, type=pa.string())
            else:
                lid, derived, base = cols
            if len(lid) and (lid.min() < 0 or lid.max() >= info["ntotal"]):
                raise SystemExit(f"docid_map.tsv of '{lang}' names row {int(lid.max())} but index.faiss holds {info['ntotal']} rows")
            order = _batch_sorted_order(lid)
            src = lid[order]
            identity = len(src) == info["ntotal"] and bool(np.array_equal(src, np.arange(len(src))))
            in_order = len(order) == 0 or bool(np.array_equal(order, np.arange(len(order))))
            segments.append((cached["index_path"], info["vec_offset"], info["ntotal"] if identity else len(src)))
            gathers.append(None if identity else src)
            if not in_order:
                idx = pa.array(order)
                derived, base = derived.take(idx), base.take(idx)
            derived_parts.append(derived)
            base_parts.append(base)
            if len(derived):
                lines = pc.binary_join_element_wise(derived, base, pa.scalar(lang + "\n"), "\t")
                map_fh.write(_arrow_string_bytes(lines))
    if all(gx is None for gx in gathers):
        index = build_device_index(segments, d, devices, log=log)
    else:
        if len(devices) > 1:
            raise SystemExit("a docid_map.tsv that does not list the rows in storage order needs the single-GPU build "
                             "(device-side gather); rerun with --gpus 1")
        flat = faiss.GpuIndexFlatIP(d, device=devices[0])
        flat.reserveMemory(sum(r for _, _, r in segments))
        for (path, off, rows), cached, src in zip(segments, cached_list, gathers):
            if src is None:
                flat.add_from_file(path, off, rows)
                continue
            tmp = faiss.GpuIndexFlatIP(d, device=devices[0])
            tmp.add_from_file(path, off, cached["info"]["ntotal"])
            for c0 in range(0, len(src), 1 << 20):
                flat.add_gather(tmp, src[c0 : c0 + (1 << 20)])
            del tmp
        index = faiss.IndexIDMap(faiss.IndexFlatIP(d))
        index.index = flat
        index._ids = [np.arange(flat.ntotal, dtype=np.int64)]
    id2doc = pa.concat_arrays([p.cast(pa.string()) for p in derived_parts]) if derived_parts else pa.array([], type=pa.string())
    all_base = pa.concat_arrays([p.cast(pa.string()) for p in base_parts]) if base_parts else pa.array([], type=pa.string())
    uniq = pc.unique(all_base)
    uniq = uniq.take(pc.sort_indices(uniq))
    docids_text = _arrow_string_bytes(pc.binary_join_element_wise(uniq, pa.scalar("\n"), "")).decode("utf-8")
    return index, id2doc, docids_text[:-1] if docids_text.endswith("\n") else docids_text


def _arrow_string_bytes(arr) -> bytes:
    """The concatenated utf-8 bytes of a pyarrow string array."""
    if hasattr(arr, "combine_chunks"):
        arr = arr.combine_chunks()
    if len(arr) == 0:
        return b""
    off = np.frombuffer(arr.buffers()[1], dtype=np.int32, count=len(arr) + 1, offset=arr.offset * 4)
    return arr.buffers()[2].to_pybytes()[off[0] : off[-1]]


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
        c = locate_cached_index(pathlib.Path(args.index_root), lang, cached_list[0]["info"]["d"] if cached_list else None)
        if c is None:
            raise SystemExit(f"No cached index for '{lang}' under {args.index_root}; encoding the corpus is out of scope.")
        cached_list.append(c)
    index, id2doc, docids_text = combine_cached_indexes(cached_list, outdir / "docid_map.tsv", _devices(args), log=logging.info)
    pathlib.Path(args.docids_out).parent.mkdir(parents=True, exist_ok=True)
    pathlib.Path(args.docids_out).write_text(docids_text)
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
