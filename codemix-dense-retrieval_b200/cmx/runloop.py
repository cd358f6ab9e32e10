"""The per-alpha run loop of the onepass vector-mix scripts, array-at-a-time.

Mirrors (same names, argument meaning, file names and byte-for-byte file content):
  * mono:      onepass_dense_mix_run_custom_lang.py:844-890  -> ``run_alpha_sweep``
  * bilingual: onepass_bilingual_mix_hub_custom_lang.py:901-962 -> ``run_alpha_sweep_bilingual``
  * ``format_alpha`` :304-308, ``parse_alpha_list`` :287-301,
    ``collapse_run_max`` onepass_bilingual_mix_hub_custom_lang.py:165-181

What changed relative to the reference loop: the ~7000 per-query ``safe_mix`` GPU
round-trips per alpha become one fused kernel (``search_mixed``), and the ~700k
f-strings per alpha (7 M at k=1000) are formatted by a multi-threaded C++ routine in
libcmx.so (``cmx_trec_mono`` / ``cmx_trec_bilingual``, csrc/trec_text.cpp).  Output
files are byte-identical to what the reference loop would write for the same (D, I).
"""
from __future__ import annotations

import json
import os
import pathlib
import time
from typing import Dict, List, Optional, Sequence

import numpy as np


# ---- alpha labels -------------------------------------------------------------
# Interface kept verbatim from the reference (onepass_dense_mix_run_custom_lang.py:287-308): the labels are
# the run-file names -- the resume keys of run_all_vector_pairs.sh:329-360 -- and the messages are what
# a user of the reference sees.  Both are pinned by tests/golden/text_golden.json.
def format_alpha(alpha: float) -> str:
    """'0' / '1' for integer-valued alphas, else 4 decimals without trailing zeros ('0.1', '0.25')."""
    nearest = round(alpha)
    if abs(alpha - nearest) < 1e-8:
        return str(int(nearest))
    return f"{alpha:.4f}".rstrip("0").rstrip(".") or "0"


def parse_alpha_list(alpha_str: str) -> List[float]:
    """Comma-separated floats; blanks between commas are skipped, anything else that is not a float ends the job."""
    if not alpha_str:
        raise SystemExit("--cm_alphas must contain at least one value.")
    alphas: List[float] = []
    for token in (t.strip() for t in alpha_str.split(",")):
        if token:
            try:
                alphas.append(float(token))
            except ValueError as exc:
                raise SystemExit(f"[ERROR] Could not parse alpha '{token}': {exc}") from exc
    if not alphas:
        raise SystemExit("No valid alpha values parsed from --cm_alphas.")
    return alphas


class StrTable:
    """A list of strings as ONE utf-8 buffer + n+1 offsets (the layout the C formatter reads)."""

    def __init__(self, items):
        enc = [str(v).encode("utf-8") for v in items]
        self.n = len(enc)
        self.buf = b"".join(enc)
        off = np.zeros(self.n + 1, dtype=np.int64)
        if self.n:
            np.cumsum(np.fromiter((len(b) for b in enc), dtype=np.int64, count=self.n), out=off[1:])
        self.off = off


class DocTable:
    """int id -> document-id string; ids missing from the table print as ``str(id)``
    (the reference's ``id_lookup.get(int(doc), str(doc))``)."""

    def __init__(self, id_lookup):
        if isinstance(id_lookup, DocTable):
            self.keys, self.table = id_lookup.keys, id_lookup.table
            return
        if isinstance(id_lookup, dict):
            keys = np.fromiter(id_lookup.keys(), dtype=np.int64, count=len(id_lookup))
            names = list(id_lookup.values())
            order = np.argsort(keys, kind="stable")
            keys = keys[order]
            names = [names[i] for i in order]
            if keys.size and keys[0] == 0 and keys[-1] == keys.size - 1 and np.all(np.diff(keys) == 1):
                keys = None  # ids are exactly 0..n-1: direct indexing
        else:  # sequence: position = id
            names, keys = list(id_lookup), None
        self.keys = keys
        self.table = StrTable(names)

    def lookup(self, ids) -> list:
        """Python-level equivalent (small inputs / tests)."""
        out = []
        buf, off = self.table.buf, self.table.off
        for v in np.asarray(ids, dtype=np.int64).reshape(-1).tolist():
            if self.keys is None:
                pos = v if 0 <= v < self.table.n else -1
            else:
                pos = int(np.searchsorted(self.keys, v))
                pos = pos if pos < self.keys.size and self.keys[pos] == v else -1
            out.append(buf[off[pos]:off[pos + 1]].decode("utf-8") if pos >= 0 else str(v))
        return out


def _host_f32_i64(D, I):
    D = np.ascontiguousarray(_to_host(D), dtype=np.float32)
    I = np.ascontiguousarray(_to_host(I), dtype=np.int64)
    assert D.ndim == 2 and D.shape == I.shape
    return D, I


def _take_bytes(ptr, length) -> bytes:
    import ctypes as C

    from . import _lib

    try:
        return C.string_at(ptr.value, length.value)
    finally:
        _lib.lib().cmx_free_text(ptr)


def mono_trec_bytes(qids: Sequence[str], D, I, docs, tag: str = "onepass-cm", nthreads: int = 0) -> bytes:
    """Text of one mono run file: lines ``qid\\tQ0\\tdoc\\trank\\tscore(.4f)\\ttag`` joined
    with newlines, no trailing newline (onepass_dense_mix_run_custom_lang.py:879-888).
    Formatted by the multi-threaded C routine ``cmx_trec_mono``."""
    import ctypes as C

    from . import _lib

    D, I = _host_f32_i64(D, I)
    nq, k = D.shape
    if nq == 0:
        return b""
    docs = docs if isinstance(docs, DocTable) else DocTable(docs)
    q = qids if isinstance(qids, StrTable) else StrTable(qids)
    assert q.n == nq
    out, n = C.c_void_p(), C.c_int64(0)
    keys_ptr = None if docs.keys is None else docs.keys.ctypes.data
    _lib.check(_lib.lib().cmx_trec_mono(D.ctypes.data, I.ctypes.data, nq, k, q.buf, q.off.ctypes.data, docs.table.buf,
                                        docs.table.off.ctypes.data, keys_ptr, docs.table.n, tag.encode("utf-8"),
                                        int(nthreads), C.byref(out), C.byref(n)))
    return _take_bytes(out, n)


def _tmp_name(path: pathlib.Path) -> pathlib.Path:
    return path.with_name(path.name + f".tmp{os.getpid()}")


def write_mono_trec(path, qids: Sequence[str], D, I, docs, tag: str = "onepass-cm", nthreads: int = 0) -> int:
    """``mono_trec_bytes`` written to ``path`` without assembling the text in memory: the C
    formatter's threads pwrite their parts into a temporary file that is then renamed over
    ``path`` (all-or-nothing, like the reference's single write_text).  Returns the file size."""
    import ctypes as C

    from . import _lib

    path = pathlib.Path(path)
    D, I = _host_f32_i64(D, I)
    nq, k = D.shape
    if nq == 0:
        _atomic_write(path, b"")
        return 0
    docs = docs if isinstance(docs, DocTable) else DocTable(docs)
    q = qids if isinstance(qids, StrTable) else StrTable(qids)
    assert q.n == nq
    n = C.c_int64(0)
    keys_ptr = None if docs.keys is None else docs.keys.ctypes.data
    tmp = _tmp_name(path)
    try:
        _lib.check(_lib.lib().cmx_trec_mono_file(D.ctypes.data, I.ctypes.data, nq, k, q.buf, q.off.ctypes.data, docs.table.buf,
                                                 docs.table.off.ctypes.data, keys_ptr, docs.table.n, tag.encode("utf-8"),
                                                 int(nthreads), os.fsencode(tmp), C.byref(n)))
        os.replace(tmp, path)
    finally:
        if tmp.exists():
            tmp.unlink()
    return int(n.value)


def _strtable_from_arrow(arr) -> "StrTable":
    """StrTable over a pyarrow string array (no per-string Python objects)."""
    n = len(arr)
    off = np.frombuffer(arr.buffers()[1], dtype=np.int32, count=n + 1, offset=arr.offset * 4).astype(np.int64)
    data = arr.buffers()[2]
    st = StrTable.__new__(StrTable)
    st.n = n
    st.buf = data.to_pybytes()[off[0]:off[-1]] if (data is not None and n) else b""
    st.off = np.ascontiguousarray(off - off[0])
    return st


class BaseTable:
    """derived ids ``base#lang`` by position, plus their base-id grouping for the collapse
    (``base = did.split('#')[0]``, bases numbered in first-seen order)."""

    def __init__(self, id2doc: Sequence[str]):
        if type(id2doc).__module__.startswith("pyarrow"):  # a string array (cmx.cli builds id2doc as one): no Python strings at all
            if self._init_arrow(id2doc):
                return
            id2doc = id2doc.to_pylist()
        names = id2doc if isinstance(id2doc, list) else list(id2doc)
        if len(names) >= 4096 and self._init_arrow(names):  # 17.7 M ids: ~9 s instead of ~29 s of per-name Python
            return
        self.docs = StrTable(names)
        code_of, bases, codes = {}, [], np.empty(len(names), dtype=np.int32)
        for i, nme in enumerate(names):
            b = nme.split("#", 1)[0]
            c = code_of.get(b)
            if c is None:
                c = code_of[b] = len(bases)
                bases.append(b)
            codes[i] = c
        self.codes = codes
        self.bases = StrTable(bases)

    def _init_arrow(self, names) -> bool:
        try:
            import pyarrow as pa
            import pyarrow.compute as pc

            arr = names if isinstance(names, (pa.Array, pa.ChunkedArray)) else pa.array(names, type=pa.string())
            if isinstance(arr, pa.ChunkedArray):
                arr = arr.combine_chunks()
            if pa.types.is_large_string(arr.type):
                arr = arr.cast(pa.string())
            if not isinstance(arr, pa.Array) or arr.null_count:
                return False
            base = pc.list_element(pc.split_pattern(arr, "#", max_splits=1), 0)
            enc = pc.dictionary_encode(base)  # dictionary in order of first appearance
            codes = enc.indices.to_numpy(zero_copy_only=False).astype(np.int32)
            docs, bases = _strtable_from_arrow(arr), _strtable_from_arrow(enc.dictionary)
        except Exception:
            return False
        self.docs, self.codes, self.bases = docs, np.ascontiguousarray(codes), bases
        return True


def bilingual_bytes(qids: Sequence[str], D, I, id2doc, tag: str, nthreads: int = 0):
    """(raw run text, collapsed run text) of one bilingual alpha
    (onepass_bilingual_mix_hub_custom_lang.py:950-958 and collapse_run_max :165-181):
    raw lines ``qid Q0 did rank score(.6f) tag\\n`` for ids inside [0, len(id2doc)) -- others are
    skipped and keep their rank number; collapse groups a query's hits by
    ``base = did.split('#',1)[0]``, takes the max of the 6-decimal ROUNDED scores, sorts stably
    descending (ties keep first-seen order), re-ranks, tag ``bilingual-mix``."""
    import ctypes as C

    from . import _lib

    D, I = _host_f32_i64(D, I)
    nq, k = D.shape
    if nq == 0:
        return b"", b""
    qids = list(qids)
    if len(set(qids)) != len(qids):  # the reference merges rows of a repeated qid: rare, keep it exact
        raw = bilingual_raw_text_py(qids, D, I, id2doc, tag)
        return raw.encode("utf-8"), collapse_text_py(raw.splitlines(keepends=True)).encode("utf-8")
    bt = id2doc if isinstance(id2doc, BaseTable) else BaseTable(id2doc)
    q = StrTable(qids)
    raw, nraw, col, ncol = C.c_void_p(), C.c_int64(0), C.c_void_p(), C.c_int64(0)
    _lib.check(_lib.lib().cmx_trec_bilingual(D.ctypes.data, I.ctypes.data, nq, k, q.buf, q.off.ctypes.data, bt.docs.buf,
                                             bt.docs.off.ctypes.data, bt.docs.n, bt.codes.ctypes.data, bt.bases.buf,
                                             bt.bases.off.ctypes.data, bt.bases.n, tag.encode("utf-8"), int(nthreads),
                                             C.byref(raw), C.byref(nraw), C.byref(col), C.byref(ncol)))
    return _take_bytes(raw, nraw), _take_bytes(col, ncol)


def collapse_max_device(D, I, codes_dev, ndocs: int):
    """Collapse-by-base-id (fuse = max) of one alpha's (D, I) [nq, k] on the GPU (``cmx_collapse_max``): D, I are CUDA
    tensors or pinned host tensors (read in place through their device alias), ``codes_dev`` the int32 base code of
    every corpus row on the device.  Returns host numpy arrays (codes [nq,k], val6 [nq,k], counts [nq]) for
    ``write_bilingual_trec(..., groups=...)``, or None when a score cannot be carried exactly on the device."""
    import ctypes as C

    import torch

    from . import _lib
    from .engine import host_register

    nq, k = int(D.shape[0]), int(D.shape[1])
    dev = codes_dev.device

    def ptr(t):
        if t.is_cuda:
            return int(t.data_ptr())
        assert t.is_pinned(), "host results must be page-locked to be read by the kernel in place"
        return host_register(t.data_ptr(), t.numel() * t.element_size())

    code = torch.empty((nq, k), dtype=torch.int32, device=dev)
    val6 = torch.empty((nq, k), dtype=torch.int64, device=dev)
    cnt = torch.empty((nq,), dtype=torch.int32, device=dev)
    needs_host = C.c_int(0)
    _lib.check(_lib.lib().cmx_collapse_max(ptr(D), ptr(I), nq, k, int(codes_dev.data_ptr()), int(ndocs), int(code.data_ptr()),
                                           int(val6.data_ptr()), int(cnt.data_ptr()), C.byref(needs_host), dev.index,
                                           int(torch.cuda.current_stream(dev).cuda_stream)))
    if needs_host.value:
        return None
    return code.cpu().numpy(), val6.cpu().numpy(), cnt.cpu().numpy()


def write_bilingual_trec(raw_path, col_path, qids: Sequence[str], D, I, id2doc, tag: str, nthreads: int = 0, groups=None):
    """``bilingual_bytes`` written straight to the raw and the collapsed run file (temporary
    names, then renamed).  ``groups`` = the collapsed lists of ``collapse_max_device`` (else the
    formatter groups on the host).  Returns the two file sizes."""
    import ctypes as C

    from . import _lib

    raw_path, col_path = pathlib.Path(raw_path), pathlib.Path(col_path)
    D, I = _host_f32_i64(D, I)
    nq, k = D.shape
    qids = list(qids)
    if nq == 0 or len(set(qids)) != len(qids):  # repeated qids: the exact (slow) Python path
        raw, col = bilingual_bytes(qids, D, I, id2doc, tag, nthreads)
        _atomic_write(raw_path, raw)
        _atomic_write(col_path, col)
        return len(raw), len(col)
    bt = id2doc if isinstance(id2doc, BaseTable) else BaseTable(id2doc)
    q = StrTable(qids)
    nraw, ncol = C.c_int64(0), C.c_int64(0)
    t_raw, t_col = _tmp_name(raw_path), _tmp_name(col_path)
    try:
        if groups is not None:
            gc, gv, gn = (np.ascontiguousarray(groups[0], dtype=np.int32), np.ascontiguousarray(groups[1], dtype=np.int64),
                          np.ascontiguousarray(groups[2], dtype=np.int32))
            assert gc.shape == (nq, k) and gv.shape == (nq, k) and gn.shape == (nq,)
            _lib.check(_lib.lib().cmx_trec_bilingual_file_pre(
                D.ctypes.data, I.ctypes.data, nq, k, q.buf, q.off.ctypes.data, bt.docs.buf, bt.docs.off.ctypes.data, bt.docs.n,
                bt.codes.ctypes.data, bt.bases.buf, bt.bases.off.ctypes.data, bt.bases.n, gc.ctypes.data, gv.ctypes.data,
                gn.ctypes.data, tag.encode("utf-8"), int(nthreads), os.fsencode(t_raw), os.fsencode(t_col), C.byref(nraw),
                C.byref(ncol)))
        else:
            _lib.check(_lib.lib().cmx_trec_bilingual_file(D.ctypes.data, I.ctypes.data, nq, k, q.buf, q.off.ctypes.data, bt.docs.buf,
                                                          bt.docs.off.ctypes.data, bt.docs.n, bt.codes.ctypes.data, bt.bases.buf,
                                                          bt.bases.off.ctypes.data, bt.bases.n, tag.encode("utf-8"), int(nthreads),
                                                          os.fsencode(t_raw), os.fsencode(t_col), C.byref(nraw), C.byref(ncol)))
        os.replace(t_raw, raw_path)
        os.replace(t_col, col_path)
    finally:
        for t in (t_raw, t_col):
            if t.exists():
                t.unlink()
    return int(nraw.value), int(ncol.value)


def bilingual_raw_text_py(qids, D, I, id2doc, tag: str) -> str:
    names = id2doc.docs if isinstance(id2doc, BaseTable) else None
    lst = list(id2doc) if names is None else [names.buf[names.off[i]:names.off[i + 1]].decode() for i in range(names.n)]
    out = []
    for r, qid in enumerate(qids):
        for rank, (sc, ix) in enumerate(zip(D[r].tolist(), I[r].tolist()), 1):
            if ix < 0 or ix >= len(lst):
                continue
            out.append(f"{qid} Q0 {lst[ix]} {rank} {sc:.6f} {tag}\n")
    return "".join(out)


def collapse_text_py(raw_lines) -> str:
    by_q: Dict[str, Dict[str, float]] = {}
    for line in raw_lines:
        line = line.strip()
        if not line:
            continue
        qid, _, did, _rk, sc, _tag = line.split()
        base = did.split("#", 1)[0]
        score = float(sc)
        g = by_q.setdefault(qid, {})
        if base not in g or score > g[base]:
            g[base] = score
    out = []
    for qid, groups in by_q.items():
        for rank, (base, val) in enumerate(sorted(groups.items(), key=lambda kv: kv[1], reverse=True), 1):
            out.append(f"{qid} Q0 {base} {rank} {val:.6f} bilingual-mix\n")
    return "".join(out)


def collapse_run_max(in_run, out_run) -> None:
    """Text-to-text form with the reference's signature (used when only the raw file exists)."""
    with open(in_run, "r", encoding="utf-8") as f:
        text = collapse_text_py(f)
    with open(out_run, "w", encoding="utf-8") as out:
        out.write(text)


def _atomic_write(path: pathlib.Path, text) -> None:
    tmp = path.with_name(path.name + f".tmp{os.getpid()}")
    if isinstance(text, bytes):
        tmp.write_bytes(text)
    else:
        tmp.write_text(text, encoding="utf-8")
    os.replace(tmp, path)


def _to_host(a):
    return a.detach().cpu().numpy() if hasattr(a, "detach") else np.asarray(a)


# ---- the sweeps -------------------------------------------------------------------
def _gpu_of(index) -> Optional[int]:
    """CUDA device of a single-GPU cmx.faiss index (through IndexIDMap wrappers), else None."""
    seen = 0
    while index is not None and seen < 4:
        if hasattr(index, "getDevice"):
            try:
                return int(index.getDevice())
            except Exception:
                return None
        index = getattr(index, "index", None)
        seen += 1
    return None


class _SweepIO:
    """Host<->device traffic of a sweep.  On a single-GPU index the cached query matrices are
    uploaded ONCE (the reference re-uploads a 4 KB row per query and alpha: safe_mix), every
    alpha's (D, I) lands in one of two pinned host slots (full PCIe rate, no per-alpha
    allocation), and the slot of alpha i is being formatted while alpha i+1 is searched.
    Any other index (sharded, host) gets the arrays as they are."""

    def __init__(self, index, P, S):
        self.P, self.S, self.slots, self.turn, self.torch = P, S, [None, None], 0, None
        dev = _gpu_of(index)
        if dev is None:
            return
        try:
            import torch
        except Exception:
            return
        if not torch.cuda.is_available():
            return
        self.torch = torch
        self.dev = torch.device("cuda", dev)
        self.P = torch.as_tensor(np.ascontiguousarray(_to_host(P), dtype=np.float32)).to(self.dev)
        self.S = torch.as_tensor(np.ascontiguousarray(_to_host(S), dtype=np.float32)).to(self.dev)

    def to_host(self, D, I):
        torch = self.torch
        if torch is None or not (hasattr(D, "is_cuda") and D.is_cuda):
            return _to_host(D), _to_host(I)
        slot = self.slots[self.turn]
        if slot is None or tuple(slot[0].shape) != tuple(D.shape):
            slot = (torch.empty(tuple(D.shape), dtype=torch.float32).pin_memory(),
                    torch.empty(tuple(I.shape), dtype=torch.int64).pin_memory())
            self.slots[self.turn] = slot
        self.turn ^= 1
        slot[0].copy_(D, non_blocking=True)
        slot[1].copy_(I, non_blocking=True)
        torch.cuda.current_stream(self.dev).synchronize()
        return slot[0].numpy(), slot[1].numpy()


LAST_SWEEP: dict = {}  # bookkeeping of the last bilingual sweep: how many alphas were collapsed on the device / on the host


class _DeviceCollapse:
    """Per-alpha collapse-by-base-id on the GPU for the bilingual sweep: the base code of every corpus row is uploaded
    once; each alpha's (D, I) is grouped where it lies (device tensors of a single-GPU index, or the pinned host
    buffer a sharded index delivers, read through its device alias).  ``run`` returns one entry per alpha: the host
    arrays for ``write_bilingual_trec(groups=...)`` or None (the formatter groups on the host: no CUDA, pageable
    results, or a score the device keys cannot carry exactly)."""

    def __init__(self, table: "BaseTable"):
        self.table, self.codes, self.torch = table, {}, None
        try:
            import torch

            if torch.cuda.is_available():
                self.torch = torch
        except Exception:
            pass

    def run(self, D, I, nA: int):
        torch = self.torch
        if torch is None:
            return [None] * nA
        try:
            Dt = D if isinstance(D, torch.Tensor) else torch.from_numpy(np.asarray(D))
            It = I if isinstance(I, torch.Tensor) else torch.from_numpy(np.asarray(I))
            if not Dt.is_cuda and not (Dt.is_pinned() and It.is_pinned()):
                return [None] * nA
            if It.dtype != torch.int64 or Dt.dtype != torch.float32 or Dt.shape[-1] > 2048:
                return [None] * nA
            dev = Dt.device if Dt.is_cuda else torch.device("cuda", torch.cuda.current_device())
            codes = self.codes.get(dev)
            if codes is None:
                codes = self.codes[dev] = torch.from_numpy(self.table.codes).to(dev)
            out = [collapse_max_device(Dt[a].contiguous(), It[a].contiguous(), codes, self.table.docs.n) for a in range(nA)]
            LAST_SWEEP["device_collapse"] = LAST_SWEEP.get("device_collapse", 0) + sum(g is not None for g in out)
            return out
        except Exception as exc:  # any surprise: the host grouping is always right
            LAST_SWEEP["device_collapse_error"] = repr(exc)
            return [None] * nA


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
    from concurrent.futures import ThreadPoolExecutor

    outdir = pathlib.Path(outdir)
    outdir.mkdir(parents=True, exist_ok=True)
    docs = DocTable(id_lookup)
    qtab = StrTable(list(qids))
    written: List[pathlib.Path] = []
    alphas = [float(a) for a in alphas]
    # two-stage pipeline: while the formatter threads (C, GIL released) turn alpha i into text and
    # write its file, the GPU already searches alpha i+1
    sio = _SweepIO(index, P, S)
    with ThreadPoolExecutor(max_workers=1) as pool:
        pending = None
        for a0 in range(0, len(alphas), max(1, alpha_batch)):
            group = alphas[a0 : a0 + max(1, alpha_batch)]
            t0 = time.perf_counter()
            D, I = index.search_mixed(sio.P, sio.S, group, k)
            if pending is not None:  # the slot about to be overwritten belongs to the alpha before the pending one
                pending.result()
            D, I = sio.to_host(D, I)
            t1 = time.perf_counter()

            def emit(group=group, D=D, I=I, dt=t1 - t0):
                for gi, alpha in enumerate(group):
                    label = format_alpha(alpha)
                    run_path = outdir / f"cm-alpha-{label}.trec"
                    write_mono_trec(run_path, qtab, D[gi], I[gi], docs, tag)
                    written.append(run_path)
                    if log:
                        log(f"Run saved: {run_path}  ({qtab.n} queries, alpha={label}, search {dt:.3f}s)")

            pending = pool.submit(emit)
        if pending is not None:
            pending.result()
    return written


def run_alpha_sweep_bilingual(index, id2doc: Sequence[str], qids: Sequence[str], P, S, alphas: Sequence[float],
                              outdir, topk: int = 500, qblock: int = 128, tag: str = "bilingual-mix",
                              meta: Optional[dict] = None, alpha_batch: int = 1, log=None) -> List[pathlib.Path]:
    """Bilingual sweep: per alpha write ``cm-alpha-<label>_raw.trec`` (derived ``base#lang``
    ids), the collapsed ``cm-alpha-<label>.trec`` (max over languages) and
    ``cm-alpha-<label>_meta.json``."""
    from concurrent.futures import ThreadPoolExecutor

    outdir = pathlib.Path(outdir)
    outdir.mkdir(parents=True, exist_ok=True)
    qids = list(qids)
    table = BaseTable(id2doc)
    written: List[pathlib.Path] = []
    alphas = [float(a) for a in alphas]
    ntotal, dim = int(index.ntotal), int(index.d)
    sio = _SweepIO(index, P, S)
    LAST_SWEEP.clear()
    collapse = _DeviceCollapse(table)
    with ThreadPoolExecutor(max_workers=1) as pool:  # text of alpha i overlaps the search of alpha i+1
        pending = None
        for a0 in range(0, len(alphas), max(1, alpha_batch)):
            group = alphas[a0 : a0 + max(1, alpha_batch)]
            D, I = index.search_mixed(sio.P, sio.S, group, topk)
            groups = collapse.run(D, I, len(group))  # collapse-by-base-id where the results lie
            if pending is not None:
                pending.result()
            D, I = sio.to_host(D, I)

            def emit(group=group, D=D, I=I, groups=groups):
                for gi, alpha in enumerate(group):
                    label = format_alpha(alpha)
                    set_name = f"cm-alpha-{label}"
                    run_raw = outdir / f"{set_name}_raw.trec"
                    run_base = outdir / f"{set_name}.trec"
                    write_bilingual_trec(run_raw, run_base, qids, D[gi], I[gi], table, tag, groups=groups[gi])
                    m = dict(meta or {})
                    m.update({"alpha": label, "runs": {"raw": str(run_raw), "base": str(run_base)},
                              "index": {"type": "IndexIDMap(IndexFlatIP)", "size": ntotal, "dim": dim},
                              "topk": int(topk), "qblock": int(qblock)})
                    _atomic_write(outdir / f"{set_name}_meta.json", json.dumps(m, indent=2))
                    written.append(run_base)
                    if log:
                        log(f"Completed set '{set_name}' -> {run_raw.name} , {run_base.name}")

            pending = pool.submit(emit)
        if pending is not None:
            pending.result()
    return written
