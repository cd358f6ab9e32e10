"""File formats on either side of the hot path.

* ``index.faiss``   -- FAISS's on-disk layout for ``IndexIDMap(IndexFlatIP)`` /
  ``IndexFlatIP`` as written by ``faiss.write_index``
  (encode_multilingual_corpus.py:471) and read by ``faiss.read_index``
  (onepass_dense_mix_run_custom_lang.py:250).  faiss itself is not installable
  here, so this follows the published format of faiss 1.8 (index_write.cpp /
  index_read.cpp) and is verified only by round trips -- best effort until checked
  against a file written by real faiss:
      "IxMp" | header | <nested index> | u64 n | n x i64 id_map
      "IxFI" | header | u64 count (= ntotal*d, in 4-byte units) | count x f32
      header = i32 d | i64 ntotal | i64 dummy(1<<20) | i64 dummy(1<<20) |
               u8 is_trained | i32 metric_type (0 = inner product)
* ``docid_map.tsv`` -- header ``int_id\\tderived_id\\tbase_id\\tlang``
  (encode_multilingual_corpus.py:474-478; readers
  onepass_dense_mix_run_custom_lang.py:632-644,
  onepass_bilingual_mix_hub_custom_lang.py:629-641)
* ``queries.npz``   -- ``np.savez_compressed(qids=<unicode array>, vecs=[n,d])``
  (onepass_dense_mix_run_custom_lang.py:196-235, cache_queries_for_mix.py:166-176)
"""
from __future__ import annotations

import pathlib
import struct
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

_HDR = struct.Struct("<iqqqBi")  # d, ntotal, dummy, dummy, is_trained, metric_type
_DUMMY = 1 << 20
_CHUNK_ROWS = 1 << 16


def _write_header(fh, d: int, ntotal: int) -> None:
    fh.write(_HDR.pack(int(d), int(ntotal), _DUMMY, _DUMMY, 1, 0))


def _read_header(fh) -> Tuple[int, int]:
    raw = fh.read(_HDR.size)
    if len(raw) != _HDR.size:
        raise RuntimeError("read_index: truncated header")
    d, ntotal, _d1, _d2, _trained, metric = _HDR.unpack(raw)
    if d <= 0 or ntotal < 0:
        raise RuntimeError(f"read_index: bad header (d={d}, ntotal={ntotal})")
    if metric != 0:
        raise RuntimeError(f"read_index: only METRIC_INNER_PRODUCT (0) is supported, file has metric {metric}")
    return d, ntotal


def write_index(index, path) -> None:
    from . import faiss as F

    path = pathlib.Path(path)
    with open(path, "wb") as fh:
        if isinstance(index, F.IndexIDMap):
            fh.write(b"IxMp")
            _write_header(fh, index.d, index.ntotal)
            _write_flat(fh, index.index)
            ids = np.ascontiguousarray(index.id_map, dtype="<i8")
            fh.write(struct.pack("<Q", ids.shape[0]))
            fh.write(ids.tobytes())
        else:
            _write_flat(fh, index)


def _write_flat(fh, flat) -> None:
    n, d = flat.ntotal, flat.d
    fh.write(b"IxFI")
    _write_header(fh, d, n)
    fh.write(struct.pack("<Q", n * d))
    for i0 in range(0, n, _CHUNK_ROWS):
        rows = flat.reconstruct_n(i0, min(_CHUNK_ROWS, n - i0))
        fh.write(np.ascontiguousarray(rows, dtype="<f4").tobytes())


def read_index(path, device: Optional[int] = None):
    """Returns a host-resident IndexFlatIP / IndexIDMap(IndexFlatIP) like faiss.read_index
    (rows are streamed in chunks; promote with index_cpu_to_gpu / search)."""
    from . import faiss as F

    path = pathlib.Path(path)
    with open(path, "rb") as fh:
        fourcc = fh.read(4)
        if fourcc == b"IxMp":
            d, ntotal = _read_header(fh)
            flat = _read_flat(fh, F)
            raw = fh.read(8)
            if len(raw) != 8:
                raise RuntimeError("read_index: truncated id_map")
            (n_ids,) = struct.unpack("<Q", raw)
            if n_ids != flat.ntotal or flat.d != d:
                raise RuntimeError(f"read_index: id_map has {n_ids} ids for {flat.ntotal} rows")
            ids = np.fromfile(fh, dtype="<i8", count=n_ids)
            if ids.shape[0] != n_ids:
                raise RuntimeError("read_index: truncated id_map")
            out = F.IndexIDMap(F.IndexFlatIP(d))
            out.index = flat
            out._ids = [ids.astype(np.int64, copy=False)]
            return out
        if fourcc == b"IxFI":
            fh.seek(0)
            fh.read(4)
            return _read_flat(fh, F, fourcc_read=True)
        raise RuntimeError(f"read_index: unsupported index type {fourcc!r} (only IxMp / IxFI)")


def inspect_index(path) -> dict:
    """Validate the layout of an ``index.faiss`` file WITHOUT loading the vectors: fourccs, both
    headers, the float count, the id_map length and the total file size must all agree.  Returns
    ``{"kind": "IxMp"|"IxFI", "d", "ntotal", "vec_offset", "ids_offset"|None}``."""
    path = pathlib.Path(path)
    size = path.stat().st_size
    with open(path, "rb") as fh:
        kind = fh.read(4)
        if kind not in (b"IxMp", b"IxFI"):
            raise RuntimeError(f"read_index: unsupported index type {kind!r} (only IxMp / IxFI)")
        outer = None
        if kind == b"IxMp":
            outer = _read_header(fh)
            inner = fh.read(4)
            if inner != b"IxFI":
                raise RuntimeError(f"read_index: nested index {inner!r} is not IndexFlatIP (IxFI)")
        d, ntotal = _read_header(fh)
        if outer is not None and outer != (d, ntotal):
            raise RuntimeError(f"read_index: IxMp header {outer} disagrees with the nested IxFI header {(d, ntotal)}")
        raw = fh.read(8)
        if len(raw) != 8:
            raise RuntimeError("read_index: truncated vector block")
        (count,) = struct.unpack("<Q", raw)
        if count != ntotal * d:
            raise RuntimeError(f"read_index: vector block holds {count} floats, expected {ntotal}*{d}")
        vec_offset = fh.tell()
        end = vec_offset + 4 * count
        ids_offset = None
        if kind == b"IxMp":
            fh.seek(end)
            raw = fh.read(8)
            if len(raw) != 8:
                raise RuntimeError("read_index: truncated id_map")
            (n_ids,) = struct.unpack("<Q", raw)
            if n_ids != ntotal:
                raise RuntimeError(f"read_index: id_map has {n_ids} ids for {ntotal} rows")
            ids_offset = end + 8
            end = ids_offset + 8 * n_ids
        if size != end:
            raise RuntimeError(f"read_index: file is {size} bytes, layout needs {end}")
    return {"kind": kind.decode(), "d": d, "ntotal": ntotal, "vec_offset": vec_offset, "ids_offset": ids_offset}


LAST_LOAD: dict = {}  # rows / bytes / read_s / wall_s of the last native (device-sink) load


def sink_is_device(sink) -> bool:
    """True for sinks whose rows live in HBM (GpuIndexFlatIP, Shard, ShardedIndex over the CUDA engine)."""
    inner = getattr(sink, "local", sink)
    return hasattr(inner, "add_from_file") and (hasattr(inner, "_h") or hasattr(inner, "_shard"))


def stream_index_rows(path, sink, row0: int = 0, row1: Optional[int] = None, chunk_rows: int = _CHUNK_ROWS,
                      info: Optional[dict] = None) -> np.ndarray:
    """Feed rows [row0, row1) of an ``index.faiss`` file to ``sink.add(x)`` in chunks of ``chunk_rows``
    (256 MB at d = 1024) and return their user ids -- the loader for indexes that do not fit, or
    should not transit, host memory: a GPU index (``GpuIndexFlatIP``) or one rank's row range of a
    ``ShardedIndex`` (``sink.add_local``).  faiss.read_index + index_cpu_to_gpu
    (onepass_dense_mix_run_custom_lang.py:250,658-664) needs the whole 36 GB twice in host RAM."""
    info = info or inspect_index(path)
    d, ntotal = info["d"], info["ntotal"]
    row1 = ntotal if row1 is None else int(row1)
    if not (0 <= row0 <= row1 <= ntotal):
        raise RuntimeError(f"stream_index_rows: bad row range [{row0}, {row1}) of {ntotal}")
    # a device sink takes the whole range in one native call: reader threads + pinned staging buffers
    # overlap the file reads with the PCIe copies (cmx_index_add_from_file)
    native = getattr(getattr(sink, "local", sink), "add_from_file", None)
    if native is not None and sink_is_device(sink):
        if row1 > row0:
            stats = native(path, info["vec_offset"] + 4 * d * row0, row1 - row0)
            LAST_LOAD.update(rows=row1 - row0, bytes=4 * d * (row1 - row0), read_s=stats[0], wall_s=stats[1])
        if info["ids_offset"] is None:
            return np.arange(row0, row1, dtype=np.int64)
        with open(path, "rb") as fh:
            fh.seek(info["ids_offset"] + 8 * row0)
            ids = np.fromfile(fh, dtype="<i8", count=row1 - row0)
        if ids.shape[0] != row1 - row0:
            raise RuntimeError("read_index: truncated id_map")
        return ids.astype(np.int64, copy=False)
    add = getattr(sink, "add_local", None) or sink.add
    with open(path, "rb") as fh:
        fh.seek(info["vec_offset"] + 4 * d * row0)
        r = row0
        while r < row1:
            n = min(chunk_rows, row1 - r)
            rows = np.fromfile(fh, dtype="<f4", count=n * d)
            if rows.shape[0] != n * d:
                raise RuntimeError("read_index: truncated vector block")
            add(rows.reshape(n, d))
            r += n
        if info["ids_offset"] is None:
            return np.arange(row0, row1, dtype=np.int64)
        fh.seek(info["ids_offset"] + 8 * row0)
        ids = np.fromfile(fh, dtype="<i8", count=row1 - row0)
        if ids.shape[0] != row1 - row0:
            raise RuntimeError("read_index: truncated id_map")
    return ids.astype(np.int64, copy=False)


def read_index_ids(path, info: Optional[dict] = None) -> Optional[np.ndarray]:
    """The id_map of an IxMp file (None for a bare IxFI file) without touching the vectors."""
    info = info or inspect_index(path)
    if info["ids_offset"] is None:
        return None
    with open(path, "rb") as fh:
        fh.seek(info["ids_offset"])
        ids = np.fromfile(fh, dtype="<i8", count=info["ntotal"])
    if ids.shape[0] != info["ntotal"]:
        raise RuntimeError("read_index: truncated id_map")
    return ids.astype(np.int64, copy=False)


def read_index_to_gpu(path, device: int = 0):
    """``index.faiss`` -> ``IndexIDMap(GpuIndexFlatIP)`` (or a bare ``GpuIndexFlatIP`` for an IxFI file)
    without a host-resident copy: the equivalent of read_index + index_cpu_to_gpu."""
    from . import faiss as F

    info = inspect_index(path)
    flat = F.GpuIndexFlatIP(info["d"], device=device)
    out = F.IndexIDMap(flat) if info["kind"] == "IxMp" else flat
    flat.reserveMemory(info["ntotal"])
    ids = stream_index_rows(path, flat, info=info)
    if info["kind"] == "IxMp":
        out._ids = [ids]
    return out


def _read_block(fh, count: int) -> np.ndarray:
    """``count`` little-endian float32 values at the file position of ``fh``; large blocks with concurrent preads
    (``cmx_read_file``: several GB/s from page cache or a striped disk instead of one sequential read)."""
    if count * 4 >= (64 << 20):
        import os

        from . import _lib

        rows = np.empty((count,), dtype="<f4")
        pos = fh.tell()
        if os.fstat(fh.fileno()).st_size - pos < 4 * count:
            raise RuntimeError("read_index: truncated vector block")
        _lib.check(_lib.lib().cmx_read_file(os.fsencode(fh.name), pos, 4 * count, rows.ctypes.data, 0))
        fh.seek(pos + 4 * count)
        return rows
    rows = np.fromfile(fh, dtype="<f4", count=count)
    if rows.shape[0] != count:
        raise RuntimeError("read_index: truncated vector block")
    return rows


def _read_flat(fh, F, fourcc_read: bool = False):
    if not fourcc_read:
        fourcc = fh.read(4)
        if fourcc != b"IxFI":
            raise RuntimeError(f"read_index: nested index {fourcc!r} is not IndexFlatIP (IxFI)")
    d, ntotal = _read_header(fh)
    raw = fh.read(8)
    if len(raw) != 8:
        raise RuntimeError("read_index: truncated vector block")
    (count,) = struct.unpack("<Q", raw)
    if count != ntotal * d:
        raise RuntimeError(f"read_index: vector block holds {count} floats, expected {ntotal}*{d}")
    flat = F.IndexFlatIP(d)
    rows = _read_block(fh, count)
    if ntotal:
        flat._blocks = [rows.reshape(ntotal, d)]
        flat._n = ntotal
    return flat


# ---- docid_map.tsv ------------------------------------------------------------
def write_docid_map(path, int_ids: Sequence[int], derived: Sequence[str], base: Sequence[str], lang: str) -> None:
    with open(path, "w", encoding="utf-8") as fh:
        print("int_id\tderived_id\tbase_id\tlang", file=fh)
        for i, de, b in zip(int_ids, derived, base):
            print(f"{i}\t{de}\t{b}\t{lang}", file=fh)


def read_docid_map(path) -> Tuple[Dict[int, str], List[str], List[str]]:
    """-> (id_lookup {int_id: base_id}, kept [base_id...], derived [derived_id...]);
    malformed lines are skipped exactly as the reference readers do."""
    id_lookup: Dict[int, str] = {}
    kept: List[str] = []
    derived: List[str] = []
    with open(path, "r", encoding="utf-8") as fh:
        next(fh, None)  # header
        for line in fh:
            parts = line.rstrip("\n").split("\t")
            if len(parts) < 3:
                continue
            try:
                local_id = int(parts[0])
            except ValueError:
                continue
            id_lookup[local_id] = str(parts[2])
            kept.append(str(parts[2]))
            derived.append(parts[1])
    return id_lookup, kept, derived


def read_docid_table(path):
    """Fast path of ``read_docid_map`` for the run loops: -> (DocTable int_id -> base_id, text of
    ``"\\n".join(sorted(set(base_ids)))`` for --docids_out, number of rows).

    A clean 8.8 M-line map (plain decimal ids, a fixed number of tab-separated columns, no CR, no
    repeated int ids) is parsed by pyarrow's multi-threaded CSV reader straight into the
    one-buffer-plus-offsets layout the C formatter reads -- ~5 s instead of ~19 s of per-line
    Python.  Anything else (or no pyarrow) goes through ``read_docid_map``, which follows the
    reference readers literally (onepass_dense_mix_run_custom_lang.py:632-644)."""
    from .runloop import DocTable, StrTable

    try:
        table = _docid_table_arrow(path, DocTable, StrTable)
        if table is not None:
            return table
    except Exception:  # any surprise in the file: take the literal reader
        pass
    id_lookup, kept, _ = read_docid_map(path)
    return DocTable(id_lookup), "\n".join(sorted(set(kept))), len(kept)


def read_docid_columns(path):
    """(int ids [n] int64 numpy, derived ids, base ids) of a CLEAN ``docid_map.tsv`` -- the two id columns as
    pyarrow string arrays, no per-line Python objects -- or None when the file needs the literal reader
    (``read_docid_map``): CR line ends, ragged columns, a first column that is not a plain decimal."""
    try:
        import pyarrow as pa
        import pyarrow.compute as pc
        import pyarrow.csv as pcsv

        with open(path, "rb") as fh:
            header = fh.readline()
            if b"\r" in header:
                return None
            names = header.decode("utf-8").rstrip("\n").split("\t")
            if len(names) < 3 or len(set(names)) != len(names):
                return None
        tbl = pcsv.read_csv(
            path,
            parse_options=pcsv.ParseOptions(delimiter="\t", quote_char=False, escape_char=False, newlines_in_values=False),
            convert_options=pcsv.ConvertOptions(column_types={n: pa.string() for n in names}, strings_can_be_null=False,
                                                include_columns=names[:3]))
        if tbl.num_rows == 0:
            return None
        first, derived, base = (tbl.column(i).combine_chunks() for i in range(3))
        for col in (derived, base):
            if col.null_count or pc.any(pc.match_substring(col, "\r")).as_py():
                return None
        if not pc.all(pc.match_substring_regex(first, "^-?[0-9]+$")).as_py():
            return None
        return pc.cast(first, pa.int64()).to_numpy(zero_copy_only=False), derived, base
    except Exception:
        return None


def _docid_table_arrow(path, DocTable, StrTable):
    import pyarrow as pa
    import pyarrow.compute as pc
    import pyarrow.csv as pcsv

    with open(path, "rb") as fh:
        header = fh.readline()
        if b"\r" in header:
            return None
        names = header.decode("utf-8").rstrip("\n").split("\t")
        if len(names) < 3 or len(set(names)) != len(names):
            return None
    tbl = pcsv.read_csv(
        path,
        parse_options=pcsv.ParseOptions(delimiter="\t", quote_char=False, escape_char=False, newlines_in_values=False),
        convert_options=pcsv.ConvertOptions(column_types={n: pa.string() for n in names}, strings_can_be_null=False,
                                            include_columns=[names[0], names[2]]))
    n = tbl.num_rows
    if n == 0:
        return None
    base = tbl.column(1).combine_chunks()
    if base.null_count or pc.any(pc.match_substring(base, "\r")).as_py():
        return None
    first = tbl.column(0).combine_chunks()
    # plain ASCII decimal only: anything int() would accept beyond that (spaces, '+', '_', other scripts) is rare
    if not pc.all(pc.match_substring_regex(first, "^-?[0-9]+$")).as_py():
        return None
    ids = pc.cast(first, pa.int64()).to_numpy(zero_copy_only=False)
    if ids.shape[0] > 1 and not np.all(ids[1:] > ids[:-1]):  # not strictly increasing: order / duplicates matter
        order = np.argsort(ids, kind="stable")
        ids_sorted = ids[order]
        if np.any(ids_sorted[1:] == ids_sorted[:-1]):
            return None  # a repeated int id: the dict semantics (last wins) live in read_docid_map
        base_sorted = base.take(pa.array(order))
    else:
        ids_sorted, base_sorted = ids, base
    off = np.frombuffer(base_sorted.buffers()[1], dtype=np.int32, count=n + 1, offset=base_sorted.offset * 4).astype(np.int64)
    data = base_sorted.buffers()[2]
    raw = data.to_pybytes()[off[0]:off[-1]] if data is not None else b""
    st = StrTable.__new__(StrTable)
    st.n, st.buf, st.off = n, raw, np.ascontiguousarray(off - off[0])
    docs = DocTable.__new__(DocTable)
    docs.keys = None if (ids_sorted[0] == 0 and ids_sorted[-1] == n - 1) else np.ascontiguousarray(ids_sorted)
    docs.table = st
    srt = base.take(pc.sort_indices(base))
    if n > 1:
        keep = np.concatenate([[True], pc.not_equal(srt.slice(1), srt.slice(0, n - 1)).to_numpy(zero_copy_only=False)])
        if not keep.all():
            srt = srt.filter(pa.array(keep))
    joined = pc.binary_join(pa.ListArray.from_arrays(pa.array([0, len(srt)], type=pa.int32()), srt), "\n")[0].as_py()
    return docs, joined, n


# ---- queries.npz ----------------------------------------------------------------
def save_query_cache(cache_dir, lang: str, qids: Sequence[str], vecs) -> None:
    """vecs: [n,d] array in qid order, or a dict qid -> row."""
    lang_dir = pathlib.Path(cache_dir) / lang
    lang_dir.mkdir(parents=True, exist_ok=True)
    if isinstance(vecs, dict):
        if not vecs:
            return
        vecs = np.stack([vecs[q] for q in qids if q in vecs], axis=0)
    np.savez_compressed(lang_dir / "queries.npz", qids=np.array(list(qids)), vecs=np.asarray(vecs))


def load_query_cache(cache_dir, lang: str, qids: Sequence[str]) -> Optional[np.ndarray]:
    """-> vecs [n,d] float32 in the requested order, or None when the cache is absent or
    its qids differ from the request (same acceptance rule as the reference)."""
    cache_file = pathlib.Path(cache_dir) / lang / "queries.npz"
    if not cache_file.exists():
        return None
    try:
        data = np.load(cache_file)
        cached = [str(x) for x in data["qids"].tolist()]
        if cached != list(qids):
            return None
        vecs = data["vecs"].astype(np.float32, copy=False)
        if vecs.shape[0] != len(qids):
            return None
        return np.ascontiguousarray(vecs)
    except Exception:
        return None


def read_queries_tsv(path, qid_field: str = "id", text_field: str = "text") -> List[Tuple[str, str]]:
    """qid<TAB>text rows; an optional header line naming the two fields is skipped and a
    malformed line ends the job (onepass_dense_mix_run_custom_lang.py:72-91)."""
    rows: List[Tuple[str, str]] = []
    with open(path, "r", encoding="utf-8") as fh:
        for ln, line in enumerate(fh, 1):
            line = line.rstrip("\n")
            if not line:
                continue
            parts = line.split("\t")
            if ln == 1 and len(parts) >= 2:
                if parts[0].lower().startswith(qid_field.lower()) and parts[1].lower().startswith(text_field.lower()):
                    continue
            if len(parts) < 2:
                raise SystemExit(f"[ERROR] Bad queries TSV line #{ln}: {line}")
            rows.append((parts[0], parts[1]))
    return rows
