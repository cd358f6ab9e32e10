"""The subset of the Python ``faiss`` API the reference's run scripts use, backed by
libcmx.so on B200.  ``import cmx.faiss as faiss`` is the drop-in.

Symbols used by the reference and provided here:
``IndexFlatIP``, ``IndexIDMap`` (+ ``.index``, ``.d``, ``.ntotal``, ``.id_map``),
``add`` / ``add_with_ids`` / ``search`` / ``reconstruct``, ``read_index`` /
``write_index``, ``downcast_index``, ``StandardGpuResources``,
``index_cpu_to_gpu``, ``index_gpu_to_cpu``; plus the faiss-named multi-GPU
entry points ``index_cpu_to_all_gpus`` / ``index_cpu_to_gpus_list`` and
``GpuIndexFlatIP``.

Reference call sites: onepass_dense_mix_run_custom_lang.py:250,265-269,604,
658-664,719,878; onepass_bilingual_mix_hub_custom_lang.py:558,644-646,931-936,950;
encode_multilingual_corpus.py:367-373,440,469-471; onepass_dense_run.py:305-312,427.

Semantics kept: ``search(x, k) -> (D [n,k] float32 descending, I [n,k] int64)``,
``-1`` / lowest-float padding when k > ntotal, AssertionError on a dimension
mismatch, RuntimeError for engine errors, ``add`` copies, ``index_cpu_to_gpu``
leaves the CPU index valid and independent.

There is no CPU search: a "CPU" ``IndexFlatIP`` keeps its rows in host memory
(like faiss) but ``search`` on it promotes the rows to the default GPU and runs
the CUDA path; with no GPU / no libcmx.so it raises RuntimeError.
"""
from __future__ import annotations

import os
from concurrent.futures import ThreadPoolExecutor
from typing import List, Optional, Sequence

import numpy as np

from . import _lib
from .engine import Shard, _as_f32_2d, _is_torch, merge_topk

try:
    import torch
except Exception:  # pragma: no cover
    torch = None

METRIC_INNER_PRODUCT = 0
METRIC_L2 = 1


def get_num_gpus() -> int:
    return _lib.device_count()


def _default_device() -> int:
    return int(os.environ.get("CMX_DEVICE", "0"))


class StandardGpuResources:
    """Placeholder with faiss' name: libcmx sizes its own bounded workspace per index
    (candidate buffers + operand planes), so there is no temp-memory arena to manage.
    Constructing it verifies that a GPU and the CUDA library are usable."""

    def __init__(self):
        if get_num_gpus() < 1:
            raise RuntimeError("cmx.faiss.StandardGpuResources: no usable CUDA device")

    def setTempMemory(self, nbytes):  # noqa: N802 (faiss naming)
        pass

    def noTempMemory(self):  # noqa: N802
        pass

    def setDefaultNullStreamAllDevices(self):  # noqa: N802
        pass


class GpuClonerOptions:
    def __init__(self):
        self.useFloat16 = False
        self.shard = False


GpuMultipleClonerOptions = GpuClonerOptions


class _Searchable:
    """search-path selection shared by all index classes ('auto' | 'stream' | 'tensor')."""

    path = "auto"


class GpuIndexFlatIP(_Searchable):
    """Flat inner-product index resident in one GPU's HBM (faiss.GpuIndexFlatIP)."""

    is_trained = True
    metric_type = METRIC_INNER_PRODUCT

    def __init__(self, res_or_d, d: Optional[int] = None, config=None, device: Optional[int] = None):
        if d is None:  # GpuIndexFlatIP(d)
            d = res_or_d
        if device is None:
            device = getattr(config, "device", None)
        if device is None:
            device = _default_device()
        self.d = int(d)
        self._shard = Shard(self.d, int(device))

    # faiss attribute / method names
    @property
    def ntotal(self) -> int:
        return self._shard.ntotal

    def getDevice(self) -> int:  # noqa: N802
        return self._shard.device

    def reserveMemory(self, n: int) -> None:  # noqa: N802
        self._shard.reserve(n)

    def add(self, x) -> None:
        self._shard.add(x)

    def reset(self) -> None:
        self._shard.reset()

    def add_from_file(self, path, offset: int, n: int, nthreads: int = 0):
        """Extension: rows straight from a file's float32 block into HBM (see ``Shard.add_from_file``)."""
        return self._shard.add_from_file(path, offset, n, nthreads)

    def add_gather(self, src: "GpuIndexFlatIP", rows) -> None:
        self._shard.add_gather(src._shard, rows)

    def search(self, x, k: int):
        return self._shard.search(x, k, path=self.path)

    def search_mixed(self, P, S, alphas, k: int):
        return self._shard.search_mixed(P, S, alphas, k, path=self.path)

    def reconstruct(self, i: int, out=None):
        i = int(i)
        if not 0 <= i < self.ntotal:
            raise RuntimeError(f"reconstruct: index {i} out of range [0, {self.ntotal})")
        if out is None:
            return self._shard.reconstruct_n(i, 1)[0]
        self._shard.reconstruct_n(i, 1, out.reshape(1, -1))
        return out

    def reconstruct_n(self, i0: int = 0, n: int = -1, out=None):
        if n < 0:
            n = self.ntotal - i0
        return self._shard.reconstruct_n(i0, n, out)

    def last_stats(self) -> dict:
        return self._shard.last_stats()


class IndexFlatIP(_Searchable):
    """faiss.IndexFlatIP: rows kept in host memory; search runs on the GPU."""

    is_trained = True
    metric_type = METRIC_INNER_PRODUCT

    def __init__(self, d: int):
        self.d = int(d)
        self._blocks: List[np.ndarray] = []
        self._n = 0
        self._gpu: Optional[GpuIndexFlatIP] = None  # lazily promoted copy
        self._gpu_rows = 0

    @property
    def ntotal(self) -> int:
        return self._n

    def add(self, x) -> None:
        if _is_torch(x):
            x = x.detach().cpu().numpy()
        x = _as_f32_2d(x, self.d, "add(x)")
        self._blocks.append(np.array(x, dtype=np.float32, copy=True))  # faiss copies
        self._n += x.shape[0]

    def reset(self) -> None:
        self._blocks, self._n = [], 0
        self._gpu, self._gpu_rows = None, 0

    def _rows(self) -> np.ndarray:
        if len(self._blocks) != 1:
            self._blocks = [np.concatenate(self._blocks, axis=0)] if self._blocks else [np.empty((0, self.d), np.float32)]
        return self._blocks[0]

    def reconstruct(self, i: int, out=None):
        i = int(i)
        if not 0 <= i < self._n:
            raise RuntimeError(f"reconstruct: index {i} out of range [0, {self._n})")
        row = self._rows()[i]
        if out is None:
            return row.copy()
        out[...] = row
        return out

    def reconstruct_n(self, i0: int = 0, n: int = -1, out=None):
        if n < 0:
            n = self._n - i0
        rows = self._rows()[i0 : i0 + n]
        if out is None:
            return rows.copy()
        out[...] = rows
        return out

    def _promote(self, device: Optional[int] = None) -> GpuIndexFlatIP:
        if self._gpu is None or (device is not None and self._gpu.getDevice() != device):
            self._gpu = GpuIndexFlatIP(self.d, device=device)
            self._gpu_rows = 0
        if self._gpu_rows < self._n:
            if self._gpu_rows == 0:
                self._gpu.reserveMemory(self._n)
            self._gpu.add(self._rows()[self._gpu_rows :])
            self._gpu_rows = self._n
        self._gpu.path = self.path
        return self._gpu

    def search(self, x, k: int):
        return self._promote().search(x, k)

    def search_mixed(self, P, S, alphas, k: int):
        return self._promote().search_mixed(P, S, alphas, k)


def _map_ids(I, id_map_ext, dev_cache: dict):
    """I = id_map[I], -1 stays -1 (faiss IndexIDMap::search).  ``id_map_ext`` is the map with a
    trailing -1, so that one gather does both (index -1 wraps to the sentinel)."""
    if _is_torch(I):
        if I.is_cuda:
            idm = dev_cache.get(I.device)
            if idm is None:  # uploaded once per device, not per search
                idm = dev_cache[I.device] = torch.from_numpy(id_map_ext).to(I.device)
            return idm[I]
        return torch.from_numpy(id_map_ext[I.numpy()])
    return id_map_ext[I]


class IndexIDMap(_Searchable):
    """faiss.IndexIDMap: user ids on top of a flat index (ids live host-side)."""

    is_trained = True
    metric_type = METRIC_INNER_PRODUCT

    def __init__(self, index):
        if index.ntotal != 0:
            raise RuntimeError("IndexIDMap: index must be empty on input")
        self.index = index
        self.d = index.d
        self._version = 0
        self._ids = []            # list of int64 arrays (property: every assignment bumps _version)
        self._map_sig = None      # _version the cached translation state below was built from
        self._map_identity = True
        self._map_ext = None
        self._map_dev: dict = {}

    @property
    def ntotal(self) -> int:
        return self.index.ntotal

    @property
    def _ids(self) -> List[np.ndarray]:
        return self._id_parts

    @_ids.setter
    def _ids(self, parts) -> None:
        self._id_parts = parts
        self._version += 1

    def _translate(self, I):
        """User ids of the rows in I.  The common map of the reference -- ids 0..n-1 in row order
        (``add_with_ids(x, np.arange(start, start + n))``) -- is recognised once and costs nothing
        per search; any other map is one gather (on the device for device results)."""
        if self._version != self._map_sig:
            idm = self.id_map
            self._map_identity = bool(idm.shape[0] == 0 or (idm[0] == 0 and idm[-1] == idm.shape[0] - 1
                                                            and np.array_equal(idm, np.arange(idm.shape[0], dtype=np.int64))))
            self._map_ext = None if self._map_identity else np.concatenate([idm, np.array([-1], dtype=np.int64)])
            self._map_dev = {}
            self._map_sig = self._version
        if self._map_identity:
            return I
        return _map_ids(I, self._map_ext, self._map_dev)

    @property
    def id_map(self) -> np.ndarray:
        if len(self._ids) != 1:
            self._ids = [np.concatenate(self._ids)] if self._ids else [np.empty((0,), np.int64)]
        return self._ids[0]

    def add(self, x):
        raise RuntimeError("add not implemented for IndexIDMap: use add_with_ids")  # faiss behaviour

    def add_with_ids(self, x, ids) -> None:
        ids = np.ascontiguousarray(ids, dtype=np.int64).reshape(-1)
        n = x.shape[0] if hasattr(x, "shape") and len(x.shape) == 2 else 1
        assert ids.shape[0] == n, "add_with_ids: one id per row"
        self.index.add(x)
        self._ids.append(ids.copy())
        self._version += 1

    def reset(self) -> None:
        self.index.reset()
        self._ids = []

    def search(self, x, k: int):
        self.index.path = self.path
        D, I = self.index.search(x, k)
        return D, self._translate(I)

    def search_mixed(self, P, S, alphas, k: int):
        self.index.path = self.path
        D, I = self.index.search_mixed(P, S, alphas, k)
        return D, self._translate(I)

    def reconstruct(self, key, out=None):
        # faiss: only IndexIDMap2 can reconstruct by user id; the reference probes the
        # *base* index instead (onepass_dense_mix_run_custom_lang.py:264-269)
        raise RuntimeError("reconstruct not implemented for IndexIDMap (use the base index / IndexIDMap2)")


def downcast_index(index):
    """faiss.downcast_index: our Python objects already have their concrete type."""
    return index


def _clone_flat_to_gpu(index, device: int) -> GpuIndexFlatIP:
    if isinstance(index, GpuIndexFlatIP):
        g = GpuIndexFlatIP(index.d, device=device)
        n = index.ntotal
        if n:
            g.reserveMemory(n)
            step = max(1, (1 << 28) // (4 * index.d))
            for i0 in range(0, n, step):
                g.add(index.reconstruct_n(i0, min(step, n - i0)))
        return g
    if isinstance(index, IndexFlatIP):
        g = GpuIndexFlatIP(index.d, device=device)
        if index.ntotal:
            g.reserveMemory(index.ntotal)
            g.add(index._rows())
        return g
    raise RuntimeError(f"index_cpu_to_gpu: unsupported index type {type(index).__name__}")


def index_cpu_to_gpu(res, device: int, index, options=None):
    """Clone onto ONE GPU; the source index stays valid and independent."""
    if isinstance(index, IndexIDMap):
        out = IndexIDMap(GpuIndexFlatIP(index.d, device=device))
        out.index = _clone_flat_to_gpu(index.index, device)
        out._ids = [index.id_map.copy()]
        return out
    return _clone_flat_to_gpu(index, int(device))


def index_gpu_to_cpu(index):
    if isinstance(index, IndexIDMap):
        out = IndexIDMap(IndexFlatIP(index.d))
        out.index = index_gpu_to_cpu(index.index)
        out._ids = [index.id_map.copy()]
        return out
    if isinstance(index, IndexShardsIP):
        cpu = IndexFlatIP(index.d)
        for sh in index.shards:
            if sh.ntotal:
                cpu.add(sh.reconstruct_n(0, sh.ntotal))
        return cpu
    if isinstance(index, GpuIndexFlatIP):
        cpu = IndexFlatIP(index.d)
        n = index.ntotal
        step = max(1, (1 << 28) // (4 * index.d))
        for i0 in range(0, n, step):
            cpu.add(index.reconstruct_n(i0, min(step, n - i0)))
        return cpu
    if isinstance(index, IndexFlatIP):
        return index
    raise RuntimeError(f"index_gpu_to_cpu: unsupported index type {type(index).__name__}")


class IndexShardsIP(_Searchable):
    """Row-sharded flat IP index over several GPUs driven from ONE process
    (faiss.index_cpu_to_all_gpus(..., shard=True) analogue; new capability, the
    reference never shards).  Shard g holds the contiguous rows
    [bounds[g], bounds[g+1]).  A search runs the SAME step as the one-process-per-GPU
    ``cmx.dist.ShardedIndex`` -- one thread per GPU executes it over a ``LocalFabric``
    (peer access between the devices, event barriers between their streams): every
    shard uploads 1/G of the query vectors and replicates it over NVLink, the shards
    exchange their global k-th approximate scores before the exact rescoring, and each
    merge kernel writes its slice of the queries straight into ONE pinned host buffer.
    Result == the single-GPU result exactly (same tie order).

    The returned (D, I) are numpy views of that pinned buffer; two buffers alternate,
    so a result stays valid until the search after the next one (what the run loops'
    "format alpha i while alpha i+1 is searched" pipeline needs)."""

    is_trained = True
    metric_type = METRIC_INNER_PRODUCT

    def __init__(self, d: int, devices: Sequence[int]):
        from .dist import LocalGroup

        self.d = int(d)
        self.devices = [int(v) for v in devices]
        self.shards = [GpuIndexFlatIP(d, device=dev) for dev in self.devices]
        self._group = LocalGroup(self.devices)
        # one persistent thread per device (its CUDA device is set once)
        self._threads = [ThreadPoolExecutor(max_workers=1, initializer=self._bind, initargs=(dev,)) for dev in self.devices]
        self._views = None      # per-shard ShardedIndex over the current row counts
        self._views_key = None
        self._out = {}          # (rank, slot, shape) -> HostResult
        self._turn = 0
        self.two_phase_used = False

    @staticmethod
    def _bind(dev: int) -> None:
        if torch is not None and torch.cuda.is_available():
            torch.cuda.set_device(dev)

    @property
    def ntotal(self) -> int:
        return sum(s.ntotal for s in self.shards)

    @property
    def bounds(self) -> List[int]:
        b = [0]
        for s in self.shards:
            b.append(b[-1] + s.ntotal)
        return b

    @staticmethod
    def split(n: int, g: int) -> List[int]:
        """row i -> shard floor(i*g/n): contiguous, sizes differ by at most one."""
        return [(n * j) // g for j in range(g + 1)]

    def add(self, x) -> None:
        """Appends to the LAST shard (keeps global row order = add order)."""
        self.shards[-1].add(x)

    def add_sharded(self, rows) -> None:
        assert self.ntotal == 0, "add_sharded needs an empty index"
        b = self.split(rows.shape[0], len(self.shards))
        for g, sh in enumerate(self.shards):
            if b[g + 1] > b[g]:
                sh.reserveMemory(b[g + 1] - b[g])
                sh.add(rows[b[g] : b[g + 1]])

    def add_from_files(self, segments) -> None:
        """``segments`` = [(path, byte offset of the float32 block, rows), ...] in global row order: every shard
        streams ITS row range (row i -> shard floor(i*G/n)) from the files with the native loader, all shards
        at once -- the multi-GPU form of read_index + index_cpu_to_gpu, nothing staged in host memory."""
        assert self.ntotal == 0, "add_from_files needs an empty index"
        total = sum(int(r) for _, _, r in segments)
        b = self.split(total, len(self.shards))

        def load(g):
            sh = self.shards[g]
            if b[g + 1] > b[g]:
                sh.reserveMemory(b[g + 1] - b[g])
            start = 0
            for path, off, rows in segments:
                lo, hi = max(b[g], start), min(b[g + 1], start + int(rows))
                if hi > lo:
                    sh.add_from_file(path, int(off) + 4 * self.d * (lo - start), hi - lo)
                start += int(rows)

        for f in [t.submit(load, g) for g, t in enumerate(self._threads)]:
            f.result()

    def reset(self) -> None:
        for s in self.shards:
            s.reset()

    def reconstruct(self, i: int, out=None):
        b = self.bounds
        for g, sh in enumerate(self.shards):
            if b[g] <= i < b[g + 1]:
                return sh.reconstruct(i - b[g], out)
        raise RuntimeError(f"reconstruct: index {i} out of range [0, {b[-1]})")

    def _fanout(self, fn):
        """fn(rank, view) on every device's thread, concurrently (the step has cross-shard barriers)."""
        def guarded(g):
            try:
                return fn(g, self._views[g])
            except BaseException:
                self._group.tbar.abort()  # the other shards would wait for this one forever
                raise

        futs = [t.submit(guarded, g) for g, t in enumerate(self._threads)]
        errs, outs = [], []
        for f in futs:
            try:
                outs.append(f.result())
            except Exception as exc:  # noqa: BLE001
                errs.append(exc)
        if errs:
            self._group.tbar.reset()
            import threading

            real = [e for e in errs if not isinstance(e, threading.BrokenBarrierError)]
            raise (real or errs)[0]
        return outs

    def _ensure_views(self) -> None:
        from .dist import LocalFabric, ShardedIndex

        b = self.bounds
        key = tuple(b)
        if self._views is not None and self._views_key == key:
            for v in self._views:
                v.path = self.path
            return
        self._views = []
        for g, sh in enumerate(self.shards):
            fab = LocalFabric(self._group, g)
            v = ShardedIndex(self.d, b[-1], device=self.devices[g], fabric=fab, local=sh._shard, bounds=b, exchange="p2p")
            v.path = self.path
            self._views.append(v)
        self._views_key = key
        self._out = {}

    def _host_out(self, rank: int, view, shape):
        key = (rank, self._turn, shape)
        o = self._out.get(key)
        if o is None:
            for k2 in [k2 for k2 in self._out if k2[0] == rank and k2[1] == self._turn]:
                del self._out[k2]
            o = self._out[key] = view.host_output(*shape)
        return o

    def search(self, x, k: int):
        """``index.search(x, k)`` = the alpha = 0 case of the fused step (P rows pass through unchanged)."""
        x = _as_f32_2d(x.detach().cpu().numpy() if _is_torch(x) else x, self.d, "search(x)")
        D, I = self.search_mixed(x, x, [0.0], k)
        return D[0], I[0]

    def search_mixed(self, P, S, alphas, k: int):
        P = _as_f32_2d(P.detach().cpu() if _is_torch(P) else P, self.d, "P")
        S = _as_f32_2d(S.detach().cpu() if _is_torch(S) else S, self.d, "S")
        assert tuple(P.shape) == tuple(S.shape)
        shape = (len(alphas), int(P.shape[0]), int(k))
        if shape[0] == 0 or shape[1] == 0:
            return np.empty(shape, np.float32), np.empty(shape, np.int64)
        self._ensure_views()

        def step(rank, view):
            out = self._host_out(rank, view, shape)
            return view.search_mixed_host(P, S, alphas, k, out=out)

        D, I = self._fanout(step)[0]
        self._turn ^= 1
        self.two_phase_used = any(v.two_phase_used for v in self._views)
        return D.numpy(), I.numpy()


def index_cpu_to_gpus_list(index, co=None, gpus: Optional[Sequence[int]] = None, ngpu: int = -1):
    """Row-shard a CPU index over several GPUs (one process)."""
    if gpus is None:
        n = get_num_gpus() if ngpu is None or ngpu < 0 else ngpu
        gpus = list(range(n))
    if len(gpus) < 1:
        raise RuntimeError("index_cpu_to_gpus_list: no GPU")
    if isinstance(index, IndexIDMap):
        out = IndexIDMap(IndexShardsIP(index.d, gpus))
        out.index = index_cpu_to_gpus_list(index.index, co, gpus)
        out._ids = [index.id_map.copy()]
        return out
    if isinstance(index, GpuIndexFlatIP):
        index = index_gpu_to_cpu(index)
    if not isinstance(index, IndexFlatIP):
        raise RuntimeError(f"index_cpu_to_gpus_list: unsupported index type {type(index).__name__}")
    sh = IndexShardsIP(index.d, gpus)
    if index.ntotal:
        sh.add_sharded(index._rows())
    return sh


def index_cpu_to_all_gpus(index, co=None, ngpu: int = -1):
    return index_cpu_to_gpus_list(index, co=co, gpus=None, ngpu=ngpu)


def write_index(index, path) -> None:
    from .io import write_index as _w

    _w(index, path)


def read_index(path, io_flags: int = 0):
    from .io import read_index as _r

    return _r(path)


def read_index_to_gpu(path, device: int = 0):
    """Extension (not in faiss): read_index + index_cpu_to_gpu without the host-resident copy --
    the file is validated, then streamed chunk by chunk into the device index."""
    from .io import read_index_to_gpu as _r

    return _r(path, device)
