"""Corpus row-sharding across GPUs.

New capability named by BASELINE north_star (4); the reference never shards an
index (it only calls index_cpu_to_gpu(res, one_gpu_id, index):
onepass_dense_mix_run_custom_lang.py:661-663).

Shard r owns the contiguous rows [bounds[r], bounds[r+1]) of the global corpus
(row i -> shard floor(i*G/N)).  Queries are replicated.  ``ShardedIndex`` is written
SPMD: the same code runs once per shard, either in one process per GPU
(``torch.distributed`` ranks, peer memory through torch symmetric memory: ``SymmFabric``)
or in one thread per GPU of a single process (``LocalFabric``, used by
``cmx.faiss.IndexShardsIP``).  A fabric provides peer-mapped buffers, a stream-ordered
cross-shard barrier and a host buffer every shard's kernels can write.

One search step (rescore precision), everything enqueued without a host round trip:
  0. every shard publishes its two error-bound numbers; with HOST query vectors each
     shard uploads 1/G of P,S and replicates that slice into the peers over NVLink
     (``cmx_peer_broadcast``)                                              [barrier]
  1. ``cmx_search_begin``: fused mix + approximate pass over the shard; the shard's k
     best APPROXIMATE scores per query and its status word land in peer-mapped buffers
                                                                           [barrier]
  2. ``cmx_union_kth``: each shard finds the GLOBAL k-th best approximate score of ITS
     slice of the queries from all shards' lists (peer reads) and stores it into every
     shard (peer stores); ORs all status words                             [barrier]
  3. ``cmx_search_end``: exact rescoring of the rows that can still reach the global
     top-k (1/G of the single-GPU rescoring work per shard)                [barrier]
  4. ``cmx_merge_topk_peers``: each shard merges its slice of the queries from all
     shards' lists in place (peer reads) and stores the merged rows into every shard's
     output buffer -- or, for host results, straight into ONE pinned host buffer all
     shards have mapped, so the 84 MB of (D, I) cross G PCIe links in parallel
                                                                           [barrier]
  then the status word is read once: if ANY shard overflowed a buffer or cannot run the
  one-pass arithmetic, ALL shards (they OR the same words) redo the step with the plain
  per-shard search + merge.  Batches of more than 8192 queries (11 alphas x 6980) run
  steps 1-3 chunk by chunk.
Merge order is (score desc, shard, position) = (score desc, global row asc), so the
G-way result equals the 1-GPU result exactly, ties included.

``exchange="allgather"`` keeps the NCCL baseline (per-shard full search, all_gather,
merge kernel on every rank).  The local engine and the merge function are injectable so
that the host-side logic (partition, id bases, exchange layout) is testable on CPU
ranks with gloo; the defaults are the CUDA engine and the CUDA merge kernel -- there is
no CPU fallback.
"""
from __future__ import annotations

import os
import threading
from typing import Callable, List, Optional, Sequence

import numpy as np

try:
    import torch
    import torch.distributed as dist
except Exception:  # pragma: no cover
    torch = None
    dist = None

QUERY_CHUNK = 8192  # queries per two-phase pass (cmx_search_begin's limit)


def shard_bounds(n: int, world: int) -> List[int]:
    """row i -> rank floor(i*world/n): contiguous ranges whose sizes differ by at most 1."""
    return [(n * r) // world for r in range(world + 1)]


class HostResult:
    """(D, I) in page-locked host memory that GPU kernels write directly (``D_dev`` / ``I_dev`` are the
    device-side aliases).  With several processes it is ONE POSIX shared-memory segment mapped by all."""

    def __init__(self, D, I, D_dev: int, I_dev: int, keep=None, unregister: Optional[int] = None):
        self.D, self.I, self.D_dev, self.I_dev = D, I, D_dev, I_dev
        self._keep, self._unregister = keep, unregister

    def __iter__(self):
        return iter((self.D, self.I))

    def __del__(self):
        if self._unregister:
            try:
                from .engine import host_unregister

                host_unregister(self._unregister)
            except Exception:
                pass
            self._unregister = None


def _host_result(buf, nbytes_d: int, shape, alias: int, unregister=None) -> HostResult:
    n = int(np.prod(shape))
    D = buf[: 4 * n].view(torch.float32).view(*shape)
    I = buf[nbytes_d : nbytes_d + 8 * n].view(torch.int64).view(*shape)
    return HostResult(D, I, alias, alias + nbytes_d, keep=buf, unregister=unregister)


# ------------------------------------------------------------------------------------ fabrics
class SymmFabric:
    """One process per GPU (torch.distributed): peer memory and barriers from torch symmetric memory."""

    def __init__(self, group, device: int):
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.device = torch.device("cuda", device)
        self._hdl = None
        self._bar_buf = None
        self._shm_serial = 0

    def alloc(self, dtype, numel: int):
        import torch.distributed._symmetric_memory as symm_mem

        t = symm_mem.empty((max(1, int(numel)),), dtype=dtype, device=self.device)
        hdl = symm_mem.rendezvous(t, self.group)
        return t, [int(p) for p in hdl.buffer_ptrs]

    def barrier(self, channel: int = 0) -> None:
        if self._hdl is None:
            # the barrier's signal pads belong to an allocation: a dedicated one that lives as long as the fabric (the
            # search buffers are re-allocated when the batch shape changes)
            import torch.distributed._symmetric_memory as symm_mem

            self._bar_buf = symm_mem.empty((64,), dtype=torch.float32, device=self.device)
            self._hdl = symm_mem.rendezvous(self._bar_buf, self.group)
        self._hdl.barrier(channel=channel)  # a kernel on the current stream: no host synchronisation

    def all_ok(self, ok: bool) -> bool:
        """collective AND (slow path only)."""
        t = torch.tensor([0.0 if ok else 1.0], device=self.device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
        return float(t.item()) == 0.0

    def host_buffer(self, nbytes: int):
        """nbytes of host memory shared by all ranks (POSIX shm), page-locked and device-mapped in each."""
        from .engine import host_register

        self._shm_serial += 1
        names = [f"/dev/shm/cmx_{os.getpid()}_{self._shm_serial}" if self.rank == 0 else None]
        dist.broadcast_object_list(names, src=dist.get_global_rank(self.group, 0), group=self.group)
        if self.rank == 0:
            with open(names[0], "wb") as fh:
                fh.truncate(nbytes)
        dist.barrier(group=self.group)
        buf = torch.from_file(names[0], shared=True, size=nbytes, dtype=torch.uint8)
        dist.barrier(group=self.group)
        if self.rank == 0:
            os.unlink(names[0])
        alias = host_register(buf.data_ptr(), nbytes)
        return buf, alias, buf.data_ptr()


class LocalGroup:
    """State shared by the G thread-ranks of one process."""

    def __init__(self, devices: Sequence[int]):
        from . import _lib

        self.devices = [int(v) for v in devices]
        self.world = len(self.devices)
        self.tbar = threading.Barrier(self.world)
        self.slots: List = [None] * self.world
        self.shared = {}
        for a in self.devices:
            for b in self.devices:
                if a != b:
                    _lib.check(_lib.lib().cmx_enable_peer_access(a, b))


class LocalFabric:
    """One thread per GPU of a single process: plain device allocations are peer-accessible once peer
    access is enabled; a barrier is a host rendezvous of the threads plus cross-stream event waits."""

    def __init__(self, group: LocalGroup, rank: int):
        self.g, self.rank, self.world = group, int(rank), group.world
        self.device = torch.device("cuda", group.devices[rank])

    def _exchange(self, value):
        self.g.slots[self.rank] = value
        self.g.tbar.wait()
        out = list(self.g.slots)
        self.g.tbar.wait()
        return out

    def alloc(self, dtype, numel: int):
        t = torch.zeros((max(1, int(numel)),), dtype=dtype, device=self.device)
        return t, self._exchange(int(t.data_ptr()))

    def barrier(self, channel: int = 0) -> None:
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        evs = self._exchange(ev)
        st = torch.cuda.current_stream(self.device)
        for g, e in enumerate(evs):
            if g != self.rank:
                st.wait_event(e)

    def all_ok(self, ok: bool) -> bool:
        return all(self._exchange(bool(ok)))

    def host_buffer(self, nbytes: int):
        from .engine import host_register

        if self.rank == 0:
            self.g.shared["host"] = torch.empty((nbytes,), dtype=torch.uint8, pin_memory=True)
        self.g.tbar.wait()
        buf = self.g.shared["host"]
        self.g.tbar.wait()
        return buf, host_register(buf.data_ptr(), nbytes), None


# ------------------------------------------------------------------------------------ index
class ShardedIndex:
    """A flat IP index whose rows are sharded over the ranks of a fabric / process group."""

    def __init__(self, d: int, ntotal_global: int, device: Optional[int] = None, group=None,
                 engine_factory: Optional[Callable] = None, merge_fn: Optional[Callable] = None,
                 exchange: str = "auto", fabric=None, local=None, bounds: Optional[Sequence[int]] = None):
        self.d = int(d)
        self.group = group
        self.fabric = fabric
        if fabric is not None:
            self.world, self.rank = fabric.world, fabric.rank
        else:
            self.world = dist.get_world_size(group) if dist is not None and dist.is_initialized() else 1
            self.rank = dist.get_rank(group) if dist is not None and dist.is_initialized() else 0
        self.bounds = [int(v) for v in bounds] if bounds is not None else shard_bounds(int(ntotal_global), self.world)
        self.ntotal = int(ntotal_global)
        self.row0, self.row1 = self.bounds[self.rank], self.bounds[self.rank + 1]
        self._custom_engine = engine_factory is not None
        if engine_factory is None:
            from .engine import Shard

            engine_factory = lambda dim: Shard(dim, 0 if device is None else device)  # noqa: E731
        if merge_fn is None:
            from .engine import merge_topk

            merge_fn = merge_topk
        self.local = local if local is not None else engine_factory(self.d)  # `local`: an existing cmx.engine.Shard
        self._merge = merge_fn
        self.path = "auto"
        # "p2p": fused peer-memory merge, "allgather": NCCL all_gather + merge, "auto": p2p when possible
        self.exchange = exchange
        self.exchange_used = "none" if self.world == 1 else "allgather"
        self.exchange_error = None
        self._device = device
        # rescore precision + several shards: exchange approximate k-th scores before the exact rescoring
        self.two_phase = True
        self.two_phase_used = False
        self.fallback_steps = 0  # steps redone without the two-phase cut (status word set on some shard)
        self.last_status = 0     # OR of the shards' status words of the last two-phase step (0 = clean)
        self.fallback_chunks = 0  # 8192-query chunks redone because a shard's buffer overflowed / a speculation missed
        self.profile = False
        self.timing = {}
        self._t_last = 0.0
        self.precision = "rescore"
        self._bufs = {}

    def _dev(self):
        return torch.device("cuda", torch.cuda.current_device() if self._device is None else self._device)

    def _tick(self, name: str) -> None:
        import time

        torch.cuda.synchronize(self._dev())
        now = time.perf_counter()
        if name != "start":
            self.timing[name] = self.timing.get(name, 0.0) + (now - self._t_last) * 1e3
        self._t_last = now

    def set_precision(self, mode: str) -> None:
        """'rescore' (default) or 'split' arithmetic of the tensor path (see cmx.h).  Set it on every shard."""
        self.precision = mode
        if hasattr(self.local, "set_precision"):
            self.local.set_precision(mode)

    # ---- storage: each rank adds ITS rows, in global row order ----------------
    def reserve_local(self) -> None:
        if hasattr(self.local, "reserve"):
            self.local.reserve(self.row1 - self.row0)

    def add_local(self, x) -> None:
        self.local.add(x)
        assert self.local.ntotal <= self.row1 - self.row0, "rank added more rows than its shard holds"

    def local_complete(self) -> bool:
        return self.local.ntotal == self.row1 - self.row0

    # ---- NCCL baseline: all_gather + merge kernel ---------------------------------------------
    def _exchange_and_merge(self, D, I, k: int):
        if self.world == 1:
            return D, I
        lead = tuple(D.shape[:-1])
        nq = int(np.prod(lead)) if lead else 1
        D2 = D.reshape(nq, k).contiguous()
        I2 = I.reshape(nq, k).contiguous()
        # output = concatenation along dim 0 (the layout both nccl and gloo accept)
        Dp = torch.empty((self.world * nq, k), dtype=D2.dtype, device=D2.device)
        Ip = torch.empty((self.world * nq, k), dtype=I2.dtype, device=I2.device)
        dist.all_gather_into_tensor(Dp, D2, group=self.group)
        dist.all_gather_into_tensor(Ip, I2, group=self.group)
        Dm, Im = self._merge(Dp.view(self.world, nq, k), Ip.view(self.world, nq, k))
        return Dm.reshape(*lead, k), Im.reshape(*lead, k)

    # ---- peer-memory path -------------------------------------------------------------------------
    def _fabric(self):
        if self.fabric is None:
            self.fabric = SymmFabric(self.group, self._dev().index)
        return self.fabric

    def _p2p_ok(self) -> bool:
        if self.world == 1 or self.exchange == "allgather" or self._custom_engine:
            return False
        return torch is not None and torch.cuda.is_available()

    def _buffers(self, nqt: int, k: int, nq: int = 0):
        """Peer-mapped buffers of one (batch, k) shape; allocation is collective, done once per shape."""
        key = (nqt, k, nq)
        b = self._bufs.get(key)
        if b is not None:
            return b
        fab = self._fabric()
        chunk = min(nqt, QUERY_CHUNK)
        b = {}
        for name, dt, numel in (("D_loc", torch.float32, nqt * k), ("I_loc", torch.int64, nqt * k),
                                ("D_out", torch.float32, nqt * k), ("I_out", torch.int64, nqt * k),
                                ("ascore", torch.float32, chunk * k), ("kth", torch.float32, chunk),
                                ("bounds", torch.float32, 4), ("flags", torch.int32, 4),
                                ("P", torch.float32, nq * self.d), ("S", torch.float32, nq * self.d)):
            if numel > 0:
                b[name] = fab.alloc(dt, numel)
        b["flag_any"] = torch.zeros((1,), dtype=torch.int32, device=fab.device)
        self._bufs = {key: b}  # keep one shape alive
        return b

    def host_output(self, nA: int, nq: int, k: int) -> HostResult:
        """Result buffers in pinned host memory for ``search_mixed_host``: with several shards on a peer-memory
        fabric ONE buffer every shard's merge kernel writes its query slice into (collective call)."""
        n = nA * nq * k
        nbytes_d = (4 * n + 255) // 256 * 256
        if self._p2p_ok():
            buf, alias, unreg = self._fabric().host_buffer(nbytes_d + 8 * n)
            return _host_result(buf, nbytes_d, (nA, nq, k), alias, unreg)
        buf = torch.empty((nbytes_d + 8 * n,), dtype=torch.uint8, pin_memory=torch.cuda.is_available())
        return _host_result(buf, nbytes_d, (nA, nq, k), buf.data_ptr())

    def _two_phase_eligible(self) -> bool:
        # global facts only: every shard must take the same branch (a shard-local condition here would
        # leave the others waiting in a barrier); shard-local trouble travels in the status words instead
        return (self.two_phase and self.world > 1 and not self._custom_engine and self.path in ("auto", "tensor")
                and self.precision == "rescore" and self.ntotal > 0)

    def _step_p2p(self, nA: int, nq: int, k: int, prepare, run_local, host_out: Optional[HostResult], upload=None):
        """One search step over peer memory (module docstring).  ``prepare()`` enqueues the prologue and returns
        the device address of the [nA*nq, d] queries; ``run_local(out)`` is the plain per-shard search."""
        from .engine import merge_topk_peers, union_kth

        fab = self._fabric()
        G, r = self.world, self.rank
        nqt = nA * nq
        b = self._buffers(nqt, k, nq if upload is not None else 0)
        dev = fab.device.index
        tm = self._tick if self.profile else (lambda name: None)
        tm("start")
        D_loc, D_ptrs = b["D_loc"]
        I_loc, I_ptrs = b["I_loc"]
        if host_out is not None:
            outs_D, outs_I = [host_out.D_dev], [host_out.I_dev]
            result = (host_out.D, host_out.I)
        else:
            outs_D, outs_I = b["D_out"][1], b["I_out"][1]
            result = (b["D_out"][0][: nqt * k].view(nA, nq, k), b["I_out"][0][: nqt * k].view(nA, nq, k))
        done = False
        if self._two_phase_eligible():
            _, bounds_ptrs = b["bounds"]
            _, flag_ptrs = b["flags"]
            flag_any = b["flag_any"]
            self.local.export_bounds(bounds_ptrs[r])
            if upload is not None:
                upload(b)
            nchunks = (nqt + QUERY_CHUNK - 1) // QUERY_CHUNK
            if flag_any.numel() < nchunks:
                flag_any = b["flag_any"] = torch.zeros((nchunks,), dtype=torch.int32, device=fab.device)
            flag_any.zero_()
            fab.barrier(0)  # error bounds (and uploaded query slices) of every shard are visible
            q_ptr = prepare(b)
            _, as_ptrs = b["ascore"]
            _, kth_ptrs = b["kth"]
            tm("prologue")
            chunks = [(c0, min(QUERY_CHUNK, nqt - c0)) for c0 in range(0, nqt, QUERY_CHUNK)]
            for ci, (c0, nqc) in enumerate(chunks):
                self.local.search_begin(q_ptr + 4 * self.d * c0, nqc, k, self.row0, bounds_ptrs, 1.0 / G, as_ptrs[r], flag_ptrs[r])
                tm("begin")
                fab.barrier(1)  # every shard's approximate top-k scores and status word are visible
                union_kth(as_ptrs, nqc, k, (nqc * r) // G, (nqc * (r + 1)) // G, kth_ptrs, dev, flag_ptrs, flag_any.data_ptr() + 4 * ci)
                fab.barrier(0)  # the global k-th scores of all queries have landed
                tm("union_kth")
                self.local.search_end([kth_ptrs[r]], D_ptrs[r] + 4 * k * c0, I_ptrs[r] + 8 * k * c0)
                tm("rescore")
            fab.barrier(1)  # every shard's exact lists are complete and visible
            merge_topk_peers(D_ptrs, I_ptrs, nqt, k, (nqt * r) // G, (nqt * (r + 1)) // G, outs_D, outs_I, dev)
            fab.barrier(0)  # every shard's slice has landed in every output
            flags = flag_any[:nchunks].tolist()  # the step's only host synchronisation
            self.last_status = 0
            for f in flags:
                self.last_status |= int(f)
            tm("merge")
            bad = [ci for ci, f in enumerate(flags) if f]
            if any(int(flags[ci]) & 0x100 for ci in bad):
                bad = None  # a shard cannot run the one-pass arithmetic at all: the whole step goes the plain way
            if bad is not None:
                done = True
                self.two_phase_used = True
                if bad:
                    # Some shard overflowed a candidate buffer / missed a speculation in these chunks (every shard sees the
                    # same words): only they are redone -- plain per-shard search of the chunk (its own reruns) into the
                    # same lists, then the merge of the chunk's queries.
                    self.fallback_chunks += len(bad)
                    err = None
                    try:
                        for ci in bad:
                            c0, nqc = chunks[ci]
                            self.local.search_ptr(q_ptr + 4 * self.d * c0, nqc, k, self.row0, D_ptrs[r] + 4 * k * c0,
                                                  I_ptrs[r] + 8 * k * c0, self.path)
                    except Exception as exc:  # noqa: BLE001
                        err = exc
                    if not fab.all_ok(err is None):
                        raise RuntimeError(f"sharded search failed on a shard (this shard: {err!r})")
                    fab.barrier(1)
                    for ci in bad:
                        c0, nqc = chunks[ci]
                        merge_topk_peers(D_ptrs, I_ptrs, nqt, k, c0 + (nqc * r) // G, c0 + (nqc * (r + 1)) // G, outs_D, outs_I, dev)
                    fab.barrier(0)
                    torch.cuda.current_stream(fab.device).synchronize()
            else:
                self.fallback_steps += 1
        if not done:
            # plain per-shard search (each shard rescoring its own band; its own reruns) + fused merge;
            # a shard that fails here makes all of them raise together
            err = None
            try:
                run_local((D_loc[: nqt * k].view(nA, nq, k), I_loc[: nqt * k].view(nA, nq, k)))
            except Exception as exc:  # noqa: BLE001
                err = exc
            if not fab.all_ok(err is None):
                raise RuntimeError(f"sharded search failed on a shard (this shard: {err!r})")
            fab.barrier(1)
            merge_topk_peers(D_ptrs, I_ptrs, nqt, k, (nqt * r) // G, (nqt * (r + 1)) // G, outs_D, outs_I, dev)
            fab.barrier(0)
            torch.cuda.current_stream(fab.device).synchronize()
        self.exchange_used = "p2p"
        return result

    def _run(self, nA, nq, k, prepare, run_local, host_out=None, upload=None):
        if self._p2p_ok() and nA * nq > 0:
            try:
                return self._step_p2p(nA, nq, k, prepare, run_local, host_out, upload)
            except Exception as exc:  # symmetric memory unavailable: keep the NCCL exchange
                if self.exchange == "p2p" or self.exchange_used == "p2p":
                    raise
                self.exchange = "allgather"
                self.exchange_error = repr(exc)
        D, I = run_local(None)
        return self._exchange_and_merge(D, I, k)

    def _to_device(self, x):
        """numpy / host tensors -> CUDA (the NCCL and peer paths work on device memory)."""
        if self._custom_engine or self.world == 1:
            return x
        if not (torch is not None and isinstance(x, torch.Tensor)):
            x = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32))
        return x.to(self._dev(), non_blocking=True) if not x.is_cuda else x

    def search(self, x, k: int):
        k = int(k)
        x = self._to_device(x)
        nq = int(x.shape[0]) if hasattr(x, "shape") and len(x.shape) == 2 else 1

        def run_local(out):
            kw = {"out": (out[0][0], out[1][0])} if out else {}
            return self.local.search(x, k, id_base=self.row0, path=self.path, **kw)

        def prepare(b):
            self._keep = x.contiguous()  # alive until the step's final synchronisation
            return int(self._keep.data_ptr())

        D, I = self._run(1, nq, k, prepare, run_local)
        return (D[0], I[0]) if D.dim() == 3 else (D, I)

    def search_mixed(self, P, S, alphas: Sequence[float], k: int):
        k = int(k)
        P, S = self._to_device(P), self._to_device(S)
        nA, nq = len(alphas), int(P.shape[0])

        def run_local(out):
            return self.local.search_mixed(P, S, alphas, k, id_base=self.row0, path=self.path, **({"out": out} if out else {}))

        return self._run(nA, nq, k, lambda b: self.local.search_prepare(P, S, alphas), run_local)

    def search_mixed_host(self, P_h, S_h, alphas: Sequence[float], k: int, out: Optional[HostResult] = None):
        """The end-to-end call: P_h, S_h in (pinned) HOST memory on every shard, result in host memory.
        One shard: ``cmx_search_mixed`` with host buffers (the final kernels write the pinned output).
        Several shards on a peer-memory fabric: each uploads 1/G of P,S and replicates it over NVLink;
        each merge kernel writes its slice of (D, I) into the shared pinned buffer ``out``."""
        k = int(k)
        nA, nq = len(alphas), int(P_h.shape[0])
        if nA * nq == 0:
            return torch.empty((nA, nq, k), dtype=torch.float32), torch.empty((nA, nq, k), dtype=torch.int64)
        if out is None:
            out = self.host_output(nA, nq, k)
        if self.world == 1:
            self.local.search_mixed(P_h, S_h, alphas, k, id_base=self.row0, path=self.path, out=(out.D, out.I))
            return out.D, out.I
        if not self._p2p_ok() or (self.d & 3) != 0:
            D, I = self.search_mixed(P_h, S_h, alphas, k)
            out.D.copy_(D, non_blocking=True)
            out.I.copy_(I, non_blocking=True)
            torch.cuda.current_stream(self._dev()).synchronize()
            return out.D, out.I
        from .engine import peer_broadcast

        G, r, d = self.world, self.rank, self.d
        a0, a1 = (nq * r) // G, (nq * (r + 1)) // G
        Ph = P_h if isinstance(P_h, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(P_h, dtype=np.float32))
        Sh = S_h if isinstance(S_h, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(S_h, dtype=np.float32))
        state = {}

        def upload(b):
            # this shard's slice of the query vectors: one H2D, then NVLink stores into every peer
            for name, src in (("P", Ph), ("S", Sh)):
                t, ptrs = b[name]
                if a1 > a0:
                    t[a0 * d : a1 * d].view(a1 - a0, d).copy_(src[a0:a1], non_blocking=True)
                    peer_broadcast(ptrs[r] + 4 * d * a0, [p + 4 * d * a0 for g, p in enumerate(ptrs) if g != r],
                                   4 * d * (a1 - a0), self._dev().index)
                state[name] = t[: nq * d].view(nq, d)

        def prepare(b):
            return self.local.search_prepare(state["P"], state["S"], alphas)

        def run_local(o):
            return self.local.search_mixed(state["P"], state["S"], alphas, k, id_base=self.row0, path=self.path,
                                           **({"out": o} if o else {}))

        if not self._two_phase_eligible():
            # the step goes straight to run_local: the upload (and its barrier) happen here
            upload(self._buffers(nA * nq, k, nq))
            self._fabric().barrier(0)
            return self._step_p2p(nA, nq, k, prepare, run_local, out, upload=lambda b: None)
        return self._step_p2p(nA, nq, k, prepare, run_local, out, upload=upload)
