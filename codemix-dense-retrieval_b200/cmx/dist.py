"""Corpus row-sharding across ranks (one process per GPU, torch.distributed).

New capability named by BASELINE north_star (4); the reference never shards an
index (it only calls index_cpu_to_gpu(res, one_gpu_id, index):
onepass_dense_mix_run_custom_lang.py:661-663).

Rank r owns the contiguous rows [bounds[r], bounds[r+1]) of the global corpus
(row i -> rank floor(i*G/N)).  Queries are replicated.  A search is
  1. local fused mix+search on the rank's shard with id_base = bounds[r]
     (global row numbers come straight out of the kernel),
  2. ONE exchange: all_gather of the per-shard (D, I) lists (nq*k*12 B per rank),
  3. the k-way merge kernel on every rank (score desc, shard, position) -- so the
     G-way result equals the 1-GPU result exactly, ties included.
Steps 2+3 are fused when the ranks can map each other's memory (NVLink / NVSwitch,
torch symmetric memory): every rank writes its lists into a symmetric buffer, and ONE
kernel per rank (`cmx_merge_topk_peers`) reads all peers' lists in place through peer
pointers for ITS slice of the queries, merges them and stores the merged rows into
every rank's output buffer -- no all_gather, no staging copy, and the merge work is
split G ways.  `exchange="allgather"` keeps the NCCL version (the baseline).
The local engine and the merge function are injectable so that the host-side
logic (partition, id bases, exchange layout) is testable on CPU ranks with gloo;
the defaults are the CUDA engine and the CUDA merge kernel -- there is no CPU
fallback.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence

import numpy as np

try:
    import torch
    import torch.distributed as dist
except Exception:  # pragma: no cover
    torch = None
    dist = None


def shard_bounds(n: int, world: int) -> List[int]:
    """row i -> rank floor(i*world/n): contiguous ranges whose sizes differ by at most 1."""
    return [(n * r) // world for r in range(world + 1)]


class ShardedIndex:
    """A flat IP index whose rows are sharded over the ranks of a process group."""

    def __init__(self, d: int, ntotal_global: int, device: Optional[int] = None, group=None,
                 engine_factory: Optional[Callable] = None, merge_fn: Optional[Callable] = None,
                 exchange: str = "auto"):
        self.d = int(d)
        self.group = group
        self.world = dist.get_world_size(group) if dist is not None and dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist is not None and dist.is_initialized() else 0
        self.bounds = shard_bounds(int(ntotal_global), self.world)
        self.ntotal = int(ntotal_global)
        self.row0, self.row1 = self.bounds[self.rank], self.bounds[self.rank + 1]
        self._default_engine = engine_factory is None
        if engine_factory is None:
            from .engine import Shard

            engine_factory = lambda dim: Shard(dim, 0 if device is None else device)  # noqa: E731
        if merge_fn is None:
            from .engine import merge_topk

            merge_fn = merge_topk
        self.local = engine_factory(self.d)
        self._merge = merge_fn
        self.path = "auto"
        # "p2p": fused peer-memory merge, "allgather": NCCL all_gather + merge, "auto": p2p when possible
        self.exchange = exchange
        self.exchange_used = "none" if self.world == 1 else "allgather"
        self._peer = {}
        self._device = device
        # rescore precision + several shards: exchange approximate k-th scores before the exact rescoring
        self.two_phase = True
        self.two_phase_used = False
        self.profile = False
        self.timing = {}
        self._t_last = 0.0
        self.precision = "rescore"
        self._flag = None
        if self.world > 1 and torch is not None and torch.cuda.is_available() and self._default_engine:
            self._flag = torch.zeros((1,), dtype=torch.float32,
                                     device=torch.device("cuda", torch.cuda.current_device() if device is None else device))
        self._custom_engine = engine_factory is not None and not hasattr(self.local, "_h")

    def _tick(self, name: str) -> None:
        import time

        torch.cuda.synchronize()
        now = time.perf_counter()
        if name != "start":
            self.timing[name] = self.timing.get(name, 0.0) + (now - self._t_last) * 1e3
        self._t_last = now

    def set_precision(self, mode: str) -> None:
        """'rescore' (default) or 'split' arithmetic of the tensor path (see cmx.h)."""
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

    # ---- search ------------------------------------------------------------------
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

    # ---- fused exchange + merge over peer memory -------------------------------------
    def _peer_buffers(self, nq: int, k: int):
        """Symmetric (peer-mappable) buffers for the per-shard lists and the merged result."""
        key = (nq, k)
        if key in self._peer:
            return self._peer[key]
        import torch.distributed._symmetric_memory as symm_mem

        dev = torch.device("cuda", torch.cuda.current_device() if self._device is None else self._device)
        grp = self.group if self.group is not None else dist.group.WORLD
        bufs = {}
        for name, dt, numel in (("D_loc", torch.float32, nq * k), ("I_loc", torch.int64, nq * k),
                                ("D_out", torch.float32, nq * k), ("I_out", torch.int64, nq * k),
                                ("ascore", torch.float32, nq * k), ("kth", torch.float32, nq)):
            t = symm_mem.empty((numel,), dtype=dt, device=dev)
            hdl = symm_mem.rendezvous(t, grp)
            bufs[name] = (t, hdl, [int(p) for p in hdl.buffer_ptrs])
        self._peer = {key: bufs}  # keep one shape alive
        return bufs

    def _p2p_ok(self, like) -> bool:
        if self.world == 1 or self.exchange == "allgather" or self._custom_engine:
            return False
        if not (torch is not None and isinstance(like, torch.Tensor) and like.is_cuda):
            return False
        return True

    def _search_p2p(self, lead_shape, k: int, run_local, two_phase=None):
        import ctypes as C

        from . import _lib

        nq = int(np.prod(lead_shape))
        bufs = self._peer_buffers(nq, k)
        D_loc, hdl, D_ptrs = bufs["D_loc"]
        I_loc, _, I_ptrs = bufs["I_loc"]
        D_out, _, Do_ptrs = bufs["D_out"]
        I_out, _, Io_ptrs = bufs["I_out"]
        done = False
        tm = self._tick if self.profile else (lambda name: None)
        tm("start")
        if two_phase is not None:
            # rescore precision: exchange the shards' k-th best APPROXIMATE scores first, so that each
            # shard rescoring only touches rows that can still reach the GLOBAL top-k
            from .engine import union_kth

            ascore, _, as_ptrs = bufs["ascore"]
            kth, _, kth_ptrs = bufs["kth"]
            flag = self._flag
            flag.fill_(1.0 if two_phase(ascore) else 0.0)
            tm("begin")
            dist.all_reduce(flag, op=dist.ReduceOp.MAX, group=self.group)  # collective fallback decision
            hdl.barrier(channel=0)  # every rank's approximate top-k scores are complete and visible
            ok = float(flag.item()) == 0.0
            tm("flag+barrier")
            if ok:
                # global k-th best approximate score of MY slice of the queries, from all shards' lists
                # (peer reads), written into every shard's kth array (peer stores)
                G = self.world
                union_kth(as_ptrs, nq, k, (nq * self.rank) // G, (nq * (self.rank + 1)) // G, kth_ptrs, kth.device.index)
                hdl.barrier(channel=1)
                tm("union_kth")
                self.local.search_end([kth_ptrs[self.rank]], D_loc, I_loc)
                done = True
                self.two_phase_used = True
                tm("rescore")
        if not done:
            run_local((D_loc.view(*lead_shape, k), I_loc.view(*lead_shape, k)))
        hdl.barrier(channel=0)  # every rank's lists are complete and visible
        G = self.world
        q0, q1 = (nq * self.rank) // G, (nq * (self.rank + 1)) // G
        arr = lambda ptrs: (C.c_void_p * G)(*ptrs)  # noqa: E731
        stream = int(torch.cuda.current_stream().cuda_stream)
        _lib.check(_lib.lib().cmx_merge_topk_peers(arr(D_ptrs), arr(I_ptrs), G, nq, k, q0, q1, arr(Do_ptrs), arr(Io_ptrs), G,
                                                   D_out.device.index, stream))
        hdl.barrier(channel=1)  # every rank's slice has landed in every output buffer
        tm("merge")
        self.exchange_used = "p2p"
        return D_out.view(*lead_shape, k), I_out.view(*lead_shape, k)

    def _run(self, like, lead_shape, k: int, run_local, two_phase=None):
        if self._p2p_ok(like):
            try:
                return self._search_p2p(lead_shape, k, run_local, two_phase)
            except Exception as exc:  # symmetric memory unavailable: keep the NCCL exchange
                if self.exchange == "p2p":
                    raise
                self.exchange = "allgather"
                self.exchange_error = repr(exc)
        if two_phase is not None and self.world > 1 and not self._custom_engine and torch is not None \
                and isinstance(like, torch.Tensor) and like.is_cuda:
            from .engine import union_kth

            nq = int(np.prod(lead_shape))
            ascore = torch.empty((nq * k,), dtype=torch.float32, device=like.device)
            flag = self._flag
            flag.fill_(1.0 if two_phase(ascore) else 0.0)
            dist.all_reduce(flag, op=dist.ReduceOp.MAX, group=self.group)
            if float(flag.item()) == 0.0:
                as_all = torch.empty((self.world * nq * k,), dtype=torch.float32, device=like.device)
                dist.all_gather_into_tensor(as_all, ascore, group=self.group)
                kth = torch.empty((nq,), dtype=torch.float32, device=like.device)
                union_kth([as_all.data_ptr() + 4 * nq * k * g for g in range(self.world)], nq, k, 0, nq, [kth.data_ptr()],
                          like.device.index)
                D = torch.empty((*lead_shape, k), dtype=torch.float32, device=like.device)
                I = torch.empty((*lead_shape, k), dtype=torch.int64, device=like.device)
                self.local.search_end([kth.data_ptr()], D, I)
                self.two_phase_used = True
                return self._exchange_and_merge(D, I, k)
        D, I = run_local(None)
        return self._exchange_and_merge(D, I, k)

    def _sync_error_bounds(self) -> None:
        """Two-phase search cuts every shard's candidates at (global k-th approximate score - margin):
        the margin must come from the corpus maxima over ALL shards (include/cmx.h)."""
        b = torch.tensor(self.local.error_bounds(), dtype=torch.float32, device=self._flag.device)
        dist.all_reduce(b, op=dist.ReduceOp.MAX, group=self.group)
        nmax, rmax = (float(v) for v in b.tolist())
        self.local.raise_error_bounds(nmax, rmax)

    def search(self, x, k: int):
        k = int(k)
        nq = int(x.shape[0]) if hasattr(x, "shape") and len(x.shape) == 2 else 1
        return self._run(x, (nq,), k, lambda out: self.local.search(x, k, id_base=self.row0, path=self.path, **({"out": out} if out else {})))

    def search_mixed(self, P, S, alphas: Sequence[float], k: int):
        k = int(k)
        lead = (len(alphas), int(P.shape[0]))
        two_phase = None
        if (self.two_phase and self.world > 1 and not self._custom_engine and lead[0] * lead[1] <= 8192
                and self.path in ("auto", "tensor") and self.precision == "rescore" and self.local.ntotal > 0):
            self._sync_error_bounds()
            two_phase = lambda kth: self.local.search_mixed_begin(P, S, alphas, k, self.row0, kth)  # noqa: E731
        return self._run(P, lead, k, lambda out: self.local.search_mixed(P, S, alphas, k, id_base=self.row0, path=self.path,
                                                                         **({"out": out} if out else {})), two_phase)
