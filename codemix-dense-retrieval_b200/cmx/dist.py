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
                 engine_factory: Optional[Callable] = None, merge_fn: Optional[Callable] = None):
        self.d = int(d)
        self.group = group
        self.world = dist.get_world_size(group) if dist is not None and dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist is not None and dist.is_initialized() else 0
        self.bounds = shard_bounds(int(ntotal_global), self.world)
        self.ntotal = int(ntotal_global)
        self.row0, self.row1 = self.bounds[self.rank], self.bounds[self.rank + 1]
        if engine_factory is None:
            from .engine import Shard

            engine_factory = lambda dim: Shard(dim, 0 if device is None else device)  # noqa: E731
        if merge_fn is None:
            from .engine import merge_topk

            merge_fn = merge_topk
        self.local = engine_factory(self.d)
        self._merge = merge_fn
        self.path = "auto"

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

    def search(self, x, k: int):
        D, I = self.local.search(x, k, id_base=self.row0, path=self.path)
        return self._exchange_and_merge(D, I, int(k))

    def search_mixed(self, P, S, alphas: Sequence[float], k: int):
        D, I = self.local.search_mixed(P, S, alphas, k, id_base=self.row0, path=self.path)
        return self._exchange_and_merge(D, I, int(k))
