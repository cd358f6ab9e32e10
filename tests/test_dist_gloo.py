"""world_size-2 gloo test (CPU) of the sharding host logic: partition, id bases, the
all_gather layout and the merge order.  The local engine and the merge are injected
oracle stand-ins (test infrastructure); the product defaults are the CUDA ones."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle


class OracleEngine:
    def __init__(self, d):
        self.d, self.rows = d, []

    @property
    def ntotal(self):
        return sum(r.shape[0] for r in self.rows)

    def add(self, x):
        self.rows.append(np.asarray(x, dtype=np.float32))

    def _X(self):
        return np.concatenate(self.rows) if self.rows else np.zeros((0, self.d), np.float32)

    def search(self, x, k, id_base=0, path="auto"):
        D, I = oracle.flat_ip_search(self._X(), np.asarray(x), k)
        I = np.where(I >= 0, I + id_base, -1)
        return torch.from_numpy(D), torch.from_numpy(I)

    def search_mixed(self, P, S, alphas, k, id_base=0, path="auto"):
        Q, _ = oracle.mix_normalize(np.asarray(P), np.asarray(S), alphas)
        outs = [self.search(Q[a], k, id_base) for a in range(len(alphas))]
        return torch.stack([o[0] for o in outs]), torch.stack([o[1] for o in outs])


def _oracle_merge(Dp, Ip):
    D, I = oracle.merge_topk(Dp.numpy(), Ip.numpy(), Dp.shape[-1])
    return torch.from_numpy(D), torch.from_numpy(I)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from cmx.dist import ShardedIndex, shard_bounds

        rng = np.random.default_rng(5)
        N, d, nq, k = 1001, 32, 9, 20
        X = rng.standard_normal((N, d)).astype(np.float32)
        X /= np.linalg.norm(X, axis=1, keepdims=True)
        X[700:720] = X[10:30]  # ties across the shard boundary
        P = X[rng.integers(0, N, nq)] + 0.1 * rng.standard_normal((nq, d)).astype(np.float32)
        S = X[rng.integers(0, N, nq)] + 0.1 * rng.standard_normal((nq, d)).astype(np.float32)
        idx = ShardedIndex(d, N, engine_factory=OracleEngine, merge_fn=_oracle_merge)
        assert idx.bounds == shard_bounds(N, world) == [0, 500, 1001]
        idx.add_local(X[idx.row0 : idx.row1])
        assert idx.local_complete()
        D, I = idx.search(P, k)
        Dm, Im = idx.search_mixed(P, S, [0.0, 0.5], k)
        np.savez(os.path.join(out_dir, f"r{rank}.npz"), D=D.numpy(), I=I.numpy(), Dm=Dm.numpy(), Im=Im.numpy(),
                 X=X, P=P, S=S)
    finally:
        dist.destroy_process_group()


def test_two_rank_sharded_search_equals_single(tmp_path):
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    r0, r1 = np.load(tmp_path / "r0.npz"), np.load(tmp_path / "r1.npz")
    for key in ("D", "I", "Dm", "Im"):
        assert np.array_equal(r0[key], r1[key])  # every rank holds the merged result
    X, P, S = r0["X"], r0["P"], r0["S"]
    D, I = oracle.flat_ip_search(X, P, 20)
    assert np.array_equal(r0["I"], I) and np.array_equal(r0["D"], D)
    Q, _ = oracle.mix_normalize(P, S, [0.0, 0.5])
    for a in range(2):
        Da, Ia = oracle.flat_ip_search(X, Q[a], 20)
        assert np.array_equal(r0["Im"][a], Ia) and np.array_equal(r0["Dm"][a], Da)


def test_shard_bounds_cover_everything():
    from cmx.dist import shard_bounds

    for n in (0, 1, 7, 8841823, 17683646):
        for w in (1, 2, 4, 8):
            b = shard_bounds(n, w)
            assert b[0] == 0 and b[-1] == n and all(x <= y for x, y in zip(b, b[1:]))
            sizes = [y - x for x, y in zip(b, b[1:])]
            assert max(sizes) - min(sizes) <= 1
