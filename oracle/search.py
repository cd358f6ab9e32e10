"""Oracle for ``IndexIDMap(IndexFlatIP).search`` (test infrastructure only).

Call sites restated: ``index.search(q_chunk, 100)``
(onepass_dense_mix_run_custom_lang.py:878), ``index_search.search(q_chunk,
args.topk)`` (onepass_bilingual_mix_hub_custom_lang.py:950) and the nq=1 calls of
onepass_dense_run.py:427,460.

The arithmetic lives in faiss-gpu 1.8.0 (un-vendored, not installable here --
"parity unpinned" for this part, see ``oracle/__init__``).  FAISS's published
CPU algorithm for ``IndexFlatIP`` is restated: exact fp32 inner products of
every (query, row) pair computed with a blocked SGEMM, the k largest per query
returned sorted by score descending, ids translated through ``id_map``, and
when fewer than k rows exist the tail is padded with ``I = -1`` and
``D = -FLT_MAX``.  FAISS leaves the order of equal scores unspecified; this
oracle (and the CUDA path) break ties by ascending row number, which is one of
the orders FAISS may legally return.
"""

from __future__ import annotations

import numpy as np

NEG_PAD = np.finfo(np.float32).min  # FAISS pads IP results with lowest float


def _stable_topk_desc(scores: np.ndarray, k: int) -> tuple[np.ndarray, np.ndarray]:
    """Top-k per row by (score desc, column asc). NaN scores are never selected."""
    s = np.where(np.isnan(scores), -np.inf, scores)
    order = np.argsort(-s, axis=1, kind="stable")[:, :k]
    vals = np.take_along_axis(s, order, axis=1)
    return vals, order


def flat_ip_search(
    X: np.ndarray,
    Q: np.ndarray,
    k: int,
    ids: np.ndarray | None = None,
    block: int = 65536,
    qblock: int = 4096,
    fast: bool = False,
    dtype=np.float32,
) -> tuple[np.ndarray, np.ndarray]:
    """Brute-force inner-product top-k with FAISS IndexFlatIP result semantics.

    X [N, d], Q [nq, d] -> D [nq, k] float32 (descending), I [nq, k] int64.
    ``fast=True`` uses torch.topk (ties in arbitrary order; for timing), the
    default uses stable sorts so ties come out in ascending row order.
    """
    X = np.ascontiguousarray(X, dtype=dtype)
    Q = np.ascontiguousarray(Q, dtype=dtype)
    assert X.ndim == 2 and Q.ndim == 2 and X.shape[1] == Q.shape[1]
    N = X.shape[0]
    nq = Q.shape[0]
    D = np.full((nq, k), NEG_PAD, dtype=np.float32)
    I = np.full((nq, k), -1, dtype=np.int64)
    if N == 0 or nq == 0:
        return D, I
    if fast:
        return _flat_ip_search_fast(X, Q, k, ids, block, qblock)
    for q0 in range(0, nq, qblock):
        q1 = min(nq, q0 + qblock)
        run_v = np.empty((q1 - q0, 0), dtype=dtype)
        run_i = np.empty((q1 - q0, 0), dtype=np.int64)
        for r0 in range(0, N, block):
            r1 = min(N, r0 + block)
            with np.errstate(all="ignore"):
                sc = Q[q0:q1] @ X[r0:r1].T
            bv, bi = _stable_topk_desc(sc, min(k, r1 - r0))
            cat_v = np.concatenate([run_v, bv], axis=1)
            cat_i = np.concatenate([run_i, bi.astype(np.int64) + r0], axis=1)
            # earlier blocks (smaller rows) come first, stable sort keeps row-asc on ties
            order = np.argsort(-cat_v, axis=1, kind="stable")[:, :k]
            run_v = np.take_along_axis(cat_v, order, axis=1)
            run_i = np.take_along_axis(cat_i, order, axis=1)
        # FAISS's heap threshold starts at lowest-float: -inf / NaN never enter
        valid = run_v > NEG_PAD
        kk = run_v.shape[1]
        D[q0:q1, :kk] = np.where(valid, run_v, NEG_PAD).astype(np.float32)
        I[q0:q1, :kk] = np.where(valid, run_i, -1)
    if ids is not None:
        ids = np.asarray(ids, dtype=np.int64)
        I = np.where(I >= 0, ids[np.clip(I, 0, None)], -1)
    return D, I


def _flat_ip_search_fast(X, Q, k, ids, block, qblock):
    """Blocked torch.mm (MKL SGEMM, all host threads) + running torch.topk merge."""
    import torch

    Xt = torch.from_numpy(X)
    Qt = torch.from_numpy(Q)
    N, nq = X.shape[0], Q.shape[0]
    D = torch.full((nq, k), float(NEG_PAD), dtype=torch.float32)
    I = torch.full((nq, k), -1, dtype=torch.int64)
    for q0 in range(0, nq, qblock):
        q1 = min(nq, q0 + qblock)
        run_v = None
        run_i = None
        for r0 in range(0, N, block):
            r1 = min(N, r0 + block)
            sc = torch.mm(Qt[q0:q1], Xt[r0:r1].T)
            bv, bi = torch.topk(sc, min(k, r1 - r0), dim=1)
            bi = bi + r0
            if run_v is None:
                run_v, run_i = bv, bi
            else:
                cat_v = torch.cat([run_v, bv], dim=1)
                cat_i = torch.cat([run_i, bi], dim=1)
                run_v, pos = torch.topk(cat_v, min(k, cat_v.shape[1]), dim=1)
                run_i = torch.gather(cat_i, 1, pos)
        kk = run_v.shape[1]
        D[q0:q1, :kk] = run_v
        I[q0:q1, :kk] = run_i
    D = D.numpy()
    I = I.numpy()
    if ids is not None:
        ids = np.asarray(ids, dtype=np.int64)
        I = np.where(I >= 0, ids[np.clip(I, 0, None)], -1)
    return D, I


def flat_ip_search_f64(X, Q, k, ids=None, block: int = 65536):
    """fp64 truth: same semantics, inner products accumulated in double."""
    D, I = flat_ip_search(X, Q, k, ids=ids, block=block, dtype=np.float64)
    return D, I


def merge_topk(D_parts: np.ndarray, I_parts: np.ndarray, k: int) -> tuple[np.ndarray, np.ndarray]:
    """k-way merge of per-shard results [G, nq, k] -> [nq, k].

    Order: score descending, then shard number, then position inside the shard
    (for contiguous row shards this equals ascending global row).
    """
    D_parts = np.asarray(D_parts, dtype=np.float32)
    I_parts = np.asarray(I_parts, dtype=np.int64)
    G, nq, kk = D_parts.shape
    cat_v = np.transpose(D_parts, (1, 0, 2)).reshape(nq, G * kk)
    cat_i = np.transpose(I_parts, (1, 0, 2)).reshape(nq, G * kk)
    order = np.argsort(-cat_v.astype(np.float64), axis=1, kind="stable")[:, :k]
    D = np.take_along_axis(cat_v, order, axis=1)
    I = np.take_along_axis(cat_i, order, axis=1)
    if D.shape[1] < k:
        pad = k - D.shape[1]
        D = np.concatenate([D, np.full((nq, pad), NEG_PAD, np.float32)], axis=1)
        I = np.concatenate([I, np.full((nq, pad), -1, np.int64)], axis=1)
    return D, I


def compare_topk(D, I, D_ref, I_ref, rtol: float = 1e-5, atol: float = 1e-6) -> dict:
    """Tie-aware comparison of two (D, I) results (north_star parity rule).

    scores: |D - D_ref| <= rtol*|D_ref| + atol at every rank.
    ids:    equal rank by rank, except where the reference scores involved lie
            within the same tolerance of each other (a tie band): there the id
            may sit at another rank of the band, or -- at the k-th-rank boundary --
            be replaced by an id whose score is within tolerance of the k-th.
    """
    D = np.asarray(D, dtype=np.float64)
    D_ref = np.asarray(D_ref, dtype=np.float64)
    I = np.asarray(I)
    I_ref = np.asarray(I_ref)
    assert D.shape == D_ref.shape == I.shape == I_ref.shape
    tol = rtol * np.abs(D_ref) + atol
    score_bad = np.abs(D - D_ref) > tol
    pad = (I_ref < 0) & (I < 0)
    score_bad &= ~pad
    pos_mismatch = I != I_ref
    rows = np.nonzero(pos_mismatch.any(axis=1))[0]
    hard = 0
    for r in rows:
        ref_pos = {int(v): p for p, v in enumerate(I_ref[r]) if v >= 0}
        kth = D_ref[r][I_ref[r] >= 0][-1] if (I_ref[r] >= 0).any() else 0.0
        for p in np.nonzero(pos_mismatch[r])[0]:
            v = int(I[r, p])
            t = rtol * abs(D_ref[r, p]) + atol
            if v < 0:
                hard += 1
            elif v in ref_pos:
                if abs(D_ref[r, ref_pos[v]] - D_ref[r, p]) > 2 * t:
                    hard += 1
            else:
                if abs(D[r, p] - kth) > 2 * t:
                    hard += 1
    return {
        "ok": bool(not score_bad.any() and hard == 0),
        "score_violations": int(score_bad.sum()),
        "max_rel_err": float(
            np.max(np.where(pad, 0.0, np.abs(D - D_ref) / np.maximum(np.abs(D_ref), 1e-30)))
        )
        if D.size
        else 0.0,
        "max_abs_err": float(np.max(np.where(pad, 0.0, np.abs(D - D_ref)))) if D.size else 0.0,
        "id_pos_mismatch": int(pos_mismatch.sum()),
        "id_hard_mismatch": int(hard),
        "id_exact_frac": float(1.0 - pos_mismatch.mean()) if I.size else 1.0,
    }
