"""Oracle for the vector-mix query step (test infrastructure only).

Follows ``safe_mix`` (onepass_dense_mix_run_custom_lang.py:342-377; the copy at
onepass_bilingual_mix_hub_custom_lang.py:390-424 is identical) applied to every
query row, as the alpha loops at :846-867 (mono) and :901-919 (bilingual) do.

Arithmetic restated:
  * ``|alpha| <= 1e-8``      -> the primary row, bit-for-bit            (:350-351)
  * ``|alpha - 1| <= 1e-8``  -> the secondary row, bit-for-bit          (:352-353)
  * otherwise ``mixed = (1.0-alpha)*p + alpha*s`` in numpy: the Python doubles
    ``1.0-alpha`` and ``alpha`` are cast to fp32, two fp32 multiplies and one
    fp32 add, each individually rounded (no FMA)                       (:356)
  * ``normalize_embeddings`` == ``torch.nn.functional.normalize(x, p=2, dim=1)``
    == ``x / max(||x||_2, 1e-12)`` in fp32 (sentence-transformers 5.0.0,
    un-vendored; requirements.txt:10)                                   (:363)
  * any non-finite output element -> fall back to the secondary row if
    ``|alpha| > 0.5`` else the primary row                              (:364-376)
"""

from __future__ import annotations

import numpy as np

EPS_ENDPOINT = 1e-8
NORM_EPS = 1e-12


def _normalize_rows_f32(x: np.ndarray) -> np.ndarray:
    """torch.nn.functional.normalize(p=2, dim=1, eps=1e-12) on the CPU (fp32)."""
    import torch

    t = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32))
    return torch.nn.functional.normalize(t, p=2, dim=1).numpy()


def mix_normalize(P: np.ndarray, S: np.ndarray, alphas) -> tuple[np.ndarray, np.ndarray]:
    """Return (Q[nA, nq, d] float32, fallback_flags[nA, nq] uint8).

    flags: 0 = mixed+normalised (or endpoint pass-through), 1 = non-finite
    fallback to primary, 2 = non-finite fallback to secondary.
    """
    P = np.ascontiguousarray(P, dtype=np.float32)
    S = np.ascontiguousarray(S, dtype=np.float32)
    assert P.shape == S.shape and P.ndim == 2
    alphas = [float(a) for a in alphas]
    nq, d = P.shape
    out = np.empty((len(alphas), nq, d), dtype=np.float32)
    flags = np.zeros((len(alphas), nq), dtype=np.uint8)
    for ai, alpha in enumerate(alphas):
        if abs(alpha) <= EPS_ENDPOINT:
            out[ai] = P
            continue
        if abs(alpha - 1.0) <= EPS_ENDPOINT:
            out[ai] = S
            continue
        with np.errstate(all="ignore"):
            mixed = ((1.0 - alpha) * P + alpha * S).astype(np.float32, copy=False)
        normed = _normalize_rows_f32(mixed)
        bad = ~np.all(np.isfinite(normed), axis=1)
        if bad.any():
            use_secondary = abs(alpha) > 0.5
            normed[bad] = S[bad] if use_secondary else P[bad]
            flags[ai, bad] = 2 if use_secondary else 1
        out[ai] = normed
    return out, flags


def mix_normalize_f64(P: np.ndarray, S: np.ndarray, alphas) -> np.ndarray:
    """fp64 'truth' of the interior-alpha arithmetic (no fallback handling)."""
    P64 = np.asarray(P, dtype=np.float64)
    S64 = np.asarray(S, dtype=np.float64)
    out = np.empty((len(alphas),) + P64.shape, dtype=np.float64)
    for ai, alpha in enumerate(alphas):
        alpha = float(alpha)
        if abs(alpha) <= EPS_ENDPOINT:
            out[ai] = P64
        elif abs(alpha - 1.0) <= EPS_ENDPOINT:
            out[ai] = S64
        else:
            w1 = np.float64(np.float32(1.0 - alpha))
            w2 = np.float64(np.float32(alpha))
            m = w1 * P64 + w2 * S64
            nrm = np.maximum(np.sqrt((m * m).sum(axis=1, keepdims=True)), NORM_EPS)
            with np.errstate(all="ignore"):
                out[ai] = m / nrm
    return out
