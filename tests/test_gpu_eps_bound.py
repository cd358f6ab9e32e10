"""Direct measurement of the rescore precision's error bound.

The default arithmetic selects on ONE fp16 tensor-core pass and proves that the exact top-k is
inside the kept candidates from |approx - exact| <= eps(q).  Every other test checks the final
(D, I); this one measures the bound itself: the approximate scores the scoring kernel produces
(``cmx_debug_approx_scores``: the dense epilogue of tc_score_kernel on chosen slabs) against fp64
inner products, for >= 1e8 (query, row) pairs per data set, and reports max |approx - exact| / eps(q).
"""
import ctypes

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

D_ = 1024
NQ = 2048
N = 1_048_576 + 300  # a ragged tail block
SLAB = 8192          # rows per hook call = the candidate capacity at k <= 1024
CALLS = 7            # 7 x 2048 x 8192 = 1.17e8 pairs


def _approx(sh, Q, pos0, nrows):
    import torch

    from cmx import _lib

    nq = Q.shape[0]
    sc = torch.empty((nq, nrows), dtype=torch.float32, device="cuda")
    rows = torch.empty((nq, nrows), dtype=torch.int64, device="cuda")
    margin = torch.empty((nq,), dtype=torch.float32, device="cuda")
    _lib.check(_lib.lib().cmx_debug_approx_scores(sh._h, Q.data_ptr(), nq, pos0, nrows, sc.data_ptr(), rows.data_ptr(),
                                                  margin.data_ptr(), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
    return sc, rows, margin


def _measure(X, Q, label):
    """max over pairs of |approx - exact| / eps(q), eps = margin / 2; the slabs are spread over the corpus."""
    import torch

    from cmx.engine import Shard

    sh = Shard(D_, 0)
    sh.set_precision("rescore")
    sh.add(X)
    n = X.shape[0]
    npos = (n + 255) // 256 * 256
    worst, pairs = 0.0, 0
    Q64 = Q.double()
    for c in range(CALLS):
        pos0 = min((npos - SLAB) // 256 * 256, (c * (npos - SLAB) // max(1, CALLS - 1)) // 256 * 256)
        sc, rows, margin = _approx(sh, Q, pos0, SLAB)
        r = rows[0]
        assert bool((rows == r[None, :]).all())  # the position -> row map does not depend on the query
        valid = r >= 0
        assert int(valid.sum()) >= SLAB - 256  # only the ragged tail block has padding positions
        exact = Q64 @ X[r[valid]].double().T
        err = (sc[:, valid].double() - exact).abs()
        eps = (margin.double() / 2.0)[:, None]
        assert bool((eps > 0).all()) and bool(torch.isfinite(eps).all())
        worst = max(worst, float((err / eps).max()))
        pairs += int(err.numel())
    print(f"[eps-bound] {label}: max |approx - exact| / eps(q) = {worst:.4f} over {pairs:.3e} pairs")
    return worst, pairs


def _unit(x):
    import torch

    return torch.nn.functional.normalize(x, dim=1)


def _data(kind):
    import torch

    g = torch.Generator(device="cuda").manual_seed(2024)
    X = torch.randn((N, D_), generator=g, device="cuda")
    Q = torch.randn((NQ, D_), generator=g, device="cuda")
    if kind == "iid":
        return _unit(X), _unit(Q)
    if kind == "aniso":  # one shared direction: cosines ~0.7 like real embedding scores
        u = _unit(torch.randn((1, D_), generator=g, device="cuda"))
        return _unit(X + 1.5 * (D_ ** 0.5) * u), _unit(Q + 1.5 * (D_ ** 0.5) * u)
    if kind == "wide":  # row / query norms over four decades, heavy-tailed coordinates
        X = X * torch.exp(2.0 * torch.randn((N, D_), generator=g, device="cuda"))
        X = _unit(X) * torch.pow(10.0, torch.rand((N, 1), generator=g, device="cuda") * 4 - 2)
        Q = _unit(Q * torch.exp(2.0 * torch.randn((NQ, D_), generator=g, device="cuda")))
        Q = Q * torch.pow(10.0, torch.rand((NQ, 1), generator=g, device="cuda") * 4 - 2)
        return X.contiguous(), Q.contiguous()
    raise ValueError(kind)


@pytest.mark.parametrize("kind", ["iid", "aniso", "wide"])
def test_eps_bound_holds_with_margin(kind):
    """>= 1e8 pairs per data set: the observed error stays below eps(q) / 1.5."""
    X, Q = _data(kind)
    worst, pairs = _measure(X, Q, kind)
    assert pairs >= 1e8
    assert worst <= 1.0 / 1.5, worst


def test_eps_bound_adversarial_alignment():
    """Rows aligned with the queries' own fp16 residuals drive the Cauchy-Schwarz term q_lo . x towards
    equality: the bound must still hold (<= 1), and the measurement shows how much of it can be reached."""
    import torch

    X, Q = _data("iid")
    # the kernel's query scale: a power of two placing max|q| of the BATCH in [2^12, 2^13)
    amax = float(Q.abs().max())
    scale = 2.0 ** (12 - int(np.floor(np.log2(amax))))
    qlo = Q - (Q * scale).half().float() / scale
    X = X.clone()
    X[:NQ] = _unit(qlo)
    X[NQ : 2 * NQ] = -_unit(qlo)
    from cmx import _lib

    _lib.check(_lib.lib().cmx_debug_set_block_order(0))  # file order: the first slab measured is rows 0..8191
    try:
        worst, _ = _measure(X, Q, "aligned rows")
    finally:
        _lib.check(_lib.lib().cmx_debug_set_block_order(1))
    assert worst <= 1.0, worst


def test_tensor_core_accumulation_error_is_inside_gamma():
    """Operands that ARE fp16 numbers (no rounding residual on either side): what remains of
    |approx - exact| is the fp32 accumulation inside tcgen05.mma -- the part of eps(q) that rests on
    a hardware assumption (gamma = 2 d 2^-23).  Measured against gamma ||q|| max||x||."""
    import torch

    from cmx.engine import Shard

    g = torch.Generator(device="cuda").manual_seed(7)
    # integers / 64 with |value| <= 2047/64: exact in fp16 after the power-of-two plane scaling
    X = torch.randint(-2047, 2048, (N, D_), generator=g, device="cuda").float() / 64.0
    Q = torch.randint(-2047, 2048, (NQ, D_), generator=g, device="cuda").float() / 64.0
    sh = Shard(D_, 0)
    sh.set_precision("rescore")
    sh.add(X)
    nb, rb = sh.error_bounds()
    assert rb == 0.0, rb  # the hi plane loses nothing of these rows
    gamma = 2.0 * D_ * 2.0 ** -23
    worst = 0.0
    npos = (N + 255) // 256 * 256
    for c in range(3):
        pos0 = (c * (npos - SLAB) // 2) // 256 * 256
        sc, rows, margin = _approx(sh, Q, pos0, SLAB)
        r = rows[0]
        valid = r >= 0
        exact = Q.double() @ X[r[valid]].double().T
        err = (sc[:, valid].double() - exact).abs()
        bound = gamma * Q.double().norm(dim=1)[:, None] * nb
        worst = max(worst, float((err / bound).max()))
    print(f"[eps-bound] fp16-exact operands: max accumulation error / (gamma ||q|| X) = {worst:.4f}")
    assert worst <= 0.5, worst
