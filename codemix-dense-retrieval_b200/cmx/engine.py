"""Thin object wrapper over the C ABI: one ``Shard`` = one ``cmx_index`` on one GPU.

Accepts numpy arrays (host memory) and torch tensors (host or CUDA; CUDA tensors
are used zero-copy through ``data_ptr()``).  PyTorch is plumbing only: device
memory, streams.  All arithmetic happens in libcmx.so.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Tuple

import numpy as np

from . import _lib
from ._lib import PATH_AUTO, PATH_STREAM, PATH_TENSOR, check

try:  # torch is optional for pure-numpy callers
    import torch
except Exception:  # pragma: no cover
    torch = None

_PATHS = {"auto": PATH_AUTO, "stream": PATH_STREAM, "tensor": PATH_TENSOR,
          PATH_AUTO: PATH_AUTO, PATH_STREAM: PATH_STREAM, PATH_TENSOR: PATH_TENSOR, None: PATH_AUTO}


def _is_torch(x) -> bool:
    return torch is not None and isinstance(x, torch.Tensor)


def _stream(device: int) -> int:
    if torch is not None and torch.cuda.is_available():
        return int(torch.cuda.current_stream(device).cuda_stream)
    return 0


def _as_f32_2d(x, d: Optional[int], what: str):
    """Coerce like faiss' Python layer: float32, C-contiguous, 2-D; assert the dim."""
    if _is_torch(x):
        if x.dtype != torch.float32 or not x.is_contiguous():
            x = x.to(torch.float32).contiguous()
        if x.dim() == 1:
            x = x.reshape(1, -1)
        assert x.dim() == 2, f"{what} must be 2-D"
    else:
        x = np.ascontiguousarray(x, dtype=np.float32)
        if x.ndim == 1:
            x = x.reshape(1, -1)
        assert x.ndim == 2, f"{what} must be 2-D"
    if d is not None:
        assert x.shape[1] == d, f"{what} has dim {x.shape[1]}, index has d={d}"
    return x


def _ptr(x) -> Tuple[int, int]:
    """(address, on_device)"""
    if _is_torch(x):
        return int(x.data_ptr()), 1 if x.is_cuda else 0
    return int(x.ctypes.data), 0


class Shard:
    """A flat inner-product row store resident on one GPU."""

    def __init__(self, d: int, device: int = 0):
        self._h = C.c_void_p()
        check(_lib.lib().cmx_index_create(int(d), int(device), C.byref(self._h)))
        self.d = int(d)
        self.device = int(device)

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            try:
                _lib.lib().cmx_index_free(h)
            except Exception:
                pass
            self._h = C.c_void_p()

    close = __del__

    # ---- storage -----------------------------------------------------------
    @property
    def ntotal(self) -> int:
        n = C.c_int64(0)
        check(_lib.lib().cmx_index_ntotal(self._h, C.byref(n)))
        return int(n.value)

    def memory(self) -> dict:
        """Device bytes: fp32 row store, fp16 operand plane(s), search workspace, and what a FAISS flat index of the
        same rows would hold (ntotal * d * 4)."""
        out = (C.c_int64 * 4)()
        check(_lib.lib().cmx_index_memory(self._h, out))
        return {"store": int(out[0]), "planes": int(out[1]), "workspace": int(out[2]), "faiss_flat": int(out[3])}

    def reserve(self, n: int) -> None:
        check(_lib.lib().cmx_index_reserve(self._h, int(n)))

    def reset(self) -> None:
        check(_lib.lib().cmx_index_reset(self._h))

    def add(self, x) -> None:
        x = _as_f32_2d(x, self.d, "add(x)")
        if _is_torch(x) and x.is_cuda:
            assert x.device.index == self.device, "tensor lives on another GPU than the shard"
            torch.cuda.current_stream(self.device).synchronize()
        p, on_dev = _ptr(x)
        check(_lib.lib().cmx_index_add(self._h, p, int(x.shape[0]), on_dev))

    def add_from_file(self, path, offset: int, n: int, nthreads: int = 0):
        """Append ``n`` rows stored as little-endian float32 at byte ``offset`` of ``path`` (the vector block of an
        ``index.faiss``): reader threads + page-locked staging buffers overlap the file reads with the PCIe
        copies (``cmx_index_add_from_file``).  Returns (seconds inside pread, wall seconds)."""
        import os

        sec = (C.c_double * 2)()
        check(_lib.lib().cmx_index_add_from_file(self._h, os.fsencode(str(path)), int(offset), int(n), int(nthreads), sec))
        return float(sec[0]), float(sec[1])

    def add_gather(self, src: "Shard", rows) -> None:
        """Append rows ``rows`` (row numbers) of another shard on the same GPU without leaving the device."""
        rows = np.ascontiguousarray(rows, dtype=np.int64)
        check(_lib.lib().cmx_index_add_gather(self._h, src._h, rows.ctypes.data, int(rows.shape[0])))

    def reconstruct_n(self, i0: int, n: int, out=None):
        if out is None:
            out = np.empty((int(n), self.d), dtype=np.float32)
        p, on_dev = _ptr(out)
        check(_lib.lib().cmx_index_reconstruct(self._h, int(i0), int(n), p, on_dev))
        return out

    def data_ptr(self) -> int:
        p = C.c_void_p()
        check(_lib.lib().cmx_index_data(self._h, C.byref(p)))
        return int(p.value or 0)

    def set_precision(self, mode) -> None:
        """'rescore' (default: one fp16 MMA pass + exact fp32 rescoring of a provably sufficient
        candidate superset) or 'split' (three fp16 MMA passes, fp32-faithful tensor-core scores)."""
        m = {"split": _lib.PRECISION_SPLIT, "rescore": _lib.PRECISION_RESCORE}.get(mode, mode)
        check(_lib.lib().cmx_index_set_precision(self._h, int(m)))

    def error_bounds(self):
        """(max row norm, max norm of what the fp16 plane loses of a row) of this shard -- the two corpus
        quantities behind the rescore-mode filter margin (builds the fp16 plane if needed)."""
        out = (C.c_float * 2)()
        check(_lib.lib().cmx_index_error_bounds(self._h, out))
        return float(out[0]), float(out[1])

    def raise_error_bounds(self, norm_max: float, resid_max: float) -> None:
        """Install the maxima over ALL shards (sharded search; see include/cmx.h)."""
        check(_lib.lib().cmx_index_raise_error_bounds(self._h, (C.c_float * 2)(norm_max, resid_max)))

    def set_cand_capacity(self, cap: int) -> None:
        check(_lib.lib().cmx_index_set_cand_capacity(self._h, int(cap)))

    # ---- search ------------------------------------------------------------
    def _alloc_out(self, like, shape_d, shape_i):
        if _is_torch(like) and like.is_cuda:
            D = torch.empty(shape_d, dtype=torch.float32, device=like.device)
            I = torch.empty(shape_i, dtype=torch.int64, device=like.device)
        elif _is_torch(like):
            D = torch.empty(shape_d, dtype=torch.float32, pin_memory=torch.cuda.is_available())
            I = torch.empty(shape_i, dtype=torch.int64, pin_memory=torch.cuda.is_available())
        else:
            D = np.empty(shape_d, dtype=np.float32)
            I = np.empty(shape_i, dtype=np.int64)
        return D, I

    def search(self, x, k: int, id_base: int = 0, path="auto", out=None):
        x = _as_f32_2d(x, self.d, "search(x)")
        k = int(k)
        assert k > 0
        nq = int(x.shape[0])
        D, I = out if out is not None else self._alloc_out(x, (nq, k), (nq, k))
        if nq == 0:
            return D, I
        px, xdev = _ptr(x)
        pd, ddev = _ptr(D)
        pi, idev = _ptr(I)
        assert xdev == ddev == idev, "inputs and outputs must all be host or all be device"
        check(_lib.lib().cmx_index_search(self._h, px, nq, k, pd, pi, xdev, int(id_base), _PATHS[path], _stream(self.device)))
        return D, I

    def search_mixed(self, P, S, alphas: Sequence[float], k: int, id_base: int = 0, path="auto", out=None,
                     want_flags: bool = False):
        """Fused mix+normalise prologue and search for a batch of alphas -> D,I [nA,nq,k]."""
        P = _as_f32_2d(P, self.d, "P")
        S = _as_f32_2d(S, self.d, "S")
        assert tuple(P.shape) == tuple(S.shape)
        nq, nA, k = int(P.shape[0]), len(alphas), int(k)
        D, I = out if out is not None else self._alloc_out(P, (nA, nq, k), (nA, nq, k))
        a = (C.c_double * nA)(*[float(v) for v in alphas])
        pp, pdev = _ptr(P)
        ps, sdev = _ptr(S)
        pd, ddev = _ptr(D)
        pi, idev = _ptr(I)
        assert pdev == sdev == ddev == idev, "inputs and outputs must all be host or all be device"
        flags = None
        pf = None
        if want_flags:
            if pdev:
                flags = torch.empty((nA, nq), dtype=torch.uint8, device=P.device)
                pf = int(flags.data_ptr())
            else:
                flags = np.empty((nA, nq), dtype=np.uint8)
                pf = int(flags.ctypes.data)
        check(_lib.lib().cmx_search_mixed(self._h, pp, ps, nq, a, nA, k, pd, pi, pf, pdev, int(id_base), _PATHS[path],
                                          _stream(self.device)))
        return (D, I, flags) if want_flags else (D, I)

    # ---- two-phase search for sharded indexes (device tensors, rescore precision) --------------
    # Building blocks that only ENQUEUE work on the current stream (include/cmx.h): the caller
    # provides the cross-shard barriers and reads the status word at the end of the step.
    def search_prepare(self, P, S, alphas: Sequence[float]) -> int:
        """Fused mix + normalise of nA alphas into the shard's own query buffer; returns the device
        address of the mixed queries [nA*nq, d]."""
        P = _as_f32_2d(P, self.d, "P")
        S = _as_f32_2d(S, self.d, "S")
        assert P.is_cuda and S.is_cuda and tuple(P.shape) == tuple(S.shape), "two-phase search works on CUDA tensors"
        nA = len(alphas)
        a = (C.c_double * nA)(*[float(v) for v in alphas])
        q = C.c_void_p()
        check(_lib.lib().cmx_search_prepare(self._h, _ptr(P)[0], _ptr(S)[0], int(P.shape[0]), a, nA, C.byref(q), _stream(self.device)))
        return int(q.value)

    def export_bounds(self, out2_ptr: int) -> None:
        """{max row norm, max fp16-residual norm} of this shard -> 2 floats at a device address."""
        check(_lib.lib().cmx_index_export_bounds(self._h, int(out2_ptr), _stream(self.device)))

    def search_begin(self, q_ptr: int, nq: int, k: int, id_base: int, bounds_ptrs: Sequence[int], est_scale: float,
                     scores_ptr: int, flag_ptr: int) -> None:
        """Phase 1 for queries at device address ``q_ptr`` [nq <= 8192, d]: approximate pass; the shard's k
        best approximate scores per query -> ``scores_ptr`` [nq, k], its status word -> ``flag_ptr``."""
        arr = (C.c_void_p * max(1, len(bounds_ptrs)))(*[int(p) for p in bounds_ptrs])
        check(_lib.lib().cmx_search_begin(self._h, int(q_ptr), int(nq), int(k), int(id_base), arr, len(bounds_ptrs),
                                          float(est_scale), int(scores_ptr), int(flag_ptr), _stream(self.device)))

    def search_end(self, kth_ptrs: Sequence[int], D_ptr: int, I_ptr: int) -> None:
        """Phase 2: exact rescoring of the rows that can still reach the global top-k, given the global
        k-th best approximate scores (device pointers, peer memory allowed) -> D, I [nq, k] at device addresses."""
        arr = (C.c_void_p * len(kth_ptrs))(*[int(p) for p in kth_ptrs])
        check(_lib.lib().cmx_search_end(self._h, arr, len(kth_ptrs), int(D_ptr), int(I_ptr), _stream(self.device)))

    def search_ptr(self, q_ptr: int, nq: int, k: int, id_base: int, D_ptr: int, I_ptr: int, path="auto") -> None:
        """``search`` on raw device addresses (queries [nq, d], outputs [nq, k]): the per-chunk fallback of the sharded step."""
        check(_lib.lib().cmx_index_search(self._h, int(q_ptr), int(nq), int(k), int(D_ptr), int(I_ptr), 1, int(id_base),
                                          _PATHS[path], _stream(self.device)))

    def last_stats(self) -> dict:
        st = _lib.SearchStats()
        check(_lib.lib().cmx_index_last_stats(self._h, C.byref(st)))
        return st.as_dict()


def mix_normalize(P, S, alphas: Sequence[float], device: int = 0, want_flags: bool = False):
    """Batched safe_mix: out[a, q] = normalise((1-alpha_a) P[q] + alpha_a S[q]) with the
    reference's endpoint pass-through and non-finite fallback.  [nA, nq, d]."""
    P = _as_f32_2d(P, None, "P")
    S = _as_f32_2d(S, None, "S")
    assert tuple(P.shape) == tuple(S.shape)
    nq, d = int(P.shape[0]), int(P.shape[1])
    nA = len(alphas)
    on_dev = _is_torch(P) and P.is_cuda
    if on_dev:
        assert _is_torch(S) and S.is_cuda and S.device == P.device
        device = P.device.index
        out = torch.empty((nA, nq, d), dtype=torch.float32, device=P.device)
        flags = torch.zeros((nA, nq), dtype=torch.uint8, device=P.device)
    else:
        if _is_torch(P):
            P = P.numpy()
        if _is_torch(S):
            S = S.numpy()
        out = np.empty((nA, nq, d), dtype=np.float32)
        flags = np.zeros((nA, nq), dtype=np.uint8)
    a = (C.c_double * nA)(*[float(v) for v in alphas])
    check(_lib.lib().cmx_mix_normalize(_ptr(P)[0], _ptr(S)[0], nq, d, a, nA, _ptr(out)[0], _ptr(flags)[0],
                                       1 if on_dev else 0, int(device), _stream(device)))
    return (out, flags) if want_flags else out


def union_kth(score_ptrs: Sequence[int], nq: int, k: int, q0: int, q1: int, out_ptrs: Sequence[int], device: int,
              flag_ptrs: Sequence[int] = (), flag_any_ptr: int = 0) -> None:
    """Global k-th best approximate score of queries [q0, q1) over all shards' exported lists
    (device pointers, peer memory allowed), written to every ``out_ptrs[o][q]``; the shards' status
    words ``flag_ptrs`` are OR-ed into the word at ``flag_any_ptr``.  Asynchronous."""
    parts = (C.c_void_p * len(score_ptrs))(*[int(p) for p in score_ptrs])
    outs = (C.c_void_p * len(out_ptrs))(*[int(p) for p in out_ptrs])
    flags = (C.c_void_p * max(1, len(flag_ptrs)))(*[int(p) for p in flag_ptrs])
    check(_lib.lib().cmx_union_kth(parts, len(score_ptrs), int(nq), int(k), int(q0), int(q1), outs, len(out_ptrs),
                                   flags, len(flag_ptrs), int(flag_any_ptr) or None, int(device), _stream(device)))


def merge_topk_peers(D_ptrs: Sequence[int], I_ptrs: Sequence[int], nq: int, k: int, q0: int, q1: int,
                     Dout_ptrs: Sequence[int], Iout_ptrs: Sequence[int], device: int) -> None:
    """Fused exchange + merge of queries [q0, q1): reads every part's [nq, k] lists in place (peer memory)
    and stores the merged rows into every output (device, peer, or the device alias of pinned host
    memory).  Asynchronous; the caller provides the barriers."""
    arr = lambda ptrs: (C.c_void_p * len(ptrs))(*[int(p) for p in ptrs])  # noqa: E731
    check(_lib.lib().cmx_merge_topk_peers(arr(D_ptrs), arr(I_ptrs), len(D_ptrs), int(nq), int(k), int(q0), int(q1),
                                          arr(Dout_ptrs), arr(Iout_ptrs), len(Dout_ptrs), int(device), _stream(device)))


def peer_broadcast(src_ptr: int, dst_ptrs: Sequence[int], nbytes: int, device: int) -> None:
    """``nbytes`` at device address ``src_ptr`` -> every ``dst_ptrs[i]`` (peer memory).  Asynchronous."""
    if not dst_ptrs or nbytes <= 0:
        return
    dsts = (C.c_void_p * len(dst_ptrs))(*[int(p) for p in dst_ptrs])
    check(_lib.lib().cmx_peer_broadcast(int(src_ptr), dsts, len(dst_ptrs), int(nbytes), int(device), _stream(device)))


def host_register(ptr: int, nbytes: int) -> int:
    """Page-lock + device-map a host range; returns the device-side alias."""
    out = C.c_void_p()
    check(_lib.lib().cmx_host_register(int(ptr), int(nbytes), C.byref(out)))
    return int(out.value)


def host_unregister(ptr: int) -> None:
    check(_lib.lib().cmx_host_unregister(int(ptr)))


def merge_topk(D_parts, I_parts, k: Optional[int] = None, device: int = 0):
    """k-way merge of per-shard results [G, nq, k] -> [nq, k] (score desc, shard, position)."""
    on_dev = _is_torch(D_parts) and D_parts.is_cuda
    if on_dev:
        D_parts = D_parts.contiguous()
        I_parts = I_parts.contiguous()
        device = D_parts.device.index
    else:
        D_parts = np.ascontiguousarray(D_parts, dtype=np.float32)
        I_parts = np.ascontiguousarray(I_parts, dtype=np.int64)
    G, nq, kk = (int(v) for v in D_parts.shape)
    k = kk if k is None else int(k)
    assert k == kk, "merge keeps k"
    if on_dev:
        D = torch.empty((nq, k), dtype=torch.float32, device=D_parts.device)
        I = torch.empty((nq, k), dtype=torch.int64, device=D_parts.device)
    else:
        D = np.empty((nq, k), dtype=np.float32)
        I = np.empty((nq, k), dtype=np.int64)
    check(_lib.lib().cmx_merge_topk(_ptr(D_parts)[0], _ptr(I_parts)[0], G, nq, k, _ptr(D)[0], _ptr(I)[0],
                                    1 if on_dev else 0, int(device), _stream(device)))
    return D, I
