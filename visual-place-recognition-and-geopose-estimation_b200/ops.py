"""Thin Python handle on the library's GEMM (the building block of every dense step).  Used by
the GPU unit tests to validate each operand layout of the tcgen05 kernel in isolation."""
from __future__ import annotations

import torch

from . import _lib as L


def gemm(A, B, M, N, K, *, a_mn=False, b_mn=False, batch=1, a_bs=0, b_bs=0, out_kind=L.OUT_F32, bias=None,
         relu=False, mask=None, mask_scale=1.0, alpha=1.0, k_splits=1, bn=0, out=None, ldo=None, out_bs=None,
         colsum=None, dropout_p=0.0, dropout_seed=0):
    """out[b,m,n] = epilogue(alpha * sum_k A(b,m,k) B(b,n,k)); A/B are 2-D/3-D row-major tensors,
    stored [rows=M|N, K] (x_mn False) or [rows=K, M|N] (x_mn True)."""
    L.require_cuda(A, "A")
    dt = L.dtype_code(A.dtype)
    lda = A.stride(-2)
    ldb = B.stride(-2)
    dev = A.device
    if out is None:
        odt = torch.bfloat16 if out_kind == L.OUT_BF16 else torch.float32
        out = torch.zeros(batch, M, N, dtype=odt, device=dev)
    ldo = ldo if ldo is not None else out.stride(-2)
    out_bs = out_bs if out_bs is not None else (out.stride(0) if out.dim() == 3 else 0)
    ld_mask = mask.stride(-2) if mask is not None else 0
    mask_bs = (mask.stride(0) if (mask is not None and mask.dim() == 3) else 0)
    with L.on_device(A):
        L.check(L.lib().vpr_gemm(dt, L.ptr(A), int(a_mn), lda, a_bs, L.ptr(B), int(b_mn), ldb, b_bs, M, N, K, batch,
                                 L.ptr(out), out_kind, ldo, out_bs, L.ptr(bias), int(relu), L.ptr(mask), ld_mask,
                                 mask_bs, float(mask_scale), float(alpha), k_splits, bn, L.ptr(colsum), float(dropout_p),
                                 int(dropout_seed), L.stream_ptr(dev)),
                "vpr_gemm")
    return out
