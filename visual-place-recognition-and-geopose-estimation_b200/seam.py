"""Backbone -> head seam: the backbone's final LayerNorm, the CLS / patch-token split and the cast to the
head's run dtype in one pass (vpr_seam_ln_split_fwd / _bwd).

[UPSTREAM] the hub model's DINOv2 wrapper ends with ``x = self.model.norm(x)``, ``t = x[:, 0]``,
``f = x[:, 1:].reshape(B, h, w, C).permute(0, 3, 1, 2)`` before the aggregator the reference reaches at
dinov2salad/dinov2salad_finetuning.py:36; ``BackboneSeam`` takes the un-normalised last hidden state and
hands ``SaladHead`` its token-major ``(x_tok [B,n,C], t [B,C])`` directly.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _lib as L


class _SeamFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, hidden, weight, bias, eps, skip, out_dtype):
        lib = L.lib()
        B, S, d = hidden.shape
        n = S - 1 - skip
        dev = hidden.device
        x_tok = torch.empty(B, n, d, dtype=out_dtype, device=dev)
        t = torch.empty(B, d, dtype=out_dtype, device=dev)
        need_bwd = any(ctx.needs_input_grad[:3])
        stats = torch.empty(B * (1 + n), 2, dtype=torch.float32, device=dev) if need_bwd else None
        with L.on_device(hidden):
            L.check(lib.vpr_seam_ln_split_fwd(L.ptr(hidden), L.ptr(weight), L.ptr(bias), float(eps), L.ptr(x_tok), L.ptr(t),
                                              L.ptr(stats), B, n, skip, d, L.dtype_code(hidden.dtype),
                                              L.dtype_code(out_dtype), L.stream_ptr(dev)), "vpr_seam_ln_split_fwd")
        if need_bwd:
            ctx.save_for_backward(hidden, weight, stats)
            ctx.skip, ctx.out_dtype = skip, out_dtype
            ctx.set_materialize_grads(False)
        return x_tok, t

    @staticmethod
    def backward(ctx, d_x_tok, d_t):
        lib = L.lib()
        hidden, weight, stats = ctx.saved_tensors
        B, S, d = hidden.shape
        n = S - 1 - ctx.skip
        dev = hidden.device
        d_hidden = torch.empty_like(hidden)
        d_w = torch.empty(d, dtype=torch.float32, device=dev)
        d_b = torch.empty(d, dtype=torch.float32, device=dev)
        dx = d_x_tok.to(ctx.out_dtype).contiguous() if d_x_tok is not None else None
        dt = d_t.to(ctx.out_dtype).contiguous() if d_t is not None else None
        with L.on_device(hidden):
            L.check(lib.vpr_seam_ln_split_bwd(L.ptr(hidden), L.ptr(weight), L.ptr(stats), L.ptr(dx), L.ptr(dt), L.ptr(d_hidden),
                                              L.ptr(d_w), L.ptr(d_b), B, n, ctx.skip, d, L.dtype_code(hidden.dtype),
                                              L.dtype_code(ctx.out_dtype), L.stream_ptr(dev)), "vpr_seam_ln_split_bwd")
        return d_hidden, d_w, d_b, None, None, None


class BackboneSeam(nn.Module):
    """``forward(hidden [B, 1 + registers + n, C]) -> (x_tok [B, n, C], t [B, C])`` in ``out_dtype``.

    ``norm`` carries the backbone's final LayerNorm parameters under the same key names DINOv2 uses
    (``norm.weight`` / ``norm.bias``), so ``seam.norm.load_state_dict(backbone.norm.state_dict())`` wires it in."""

    def __init__(self, dim: int = 768, eps: float = 1e-6, num_register_tokens: int = 0,
                 out_dtype: torch.dtype = torch.bfloat16):
        super().__init__()
        self.norm = nn.LayerNorm(dim, eps=eps)
        self.num_register_tokens = int(num_register_tokens)
        self.out_dtype = out_dtype

    def forward(self, hidden: torch.Tensor):
        L.require_cuda(hidden, "hidden state")
        if hidden.dim() != 3 or hidden.shape[2] != self.norm.normalized_shape[0]:
            raise ValueError(f"expected hidden [B, tokens, {self.norm.normalized_shape[0]}], got {tuple(hidden.shape)}")
        if hidden.shape[1] < 2 + self.num_register_tokens:
            raise ValueError("hidden state has no patch tokens")
        hidden = hidden.contiguous()
        w = self.norm.weight.float().contiguous()
        b = self.norm.bias.float().contiguous()
        return _SeamFn.apply(hidden, w, b, self.norm.eps, self.num_register_tokens, self.out_dtype)
