"""Geopose heads + losses on the SALAD descriptor, forward and backward fused into one library
call (vpr_heads_loss_fwd_bwd).

Mirrors the reference modules:
  * ``regressor = Sequential(Linear(8448,512), ReLU(), [Dropout(p)], Linear(512,2))`` with MSE or
    Huber(delta) on z-scored lat/lon targets -- dinov2salad/dinov2salad_finetuning.py:28-32,96 and
    dinov2salad/dinov2salad_finetuning_2.py:118-123,154;
  * ``head = Linear(D, 2)`` = [sin, cos] with the atan2 -> degrees -> MAAE loss of
    angle_prediction/dinov2salad/dino_v2_gemini.py:106-141, the unit-normalised acos loss of
    angle_prediction/swin/swin_angle_finetuning_sin_cos.py:56-69, or MSE on [sin, cos]
    (angle_prediction/swin/swin_angle_finetuning_gemini.py:183).
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.nn as nn

from . import _lib as L

_LATLON = {"mse": L.LATLON_MSE, "huber": L.LATLON_HUBER}
_ANGLE = {None: L.ANGLE_NONE, "none": L.ANGLE_NONE, "atan2_maae": L.ANGLE_ATAN2_MAAE, "acos": L.ANGLE_ACOS,
          "mse": L.ANGLE_MSE}


def _run(desc, params, latlon_target, angle_target, cfg, want_d_desc):
    lib = L.lib()
    B, D = desc.shape
    dt = L.dtype_code(desc.dtype)
    dev = desc.device
    shape = L.HeadsShape(B, D, cfg["hidden"], cfg["latlon"], cfg["angle"], cfg["delta"], cfg["loss_scale"],
                         cfg["global_B"] or B, cfg["reg_dropout_p"], cfg["seed"])
    ws_bytes = lib.vpr_heads_workspace_bytes(C.byref(shape), dt)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    sizes = [p.numel() if p is not None else 0 for p in params]      # the angle head may be absent (None, None)
    flat = torch.empty(sum(sizes), dtype=torch.float32, device=dev)
    grads = list(torch.split(flat, sizes))
    ps = L.HeadsParams(*[p.data_ptr() if p is not None else None for p in params])
    gs = L.HeadsGrads(*[g.data_ptr() if p is not None else None for g, p in zip(grads, params)])
    latlon_pred = torch.empty(B, 2, dtype=torch.float32, device=dev)
    sincos_pred = torch.zeros(B, 2, dtype=torch.float32, device=dev)
    losses = torch.empty(2, dtype=torch.float32, device=dev)
    d_desc = torch.empty_like(desc) if want_d_desc else None
    L.check(lib.vpr_heads_loss_fwd_bwd(C.byref(ps), L.ptr(desc), L.ptr(latlon_target), L.ptr(angle_target),
                                       L.ptr(latlon_pred), L.ptr(sincos_pred), L.ptr(losses), C.byref(gs),
                                       L.ptr(d_desc), L.ptr(ws), ws_bytes, C.byref(shape), dt, L.stream_ptr(dev)),
            "vpr_heads_loss_fwd_bwd")
    return (latlon_pred, sincos_pred, losses, [g.view_as(p) if p is not None else None for g, p in zip(grads, params)],
            d_desc, flat)


class _HeadsLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, desc, latlon_target, angle_target, cfg, *params):
        latlon_pred, sincos_pred, losses, grads, d_desc, _ = _run(
            desc, params, latlon_target, angle_target, cfg, ctx.needs_input_grad[0])
        ctx.present = [g is not None for g in grads]
        ctx.save_for_backward(*( [d_desc] if d_desc is not None else [] ), *[g for g in grads if g is not None])
        ctx.has_d_desc = d_desc is not None
        ctx.mark_non_differentiable(latlon_pred, sincos_pred, losses)
        total = losses[0] + losses[1] if cfg["angle"] != L.ANGLE_NONE else losses[0].clone()
        return total, losses, latlon_pred, sincos_pred

    @staticmethod
    def backward(ctx, g_total, g_losses, _g1, _g2):
        saved = list(ctx.saved_tensors)
        d_desc = saved.pop(0) if ctx.has_d_desc else None
        scale = g_total
        out_d = (d_desc * scale.to(d_desc.dtype)) if d_desc is not None else None
        it = iter(saved)
        pg = [next(it) * scale if present else None for present in ctx.present]
        return (out_d, None, None, None, *pg)


class GeoposeHeads(nn.Module):
    """lat/lon regressor + sin/cos angle head on a descriptor batch, with fused losses."""

    def __init__(self, in_dim: int = 8448, hidden: int = 512, dropout: float = 0.0, latlon_loss: str = "mse",
                 angle_loss: str | None = "atan2_maae", huber_delta: float = 1.0):
        super().__init__()
        layers = [nn.Linear(in_dim, hidden), nn.ReLU()]
        if dropout > 0:
            layers.append(nn.Dropout(dropout))
        layers.append(nn.Linear(hidden, 2))
        self.regressor = nn.Sequential(*layers)       # keys regressor.0.*, regressor.2|3.* as in the reference
        self.head = nn.Linear(in_dim, 2)              # [sin, cos]
        self.in_dim, self.hidden = in_dim, hidden
        self.reg_dropout_p = float(dropout)
        if latlon_loss not in _LATLON or angle_loss not in _ANGLE:
            raise ValueError("unknown loss variant")
        self.latlon_loss, self.angle_loss, self.huber_delta = latlon_loss, angle_loss, float(huber_delta)
        self._step = 0
        self.seed = 0xA17

    def _params(self):
        last = self.regressor[-1]
        return [self.regressor[0].weight, self.regressor[0].bias, last.weight, last.bias, self.head.weight,
                self.head.bias]

    def _cfg(self, loss_scale=1.0, global_B=0):
        seed = (self.seed * 1000003 + self._step) & 0xFFFFFFFFFFFFFFFF
        if self.training:
            self._step += 1
        return {"hidden": self.hidden, "latlon": _LATLON[self.latlon_loss], "angle": _ANGLE[self.angle_loss],
                "delta": self.huber_delta, "loss_scale": float(loss_scale), "global_B": int(global_B),
                "reg_dropout_p": self.reg_dropout_p if self.training else 0.0, "seed": seed}

    def forward(self, desc: torch.Tensor, latlon_target: torch.Tensor, angle_target_deg: torch.Tensor | None = None,
                global_B: int = 0):
        """Returns (total_loss, losses[2] = (lat/lon, angle), latlon_pred [B,2], sincos_pred [B,2]).
        ``total_loss.backward()`` delivers the gradients that the same kernel call already computed."""
        L.require_cuda(desc, "descriptor")
        desc = desc.contiguous()
        B = desc.shape[0]
        lt = latlon_target.to(device=desc.device, dtype=torch.float32).contiguous()
        if self.angle_loss in (None, "none"):
            at = None
        else:
            if angle_target_deg is None:
                raise ValueError("angle_target_deg is required for the angle head")
            at = angle_target_deg.to(device=desc.device, dtype=torch.float32).contiguous()
        params = [p.contiguous() for p in self._params()]
        return _HeadsLossFn.apply(desc, lt, at, self._cfg(global_B=global_B), *params)

    @torch.no_grad()
    def predict(self, desc: torch.Tensor):
        """(latlon_pred [B,2], sincos_pred [B,2]) without targets (eval loops of the reference,
        dinov2salad_finetuning.py:147, dinov2salad_validation.py:81)."""
        B = desc.shape[0]
        z = torch.zeros(B, 2, dtype=torch.float32, device=desc.device)
        a = torch.zeros(B, dtype=torch.float32, device=desc.device)
        cfg = self._cfg()
        cfg["reg_dropout_p"] = 0.0
        latlon_pred, sincos_pred, *_ = _run(desc.contiguous(), [p.contiguous() for p in self._params()], z,
                                            a if cfg["angle"] != L.ANGLE_NONE else None, cfg, False)
        return latlon_pred, sincos_pred


def sincos_to_degrees(sin_cos: torch.Tensor) -> torch.Tensor:
    """reference angle_prediction/dinov2salad/dino_v2_gemini.py:135-141 (host-side convenience)."""
    return torch.rad2deg(torch.atan2(sin_cos[:, 0], sin_cos[:, 1])) % 360.0
