"""Geopose heads + losses on the SALAD descriptor, forward and backward fused into one library
call (vpr_heads_loss_fwd_bwd).

Mirrors the reference modules:
  * ``regressor = Sequential(Linear(8448,512), ReLU(), [Dropout(p)], Linear(512,2))`` with MSE or
    Huber(delta) on z-scored lat/lon targets -- dinov2salad/dinov2salad_finetuning.py:28-32,96 and
    dinov2salad/dinov2salad_finetuning_2.py:118-123,154;
  * ``dropout = Dropout(p); head = Linear(D, 2)`` = [sin, cos] with the atan2 -> degrees -> MAAE loss of
    angle_prediction/dinov2salad/dino_v2_gemini.py:99-141, the unit-normalised acos loss of
    angle_prediction/swin/swin_angle_finetuning_sin_cos.py:56-69, or MSE on [sin, cos]
    (angle_prediction/swin/swin_angle_finetuning_gemini.py:183);
  * ``head = Linear(D, 1)``, ``pred = head(x).squeeze(-1) % 360`` with the MAAE loss of
    angle_prediction/dinov2salad/dino_angle_finetuning.py:37-52 (``angle_loss="scalar_maae"``).
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.nn as nn

from . import _lib as L

_LATLON = {"mse": L.LATLON_MSE, "huber": L.LATLON_HUBER, "external": L.LATLON_EXTERNAL}
_ANGLE = {None: L.ANGLE_NONE, "none": L.ANGLE_NONE, "atan2_maae": L.ANGLE_ATAN2_MAAE, "acos": L.ANGLE_ACOS,
          "mse": L.ANGLE_MSE, "scalar_maae": L.ANGLE_SCALAR_MAAE}
_U64 = 0xFFFFFFFFFFFFFFFF


def make_cfg(hidden, latlon, angle=L.ANGLE_NONE, delta=1.0, loss_scale=1.0, global_B=0, reg_dropout_p=0.0, seed=0,
             head_dropout_p=0.0, seed_dev=None):
    return {"hidden": int(hidden), "latlon": int(latlon), "angle": int(angle), "delta": float(delta),
            "loss_scale": float(loss_scale), "global_B": int(global_B), "reg_dropout_p": float(reg_dropout_p),
            "seed": int(seed) & _U64, "head_dropout_p": float(head_dropout_p), "seed_dev": seed_dev}


def _shape(B, D, cfg, flags):
    return L.HeadsShape(B, D, cfg["hidden"], cfg["latlon"], cfg["angle"], cfg["delta"], cfg["loss_scale"],
                        cfg["global_B"] or B, cfg["reg_dropout_p"], cfg["seed"], cfg["head_dropout_p"], flags,
                        cfg["seed_dev"].data_ptr() if cfg.get("seed_dev") is not None else None)


def workspace_bytes(B, D, cfg, dtype) -> int:
    shape = _shape(B, D, cfg, 0)
    return int(L.lib().vpr_heads_workspace_bytes(C.byref(shape), L.dtype_code(dtype)))


def pack_weights(params, ws, B, D, cfg, dtype) -> None:
    """vpr_heads_pack: cast [regressor.0.weight ; angle head] to the run dtype into the head of ``ws``."""
    shape = _shape(B, D, cfg, 0)
    ps = L.HeadsParams(*[p.data_ptr() if p is not None else None for p in params])
    with L.on_device(ws):
        L.check(L.lib().vpr_heads_pack(C.byref(ps), L.ptr(ws), ws.numel(), C.byref(shape), L.dtype_code(dtype),
                                       L.stream_ptr(ws.device)), "vpr_heads_pack")


class _HeadsCall:
    """Buffers and argument structs of ONE vpr_heads_loss_fwd_bwd call.  Split from the launch so that the part that
    does not depend on the descriptor -- clearing the call's accumulators -- can be issued ahead of time on another
    stream (``zero()``, then ``run(..., prezeroed=True)``)."""

    def __init__(self, B, D, dtype, dev, params, cfg, want_d_desc, ws=None, forward_only=False, flat=None):
        lib = L.lib()
        self.B, self.D, self.cfg, self.dev, self.forward_only = B, D, cfg, dev, forward_only
        self.dt = L.dtype_code(dtype)
        shape = _shape(B, D, cfg, 0)
        ws_bytes = lib.vpr_heads_workspace_bytes(C.byref(shape), self.dt)
        if ws is None:
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        elif ws.numel() < ws_bytes or ws.device != dev:
            raise RuntimeError("heads workspace too small or on another device")
        self.ws = ws
        self.ps = L.HeadsParams(*[p.data_ptr() if p is not None else None for p in params])
        self._params = params                       # keep the (possibly .contiguous()) tensors alive
        self.latlon_pred = torch.empty(B, 2, dtype=torch.float32, device=dev)
        self.sincos_pred = (torch.empty if cfg["angle"] != L.ANGLE_NONE else torch.zeros)(B, 2, dtype=torch.float32, device=dev)
        if forward_only:
            self.flat = torch.empty(3, dtype=torch.float32, device=dev)
            self.losses, self.grads, self.gs, self.d_desc = self.flat, [None] * len(params), None, None
        else:
            sizes = [p.numel() if p is not None else 0 for p in params]      # the angle head may be absent (None, None)
            if flat is None:
                flat = torch.empty(sum(sizes) + 3, dtype=torch.float32, device=dev)
            elif flat.numel() < sum(sizes) + 3 or flat.dtype != torch.float32 or flat.device != dev:
                raise RuntimeError("flat gradient buffer must be an fp32 tensor on the descriptor's device")
            self.flat = flat
            *grads, self.losses = torch.split(flat[: sum(sizes) + 3], sizes + [3])
            self.gs = L.HeadsGrads(*[g.data_ptr() if p is not None else None for g, p in zip(grads, params)])
            self.grads = [g.view_as(p) if p is not None else None for g, p in zip(grads, params)]
            self.d_desc = torch.empty(B, D, dtype=dtype, device=dev) if want_d_desc else None

    def _flags(self, extra=0):
        return (L.HEADS_FORWARD_ONLY if self.forward_only else 0) | extra

    def zero(self):
        """vpr_heads_zero on the current stream."""
        shape = _shape(self.B, self.D, self.cfg, self._flags())
        with torch.cuda.device(self.dev):
            L.check(L.lib().vpr_heads_zero(L.ptr(self.losses), C.byref(self.gs) if self.gs is not None else None,
                                           L.ptr(self.ws), self.ws.numel(), C.byref(shape), self.dt,
                                           L.stream_ptr(self.dev)), "vpr_heads_zero")

    def _call(self, desc, latlon_target, angle_target, flags):
        shape = _shape(self.B, self.D, self.cfg, flags)
        with L.on_device(desc):
            L.check(L.lib().vpr_heads_loss_fwd_bwd(
                C.byref(self.ps), L.ptr(desc), L.ptr(latlon_target), L.ptr(angle_target), L.ptr(self.latlon_pred),
                L.ptr(self.sincos_pred), L.ptr(self.losses), C.byref(self.gs) if self.gs is not None else None,
                L.ptr(self.d_desc), L.ptr(self.ws), self.ws.numel(), C.byref(shape), self.dt, L.stream_ptr(self.dev)),
                "vpr_heads_loss_fwd_bwd")

    def run(self, desc, latlon_target, angle_target, prepacked=False, defer_wgrad=False, prezeroed=False):
        cfg = self.cfg
        base = self._flags((L.HEADS_PREPACKED if prepacked else 0) | (L.HEADS_PREZEROED if prezeroed else 0))
        self._call(desc, latlon_target, angle_target, base | (L.HEADS_SKIP_WGRAD if defer_wgrad else 0))
        seed_dev = cfg.get("seed_dev")
        if seed_dev is not None and not self.forward_only and (cfg["reg_dropout_p"] > 0 or cfg["head_dropout_p"] > 0):
            with L.on_device(desc):
                L.check(L.lib().vpr_counter_advance(L.ptr(seed_dev), L.stream_ptr(self.dev)), "vpr_counter_advance")
        out = (self.latlon_pred, self.sincos_pred, self.losses, self.grads, self.d_desc, self.flat)
        if defer_wgrad:
            return out + ((lambda: self._call(desc, latlon_target, angle_target, self._flags(L.HEADS_WGRAD_ONLY))),)
        return out


def _run(desc, params, latlon_target, angle_target, cfg, want_d_desc, ws=None, forward_only=False, prepacked=False,
         defer_wgrad=False, flat=None):
    """One vpr_heads_loss_fwd_bwd call.  Returns (latlon_pred, sincos_pred, losses, grads, d_desc, flat): ``losses`` =
    (lat/lon, angle, their sum); ``flat`` is ONE fp32 buffer holding every parameter gradient back to back followed by
    the three losses, so that a data-parallel caller sums gradients and losses across ranks with a single collective.
    ``defer_wgrad``: leave out the two weight-gradient launches nothing downstream of ``d_desc`` waits for and return a
    seventh value, a callable that issues them on whatever stream is current when it is called (the caller orders
    that stream after this call and joins it before reading ``regressor.0.weight.grad`` / ``head.weight.grad``)."""
    B, D = desc.shape
    call = _HeadsCall(B, D, desc.dtype, desc.device, params, cfg, want_d_desc, ws=ws, forward_only=forward_only, flat=flat)
    return call.run(desc, latlon_target, angle_target, prepacked=prepacked, defer_wgrad=defer_wgrad)


class _HeadsLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, desc, latlon_target, angle_target, cfg, ws, prepacked, *params):
        latlon_pred, sincos_pred, losses, grads, d_desc, _ = _run(
            desc, params, latlon_target, angle_target, cfg, ctx.needs_input_grad[0], ws=ws, prepacked=prepacked)
        ctx.present = [g is not None for g in grads]
        ctx.save_for_backward(*( [d_desc] if d_desc is not None else [] ), *[g for g in grads if g is not None])
        ctx.has_d_desc = d_desc is not None
        ctx.mark_non_differentiable(latlon_pred, sincos_pred, losses)
        return losses[2].clone(), losses, latlon_pred, sincos_pred

    @staticmethod
    def backward(ctx, g_total, g_losses, _g1, _g2):
        saved = list(ctx.saved_tensors)
        d_desc = saved.pop(0) if ctx.has_d_desc else None
        scale = g_total
        out_d = (d_desc * scale.to(d_desc.dtype)) if d_desc is not None else None
        it = iter(saved)
        pg = [next(it) * scale if present else None for present in ctx.present]
        return (out_d, None, None, None, None, None, *pg)


class _RegressorFn(torch.autograd.Function):
    """``preds = regressor(features)`` with a real autograd edge, for the reference's training idiom
    ``preds = model(inputs); loss = loss_fn(preds, targets); loss.backward()`` (dinov2salad_finetuning.py:120-124).
    forward = the fused call in forward-only mode; backward = the same call again with the caller's dL/dpreds
    (VPR_LATLON_EXTERNAL) -- same seed, so the regressor's dropout mask is the one the forward drew."""

    @staticmethod
    def forward(ctx, desc, cfg, *params):
        B = desc.shape[0]
        zeros = torch.zeros(B, 2, dtype=torch.float32, device=desc.device)
        cfg_f = dict(cfg, seed_dev=None)
        preds, *_ = _run(desc, params, zeros, None, cfg_f, False, forward_only=True)
        ctx.cfg = cfg_f
        ctx.save_for_backward(desc, *[p for p in params if p is not None])
        ctx.present = [p is not None for p in params]
        return preds

    @staticmethod
    def backward(ctx, d_preds):
        desc, *ps = ctx.saved_tensors
        it = iter(ps)
        params = [next(it) if present else None for present in ctx.present]
        cfg = dict(ctx.cfg, latlon=L.LATLON_EXTERNAL, angle=L.ANGLE_NONE)
        _, _, _, grads, d_desc, _ = _run(desc, params, d_preds.float().contiguous(), None, cfg,
                                         ctx.needs_input_grad[0])
        return (d_desc, None, *grads)


class GeoposeHeads(nn.Module):
    """lat/lon regressor + angle head on a descriptor batch, with fused losses.

    ``dropout``: the regressor's Dropout between ReLU and the last Linear (dinov2salad_finetuning_2.py:121);
    ``head_dropout``: Dropout on the angle head's input (dino_v2_gemini.py:104,111, rate 0.1 there)."""

    def __init__(self, in_dim: int = 8448, hidden: int = 512, dropout: float = 0.0, latlon_loss: str = "mse",
                 angle_loss: str | None = "atan2_maae", huber_delta: float = 1.0, head_dropout: float = 0.0):
        super().__init__()
        if latlon_loss not in ("mse", "huber") or angle_loss not in _ANGLE:
            raise ValueError("unknown loss variant")
        layers = [nn.Linear(in_dim, hidden), nn.ReLU()]
        if dropout > 0:
            layers.append(nn.Dropout(dropout))
        layers.append(nn.Linear(hidden, 2))
        self.regressor = nn.Sequential(*layers)       # keys regressor.0.*, regressor.2|3.* as in the reference
        self.dropout = nn.Dropout(head_dropout)       # dino_v2_gemini.py:104 (no parameters; the kernel applies it)
        self.head = nn.Linear(in_dim, 1 if angle_loss == "scalar_maae" else 2)    # [sin, cos] | angle in degrees
        self.in_dim, self.hidden = in_dim, hidden
        self.reg_dropout_p = float(dropout)
        self.head_dropout_p = float(head_dropout)
        self.latlon_loss, self.angle_loss, self.huber_delta = latlon_loss, angle_loss, float(huber_delta)
        self._step = 0
        self.seed = 0xA17
        self._seed_dev = None
        self._ws = {}          # (B, dtype, device) -> [workspace, pack key]

    def use_device_step_counter(self, enabled: bool = True) -> None:
        """Keep the dropout step counter on the device (needed under CUDA-graph capture: a host counter would be
        frozen into the graph and every replay would draw the same masks)."""
        self._seed_dev = torch.zeros(1, dtype=torch.int64, device=self.head.weight.device) if enabled else None

    def _params(self):
        last = self.regressor[-1]
        return [self.regressor[0].weight, self.regressor[0].bias, last.weight, last.bias, self.head.weight,
                self.head.bias]

    def _cfg(self, loss_scale=1.0, global_B=0):
        seed = (self.seed * 1000003 + self._step) & _U64
        drop = self.training and (self.reg_dropout_p > 0 or self.head_dropout_p > 0)
        seed_dev = self._seed_dev if drop else None
        if drop and seed_dev is None:
            self._step += 1
        return make_cfg(self.hidden, _LATLON[self.latlon_loss], _ANGLE[self.angle_loss], self.huber_delta, loss_scale,
                        global_B, self.reg_dropout_p if self.training else 0.0, seed,
                        self.head_dropout_p if self.training else 0.0, seed_dev)

    def current_seed(self) -> int:
        """Seed the NEXT training call will use (synchronises when the counter lives on the device)."""
        base = (self.seed * 1000003 + self._step) & _U64
        if self._seed_dev is not None:
            base = (base + int(self._seed_dev.item())) & _U64
        return base

    def _workspace(self, desc, cfg, params):
        """Per-shape persistent workspace whose head holds the run-dtype copy of the big weight matrices; the copy is
        redone only when regressor.0.weight / head.weight changed (tensor version counters) since the last call."""
        B, D = desc.shape
        key = (B, desc.dtype, desc.device)
        ent = self._ws.get(key)
        if ent is None:
            ent = [torch.empty(workspace_bytes(B, D, cfg, desc.dtype), dtype=torch.uint8, device=desc.device), None]
            self._ws[key] = ent
        pk = tuple((p.data_ptr(), p._version) for p in (params[0], params[4])) + (cfg["angle"],)
        prepacked = ent[1] == pk
        ent[1] = pk
        return ent[0], prepacked

    def forward(self, desc: torch.Tensor, latlon_target: torch.Tensor, angle_target_deg: torch.Tensor | None = None,
                global_B: int = 0):
        """Returns (total_loss, losses[3] = (lat/lon, angle, sum), latlon_pred [B,2], sincos_pred [B,2]).
        ``total_loss.backward()`` delivers the gradients that the same kernel call already computed.
        With ``angle_loss="scalar_maae"`` ``sincos_pred[:, 0]`` is the predicted angle in [0, 360)."""
        L.require_cuda(desc, "descriptor")
        desc = desc.contiguous()
        lt = latlon_target.to(device=desc.device, dtype=torch.float32).contiguous()
        if self.angle_loss in (None, "none"):
            at = None
        else:
            if angle_target_deg is None:
                raise ValueError("angle_target_deg is required for the angle head")
            at = angle_target_deg.to(device=desc.device, dtype=torch.float32).contiguous()
        params = [p.contiguous() for p in self._params()]
        cfg = self._cfg(global_B=global_B)
        ws, prepacked = self._workspace(desc, cfg, params)
        return _HeadsLossFn.apply(desc, lt, at, cfg, ws, prepacked, *params)

    @torch.no_grad()
    def predict(self, desc: torch.Tensor):
        """(latlon_pred [B,2], sincos_pred [B,2]) without targets (eval loops of the reference,
        dinov2salad_finetuning.py:147, dinov2salad_validation.py:81): forward-only call, no gradient work."""
        L.require_cuda(desc, "descriptor")
        desc = desc.contiguous()
        B = desc.shape[0]
        z = torch.zeros(B, 2, dtype=torch.float32, device=desc.device)
        a = torch.zeros(B, dtype=torch.float32, device=desc.device)
        cfg = make_cfg(self.hidden, _LATLON[self.latlon_loss], _ANGLE[self.angle_loss], self.huber_delta)
        params = [p.contiguous() for p in self._params()]
        ws, prepacked = self._workspace(desc, cfg, params)
        latlon_pred, sincos_pred, *_ = _run(desc, params, z, a if cfg["angle"] != L.ANGLE_NONE else None, cfg, False,
                                            ws=ws, forward_only=True, prepacked=prepacked)
        return latlon_pred, sincos_pred


def sincos_to_degrees(sin_cos: torch.Tensor) -> torch.Tensor:
    """reference angle_prediction/dinov2salad/dino_v2_gemini.py:135-141 (host-side convenience)."""
    return torch.rad2deg(torch.atan2(sin_cos[:, 0], sin_cos[:, 1])) % 360.0
