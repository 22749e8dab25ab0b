"""Host-side mirror of the reference's model classes around the hot path, and checkpoint compatibility.

  * ``VPRModel``  -- [UPSTREAM] hub model layout: ``backbone`` + ``aggregator`` (the module
    ``torch.hub.load("serizba/salad", "dinov2_salad")`` returns at dinov2salad/dinov2salad_finetuning.py:19),
    with ``SaladHead`` as the aggregator and, optionally, ``BackboneSeam`` between the two.
  * ``DINOv2RegressionModel`` -- dinov2salad/dinov2salad_finetuning.py:21-37 (dup dinov2salad_validation.py:36-52):
    frozen ``feature_extractor`` + ``regressor``; same constructor argument, same ``state_dict`` keys.
  * ``load_reference_checkpoint`` / ``save_reference_checkpoint`` -- the ``checkpoint_*.pth`` dicts written at
    dinov2salad_finetuning.py:130-135 and read at dinov2salad_validation.py:68-69.
  * ``load_upstream_salad_weights`` -- [UPSTREAM] ``dino_salad.ckpt`` (``aggregator.*`` keys).
"""
from __future__ import annotations

from typing import Mapping

import torch
import torch.nn as nn

from . import _lib as L
from .heads import _HeadsLossFn, _RegressorFn, _run, make_cfg
from .salad import SaladHead

_LATLON = {"mse": L.LATLON_MSE, "huber": L.LATLON_HUBER}


class VPRModel(nn.Module):
    """``forward(images) -> descriptor``.  ``backbone(images)`` returns either ``(f, t)`` as the upstream DINOv2
    wrapper does, or -- when ``seam`` is given -- the un-normalised last hidden state ``[B, 1 + R + n, C]``."""

    def __init__(self, backbone: nn.Module, aggregator: SaladHead | None = None, seam: nn.Module | None = None):
        super().__init__()
        self.backbone = backbone
        self.aggregator = aggregator if aggregator is not None else SaladHead()
        if seam is not None:
            self.seam = seam

    def forward(self, x):
        feats = self.backbone(x)
        if hasattr(self, "seam"):
            feats = self.seam(feats)
        return self.aggregator(feats)


class DINOv2RegressionModel(nn.Module):
    """Frozen descriptor extractor + lat/lon regressor.

    ``forward(x)`` is the reference's call ``preds = model(inputs)`` in both of its uses: inference
    (dinov2salad_finetuning.py:147, dinov2salad_validation.py:81: forward-only kernels) and the training loop
    ``preds = model(inputs); loss = loss_fn(preds, targets); loss.backward()`` (:120-124), where ``preds`` carries an
    autograd edge whose backward hands ``dL/dpreds`` to the fused kernels.  ``model.loss(inputs, targets)`` is the
    faster spelling of those three lines for the two losses the reference uses (MSE / Huber): regressor forward,
    loss and every regressor gradient come out of ONE library call."""

    def __init__(self, base_model: nn.Module, in_dim: int = 8448, hidden: int = 512, dropout_prob: float = 0.0,
                 loss: str = "mse", huber_delta: float = 1.0):
        super().__init__()
        self.feature_extractor = base_model
        for param in self.feature_extractor.parameters():
            param.requires_grad = False
        layers = [nn.Linear(in_dim, hidden), nn.ReLU()]
        if dropout_prob > 0:
            layers.append(nn.Dropout(dropout_prob))       # dinov2salad_finetuning_2.py:118-123 key layout (0, 3)
        layers.append(nn.Linear(hidden, 2))
        self.regressor = nn.Sequential(*layers)
        if loss not in _LATLON:
            raise ValueError("loss must be 'mse' or 'huber'")
        self.in_dim, self.hidden, self.dropout_prob = in_dim, hidden, float(dropout_prob)
        self.loss_kind, self.huber_delta = loss, float(huber_delta)
        self.seed, self._step = 0xA17, 0

    def _params(self):
        last = self.regressor[-1]
        return [self.regressor[0].weight, self.regressor[0].bias, last.weight, last.bias, None, None]

    def _cfg(self, train: bool):
        seed = (self.seed * 1000003 + self._step) & 0xFFFFFFFFFFFFFFFF
        if train and self.dropout_prob > 0:
            self._step += 1
        return make_cfg(self.hidden, _LATLON[self.loss_kind], L.ANGLE_NONE, self.huber_delta,
                        reg_dropout_p=self.dropout_prob if train else 0.0, seed=seed)

    def features(self, x) -> torch.Tensor:
        with torch.no_grad():
            return self.feature_extractor(x)              # shape: (B, 8448)

    def forward(self, x) -> torch.Tensor:
        feats = self.features(x).contiguous()
        L.require_cuda(feats, "descriptor")
        params = [p.contiguous() if p is not None else None for p in self._params()]
        if torch.is_grad_enabled() and any(p is not None and p.requires_grad for p in params):
            return _RegressorFn.apply(feats, self._cfg(self.training), *params)      # differentiable w.r.t. the regressor
        z = torch.zeros(feats.shape[0], 2, dtype=torch.float32, device=feats.device)
        with torch.no_grad():
            preds, *_ = _run(feats, params, z, None, self._cfg(self.training), False, forward_only=True)
        return preds

    def loss(self, x, targets: torch.Tensor, global_B: int = 0) -> torch.Tensor:
        """Mean loss over B*2 elements (``nn.MSELoss()`` / ``nn.HuberLoss(delta)``); ``.backward()`` fills the
        regressor gradients computed by the same call."""
        feats = self.features(x).contiguous()
        L.require_cuda(feats, "descriptor")
        tg = targets.to(device=feats.device, dtype=torch.float32).contiguous()
        cfg = self._cfg(self.training)
        cfg["global_B"] = int(global_B)
        params = [p.contiguous() if p is not None else None for p in self._params()]
        total, _losses, _preds, _sc = _HeadsLossFn.apply(feats, tg, None, cfg, None, False, *params)
        return total


# ---------------------------------------------------------------------------------------------------------------
# checkpoints
# ---------------------------------------------------------------------------------------------------------------
def _unwrap(obj) -> Mapping[str, torch.Tensor]:
    if isinstance(obj, (str, bytes)) or hasattr(obj, "__fspath__"):
        obj = torch.load(obj, map_location="cpu", weights_only=False)
    for key in ("model_state_dict", "state_dict"):           # reference checkpoints / Lightning checkpoints
        if isinstance(obj, Mapping) and key in obj and isinstance(obj[key], Mapping):
            return obj[key]
    return obj


def _load_prefixed(module: nn.Module, sd: Mapping[str, torch.Tensor], prefix: str, strict: bool):
    sub = {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}
    if not sub:
        if strict:
            raise KeyError(f"checkpoint has no '{prefix}*' entries")
        return []
    own = module.state_dict()
    fixed = {}
    for k, v in sub.items():
        if k in own and v.dim() == 2 and own[k].dim() == 4:   # Linear-shaped weights for the 1x1 convolutions
            v = v.reshape(own[k].shape)
        fixed[k] = v
    res = module.load_state_dict(fixed, strict=strict)
    return list(res.missing_keys) + list(res.unexpected_keys)


def load_upstream_salad_weights(head: SaladHead, ckpt, strict: bool = True):
    """Load the ``aggregator.*`` entries of [UPSTREAM] ``dino_salad.ckpt`` (or any state dict that holds them,
    e.g. under ``feature_extractor.aggregator.``) into a ``SaladHead``.  Returns the keys that did not match."""
    sd = _unwrap(ckpt)
    for prefix in ("feature_extractor.aggregator.", "aggregator.", "model.aggregator."):
        if any(k.startswith(prefix) for k in sd):
            return _load_prefixed(head, sd, prefix, strict)
    if strict:
        raise KeyError("no aggregator.* entries in the checkpoint")
    return []


def load_reference_checkpoint(model: DINOv2RegressionModel, ckpt, strict_backbone: bool = False) -> dict:
    """Load a reference ``checkpoint_*.pth`` (dinov2salad_finetuning.py:130-135).  The aggregator and regressor
    entries must match exactly; backbone entries are loaded when the model's backbone has the same keys and are
    reported as skipped otherwise (the backbone itself is outside this path)."""
    raw = ckpt
    if isinstance(raw, (str, bytes)) or hasattr(raw, "__fspath__"):
        raw = torch.load(raw, map_location="cpu", weights_only=False)
    sd = _unwrap(raw)
    fe = model.feature_extractor
    info = {"epoch": raw.get("epoch") if isinstance(raw, Mapping) else None, "skipped": []}
    if isinstance(fe, SaladHead):
        _load_prefixed(fe, sd, "feature_extractor.aggregator.", True)
    else:
        agg = getattr(fe, "aggregator", None)
        if agg is None:
            raise RuntimeError("model.feature_extractor has no aggregator to load into")
        _load_prefixed(agg, sd, "feature_extractor.aggregator.", True)
        bb = getattr(fe, "backbone", None)
        bb_keys = [k for k in sd if k.startswith("feature_extractor.backbone.")]
        if bb is not None and bb_keys:
            try:
                _load_prefixed(bb, sd, "feature_extractor.backbone.", True)
            except (RuntimeError, KeyError):
                if strict_backbone:
                    raise
                info["skipped"] = bb_keys
        else:
            info["skipped"] = bb_keys
    _load_prefixed(model.regressor, sd, "regressor.", True)
    return info


def save_reference_checkpoint(path, model: nn.Module, optimizer=None, epoch: int = 0, loss=None) -> None:
    """Write the dict layout of dinov2salad_finetuning.py:130-135 so the reference's validation script can read it."""
    torch.save({"epoch": epoch, "model_state_dict": model.state_dict(),
                "optimizer_state_dict": optimizer.state_dict() if optimizer is not None else {}, "loss": loss}, path)
