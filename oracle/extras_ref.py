"""CPU oracle for the rows next to the hot path: the backbone->head seam, the
validation metrics and the lat/lon scaler.

TEST INFRASTRUCTURE ONLY (same rule as oracle/salad_ref.py): only tests/, __graft_entry__.smoke() and
bench.py's CPU legs may import this module.

Parity status
-------------
* seam: [UPSTREAM] DINOv2 ``forward_features`` (``x_norm = self.norm(x)``; ``x_norm[:, 0]``;
  ``x_norm[:, 1 + num_register_tokens:]``) is not vendored; restated with ``torch.nn.functional.layer_norm``.
  **Pinned** against HuggingFace ``Dinov2Model`` (the backbone SURVEY 8c names for config 1):
  tests/golden/seam_golden.npz holds ``hidden_states[-1]`` and ``last_hidden_state`` of a random-init model.
* lat/lon metrics: **pinned** -- tests/golden/metrics_golden.npz is produced by the reference's own expressions
  (sklearn ``StandardScaler`` / ``mean_squared_error`` / ``mean_absolute_error`` as called at
  dinov2salad/dinov2salad_finetuning_2.py:202-215, numpy as written at dinov2salad_validation.py:84-101).
* MAAE: **pinned** -- angle_prediction/efficient_net/validation_script.py:180-202 evaluated as written, and
  dino_v2_gemini.py's own functions (heads_golden.npz).
* recall@k: no reference code; definition restated here.  **parity unpinned**.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F


def seam_ref(hidden: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, eps: float = 1e-6, skip: int = 0):
    """(x_tok [B,n,C], t [B,C]) = split(LayerNorm(hidden)); [UPSTREAM] dinov2 forward_features."""
    x = F.layer_norm(hidden.float(), (hidden.shape[-1],), weight.float(), bias.float(), eps)
    return x[:, 1 + skip:], x[:, 0]


def latlon_metrics_ref(pred_scaled, target, mean=None, scale=None, targets_scaled=False) -> dict:
    """dinov2_validation.py:84-101 and dinov2salad_finetuning_2.py:202-215 in float64."""
    p = np.asarray(pred_scaled, dtype=np.float64)
    t = np.asarray(target, dtype=np.float64)
    if mean is not None:
        p = p * np.asarray(scale, dtype=np.float64) + np.asarray(mean, dtype=np.float64)
        if targets_scaled:
            t = t * np.asarray(scale, dtype=np.float64) + np.asarray(mean, dtype=np.float64)
    d = p - t
    n = len(p)
    mse = float((d ** 2).mean())
    return {"n": n, "final_loss": float(0.5 * ((d[:, 0] ** 2).sum() + (d[:, 1] ** 2).sum()) / n), "mse": mse,
            "rmse": math.sqrt(mse), "mae": float(np.abs(d).mean()), "mae_lat": float(np.abs(d[:, 0]).mean()),
            "mae_lon": float(np.abs(d[:, 1]).mean())}


def maae_ref(sincos_pred: torch.Tensor, true_deg: torch.Tensor, order: str = "sin_cos"):
    p = sincos_pred.float()
    t = true_deg.float()
    if order == "cos_sin":          # angle_prediction/efficient_net/validation_script.py:180-189
        deg = (torch.rad2deg(torch.atan2(p[:, 1], p[:, 0])) + 360.0) % 360.0
    else:                           # angle_prediction/dinov2salad/dino_v2_gemini.py:117-141
        deg = torch.rad2deg(torch.atan2(p[:, 0], p[:, 1])) % 360.0
        t = t % 360
    diff = torch.abs(deg - t)
    err = torch.minimum(diff, 360.0 - diff)
    return float(err.double().sum() / len(p)), deg


def recall_at_k_ref(idx: torch.Tensor, query_labels: torch.Tensor, gallery_labels: torch.Tensor, ks=(1, 5, 10, 25)) -> dict:
    """Fraction of queries with at least one same-label gallery row among their first k neighbours."""
    hit = gallery_labels[idx] == query_labels[:, None]            # [Q, k]
    return {int(k): float(hit[:, :k].any(dim=1).double().mean()) for k in ks}
