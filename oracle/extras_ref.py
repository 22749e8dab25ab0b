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


# ----------------------------------------------------------------------------------------------------------------
# preprocessing: transforms.Resize((224,224)) -> ToTensor() -> Normalize(mean, std) on a PIL image,
# dinov2salad/dinov2salad_finetuning.py:45-50.  torchvision delegates the resize to Pillow; what follows restates
# [THIRD-PARTY] Pillow src/libImaging/Resample.c (8-bit path, bilinear filter).  **Pinned**: tests/golden/
# preprocess_golden.npz holds the outputs of the reference's own torchvision/Pillow pipeline.
# ----------------------------------------------------------------------------------------------------------------
PRECISION_BITS = 32 - 8 - 2


def _pil_coeffs(in_size: int, out_size: int):
    scale = float(in_size) / out_size
    filterscale = max(scale, 1.0)
    support = 1.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds, coefs = [], np.zeros((out_size, ksize), dtype=np.int64)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = 0.0 + (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        w = np.array([max(0.0, 1.0 - abs((x + xmin - center + 0.5) * ss)) for x in range(xmax)], dtype=np.float64)
        ww = 0.0
        for v in w:
            ww += v
        if ww != 0.0:
            w = w / ww
        coefs[xx, :xmax] = [int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS)) for v in w]
        bounds.append((xmin, xmax))
    return bounds, coefs


def _clip8(v):
    return np.clip(v >> PRECISION_BITS, 0, 255)


def pil_resize_ref(img: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """img [H, W, 3] uint8 -> [out_h, out_w, 3] uint8, Pillow's bilinear 8-bit resampler (horizontal pass first,
    rounded to uint8, then vertical)."""
    H, W, _ = img.shape
    src = img.astype(np.int64)
    bh, ch = _pil_coeffs(W, out_w)
    tmp = np.empty((H, out_w, 3), dtype=np.int64)
    for xx, (xmin, cnt) in enumerate(bh):
        acc = (src[:, xmin:xmin + cnt, :] * ch[xx, :cnt, None]).sum(axis=1) + (1 << (PRECISION_BITS - 1))
        tmp[:, xx, :] = _clip8(acc)
    bv, cv = _pil_coeffs(H, out_h)
    out = np.empty((out_h, out_w, 3), dtype=np.int64)
    for yy, (ymin, cnt) in enumerate(bv):
        acc = (tmp[ymin:ymin + cnt] * cv[yy, :cnt, None, None]).sum(axis=0) + (1 << (PRECISION_BITS - 1))
        out[yy] = _clip8(acc)
    return out.astype(np.uint8)


def preprocess_ref(img: np.ndarray, out_h: int = 224, out_w: int = 224, mean=(0.485, 0.456, 0.406),
                   std=(0.229, 0.224, 0.225)) -> torch.Tensor:
    """[H, W, 3] uint8 -> [3, out_h, out_w] float32 (ToTensor: /255 in float32; Normalize: sub then div, float32)."""
    u8 = torch.from_numpy(pil_resize_ref(img, out_h, out_w)).permute(2, 0, 1).contiguous()
    x = u8.to(torch.float32).div(255)
    m = torch.as_tensor(mean, dtype=torch.float32)[:, None, None]
    s = torch.as_tensor(std, dtype=torch.float32)[:, None, None]
    return x.sub_(m).div_(s)
