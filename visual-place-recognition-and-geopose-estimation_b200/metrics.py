"""Validation metrics on the device (vpr_metrics_latlon / vpr_metrics_angle / vpr_recall_at_k) and the
lat/lon z-score scaler of the reference.

Reference lines: the scaler is ``StandardScaler().fit(train[['latitude','longitude']])`` saved with joblib
(dinov2salad/dinov2salad_finetuning.py:79-81,110-111) and loaded for validation
(dinov2salad/dinov2salad_validation.py:73); the metrics are ``final_loss`` (dinov2salad_validation.py:101),
MSE/RMSE/MAE (dinov2salad_finetuning_2.py:211-215) and MAAE (angle_prediction/efficient_net/
validation_script.py:180-202, angle_prediction/dinov2salad/dino_v2_gemini.py:117-141).
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Sequence

import numpy as np
import torch

from . import _lib as L


class LatLonScaler:
    """Two-column StandardScaler (mean_, scale_) without the sklearn dependency at run time."""

    def __init__(self, mean, scale):
        self.mean_ = np.asarray(mean, dtype=np.float64).reshape(2)
        self.scale_ = np.asarray(scale, dtype=np.float64).reshape(2)
        self._dev = {}

    @classmethod
    def fit(cls, latlon) -> "LatLonScaler":
        x = np.asarray(latlon, dtype=np.float64).reshape(-1, 2)
        scale = x.std(axis=0)                       # population std, as StandardScaler
        scale[scale == 0.0] = 1.0
        return cls(x.mean(axis=0), scale)

    @classmethod
    def from_sklearn(cls, scaler) -> "LatLonScaler":
        return cls(scaler.mean_, scaler.scale_)

    @classmethod
    def from_joblib(cls, path: str) -> "LatLonScaler":
        import joblib
        return cls.from_sklearn(joblib.load(path))

    def to_sklearn(self):
        from sklearn.preprocessing import StandardScaler
        s = StandardScaler()
        s.mean_, s.scale_, s.var_ = self.mean_.copy(), self.scale_.copy(), self.scale_ ** 2
        s.n_features_in_, s.n_samples_seen_ = 2, 1
        return s

    def transform(self, latlon):
        return (np.asarray(latlon, dtype=np.float64) - self.mean_) / self.scale_

    def inverse_transform(self, z):
        if isinstance(z, torch.Tensor):
            m, s = self.device_tensors(z.device)
            return z.double() * s + m
        return np.asarray(z, dtype=np.float64) * self.scale_ + self.mean_

    def device_tensors(self, device):
        key = str(device)
        if key not in self._dev:
            self._dev[key] = (torch.tensor(self.mean_, dtype=torch.float64, device=device),
                              torch.tensor(self.scale_, dtype=torch.float64, device=device))
        return self._dev[key]


def latlon_metrics(pred: torch.Tensor, target: torch.Tensor, scaler: LatLonScaler | None = None,
                   targets_scaled: bool = False) -> dict:
    """pred [N,2] z-scored predictions, target [N,2]; returns final_loss, mse, rmse, mae, mae_lat, mae_lon
    in original units."""
    L.require_cuda(pred, "pred")
    lib = L.lib()
    pred = pred.float().contiguous()
    target = target.to(device=pred.device, dtype=torch.float32).contiguous()
    if pred.dim() != 2 or pred.shape[1] != 2 or target.shape != pred.shape:
        raise ValueError("pred and target must both be [N, 2]")
    N = pred.shape[0]
    if N == 0:
        return {"n": 0}
    m, s = scaler.device_tensors(pred.device) if scaler is not None else (None, None)
    out = torch.empty(4, dtype=torch.float64, device=pred.device)
    with L.on_device(pred):
        L.check(lib.vpr_metrics_latlon(L.ptr(pred), L.ptr(target), L.ptr(m), L.ptr(s), int(targets_scaled), N, L.ptr(out),
                                       L.stream_ptr(pred.device)), "vpr_metrics_latlon")
    s0, s1, a0, a1 = out.tolist()
    mse = (s0 + s1) / (2 * N)
    return {"n": N, "final_loss": 0.5 * (s0 + s1) / N, "mse": mse, "rmse": math.sqrt(mse), "mae": (a0 + a1) / (2 * N),
            "mae_lat": a0 / N, "mae_lon": a1 / N}


def maae(sincos_pred: torch.Tensor, true_deg: torch.Tensor, order: str = "sin_cos"):
    """Mean absolute angular error in degrees and the predicted angles.  ``order``: "sin_cos"
    (dino_v2_gemini.py head) or "cos_sin" (efficient_net/validation_script.py)."""
    if order not in ("sin_cos", "cos_sin"):
        raise ValueError("order must be 'sin_cos' or 'cos_sin'")
    L.require_cuda(sincos_pred, "sincos_pred")
    lib = L.lib()
    p = sincos_pred.float().contiguous()
    t = true_deg.to(device=p.device, dtype=torch.float32).contiguous()
    if p.dim() != 2 or p.shape[1] != 2 or t.shape != (p.shape[0],):
        raise ValueError("sincos_pred must be [N, 2] and true_deg [N]")
    N = p.shape[0]
    if N == 0:
        return float("nan"), torch.empty(0, dtype=torch.float32, device=p.device)
    pred_deg = torch.empty(N, dtype=torch.float32, device=p.device)
    out = torch.empty(1, dtype=torch.float64, device=p.device)
    with L.on_device(p):
        L.check(lib.vpr_metrics_angle(L.ptr(p), L.ptr(t), int(order == "cos_sin"), N, L.ptr(pred_deg), L.ptr(out),
                                      L.stream_ptr(p.device)), "vpr_metrics_angle")
    return out.item() / N, pred_deg


def recall_at_k(idx: torch.Tensor, query_labels: torch.Tensor, gallery_labels: torch.Tensor,
                ks: Sequence[int] = (1, 5, 10, 25)) -> dict:
    """idx [Q,k] from knn_topk; a query counts at k if one of its first k neighbours has its label."""
    L.require_cuda(idx, "idx")
    lib = L.lib()
    idx = idx.to(torch.int64).contiguous()
    Q, k = idx.shape
    ks = [int(v) for v in ks]
    if not ks or len(ks) > 16 or min(ks) < 1:
        raise ValueError("ks must hold 1..16 positive cut-offs")
    if max(ks) > k:
        raise ValueError(f"recall@{max(ks)} needs at least {max(ks)} neighbours per query, idx has {k}")
    ql = query_labels.to(device=idx.device, dtype=torch.int64).contiguous()
    gl = gallery_labels.to(device=idx.device, dtype=torch.int64).contiguous()
    if ql.shape != (Q,):
        raise ValueError("query_labels must be [Q]")
    if Q == 0:
        return {kk: float("nan") for kk in ks}
    ks_dev = torch.tensor(ks, dtype=torch.int32, device=idx.device)
    hits = torch.empty(len(ks), dtype=torch.int64, device=idx.device)      # uint64 counts < 2^63
    with L.on_device(idx):
        L.check(lib.vpr_recall_at_k(L.ptr(idx), L.ptr(gl), L.ptr(ql), Q, k, gl.numel(), L.ptr(ks_dev), len(ks), L.ptr(hits),
                                    L.stream_ptr(idx.device)), "vpr_recall_at_k")
    h = hits.tolist()
    return {kk: hh / Q for kk, hh in zip(ks, h)}


# ---------------------------------------------------------------------------------------------------------------
# prediction CSV files (host side; the last step of the reference's validation scripts)
# ---------------------------------------------------------------------------------------------------------------
def _host_rows(a, cols):
    if isinstance(a, torch.Tensor):
        a = a.detach().cpu().double().numpy()
    a = np.asarray(a, dtype=np.float64)
    return a.reshape(-1, cols) if cols > 1 else a.reshape(-1)


def write_latlon_predictions_csv(path, filenames, true_latlon, pred_latlon) -> int:
    """The per-image prediction file of swin_transformer/validation_swin_2.py:205-218 (same writer in
    val_and_test_swin_2.py:280-290): columns ``filename,true_latitude,true_longitude,predicted_latitude,
    predicted_longitude,error_latitude,error_longitude``, the errors being ``np.abs(true - predicted)``, written
    like ``DataFrame.to_csv(index=False, float_format='%.6f')``.  Values are in original units (run predictions
    through ``LatLonScaler.inverse_transform`` first, as the reference does at :176).  Returns the row count."""
    import csv
    t, p = _host_rows(true_latlon, 2), _host_rows(pred_latlon, 2)
    names = list(filenames)
    if not (len(names) == len(t) == len(p)):
        raise ValueError("filenames, true_latlon and pred_latlon must have one row per image")
    err = np.abs(t - p)
    with open(path, "w", newline="") as f:
        w = csv.writer(f, lineterminator="\n")
        w.writerow(["filename", "true_latitude", "true_longitude", "predicted_latitude", "predicted_longitude",
                    "error_latitude", "error_longitude"])
        for i, name in enumerate(names):
            w.writerow([name] + ["%.6f" % v for v in (t[i, 0], t[i, 1], p[i, 0], p[i, 1], err[i, 0], err[i, 1])])
    return len(names)


def write_angle_predictions_csv(path, filenames, true_angle, pred_angle) -> int:
    """The per-image file of angle_prediction/efficient_net/validation_script.py:211-220: columns
    ``filename,true_angle,predicted_angle,angular_error`` with ``angular_error = min(|p - t|, 360 - |p - t|)``,
    written like ``DataFrame.to_csv(index=False)`` (shortest round-trip float text; integer-valued targets that arrive
    as integers stay integers)."""
    import csv
    names = list(filenames)
    tl = true_angle.detach().cpu().tolist() if isinstance(true_angle, torch.Tensor) else list(true_angle)
    pl = pred_angle.detach().cpu().tolist() if isinstance(pred_angle, torch.Tensor) else list(pred_angle)
    if not (len(names) == len(tl) == len(pl)):
        raise ValueError("filenames, true_angle and pred_angle must have one entry per image")

    def txt(v):
        if isinstance(v, (int, np.integer)):
            return str(int(v))
        return repr(float(v))

    with open(path, "w", newline="") as f:
        w = csv.writer(f, lineterminator="\n")
        w.writerow(["filename", "true_angle", "predicted_angle", "angular_error"])
        for name, t, p in zip(names, tl, pl):
            d = abs(p - t)
            w.writerow([name, txt(t), txt(p), txt(min(d, 360 - d))])
    return len(names)
