"""CPU oracle for the SALAD head, geopose heads/losses and descriptor kNN.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package may import this
module: only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs do, and there only as the checker
(or as the timed CPU baseline), never as the thing shipped.

Parity status
-------------
* SALAD aggregator (``SaladRef``, ``get_matching_probs``, ``log_otp_solver``):
  **parity unpinned**.  The reference never contains this arithmetic; it pulls
  it at run time with ``torch.hub.load("serizba/salad", "dinov2_salad")``
  (reference ``dinov2salad/dinov2salad_finetuning.py:19,92`` and
  ``dinov2salad/dinov2salad_validation.py:65``), without a version pin, and the
  hub repo is not vendored under /root/reference and cannot be fetched (no
  network).  What follows restates the published algorithm of
  ``serizba/salad  models/aggregators/salad.py``.  The
  only facts the reference itself pins are the output width 8448
  (``dinov2salad_finetuning.py:29,36``) and the ImageNet-normalised 224x224
  input (``:45-50``); both are asserted in tests/test_oracle.py.
* Geopose heads and losses: **pinned**.  tests/golden/make_golden.py extracts
  the reference's own loss/metric functions from the reference source files
  (AST extraction, no training code executed) and stores their outputs on
  seeded inputs in tests/golden/heads_golden.npz; tests/test_oracle.py checks
  every function below against those vectors.
* kNN: no reference code exists (finding 4); the oracle is
  ``topk(Q @ G.T)`` on identical (bf16-rounded) inputs.  **parity unpinned**.
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

# ----------------------------------------------------------------------------
# Sinkhorn (upstream serizba/salad  models/aggregators/salad.py: log_otp_solver,
# get_matching_probs)
# ----------------------------------------------------------------------------


def log_otp_solver(log_a: torch.Tensor, log_b: torch.Tensor, M: torch.Tensor,
                   num_iters: int = 3, reg: float = 1.0) -> torch.Tensor:
    """Log-domain Sinkhorn.  M: [B, m+1, n]; log_a: [B, m+1]; log_b: [B, n].

    u then v, v updated last; returns M/reg + u + v.
    """
    M = M / reg
    u = torch.zeros_like(log_a)
    v = torch.zeros_like(log_b)
    for _ in range(num_iters):
        u = log_a - torch.logsumexp(M + v.unsqueeze(1), dim=2)
        v = log_b - torch.logsumexp(M + u.unsqueeze(2), dim=1)
    return M + u.unsqueeze(2) + v.unsqueeze(1)


def get_matching_probs(S: torch.Tensor, dustbin_score, num_iters: int = 3,
                       reg: float = 1.0, const_dtype: torch.dtype = torch.float32) -> torch.Tensor:
    """S: [B, m, n] scores -> [B, m+1, n] log-probabilities (dustbin row last).

    ``norm``, ``log_a`` and ``log_b`` are built from ``math.log`` in fp32 upstream
    whatever the dtype of ``S``; ``const_dtype=torch.float64``
    gives the deliberate fp64 variant used by the gradcheck tests.
    """
    B, m, n = S.shape
    S_aug = S.new_empty(B, m + 1, n)
    S_aug[:, :m, :] = S
    S_aug[:, m, :] = dustbin_score
    norm = -torch.tensor(math.log(n + m), dtype=const_dtype, device=S.device)
    log_a = norm.expand(m + 1).contiguous()
    log_b = norm.expand(n).contiguous()
    log_a[-1] = log_a[-1] + math.log(n - m)
    log_a = log_a.expand(B, -1)
    log_b = log_b.expand(B, -1)
    log_P = log_otp_solver(log_a, log_b, S_aug, num_iters=num_iters, reg=reg)
    return log_P - norm


# ----------------------------------------------------------------------------
# SALAD aggregator (upstream SALAD.forward)
# ----------------------------------------------------------------------------


class SaladRef(nn.Module):
    """Restatement of upstream ``SALAD`` with identical parameter names/shapes.

    The hub entry the reference calls (``dinov2salad_finetuning.py:19``) builds
    it with num_channels=768, num_clusters=64, cluster_dim=128, token_dim=256.
    ``drop_mask`` ([B, 1024, h, w] of {0,1}, cluster half first then score half)
    replaces the RNG of the two Dropout layers so that training mode can be
    compared against the CUDA path, whose RNG stream cannot match torch's.
    """

    def __init__(self, num_channels=768, num_clusters=64, cluster_dim=128,
                 token_dim=256, dropout=0.3, aggregate="repeat"):
        super().__init__()
        self.num_channels = num_channels
        self.num_clusters = num_clusters
        self.cluster_dim = cluster_dim
        self.token_dim = token_dim
        self.p = float(dropout)
        self.aggregate = aggregate
        drop = nn.Dropout(dropout) if dropout > 0 else nn.Identity()
        self.token_features = nn.Sequential(
            nn.Linear(num_channels, 512), nn.ReLU(), nn.Linear(512, token_dim))
        self.cluster_features = nn.Sequential(
            nn.Conv2d(num_channels, 512, 1), drop, nn.ReLU(), nn.Conv2d(512, cluster_dim, 1))
        self.score = nn.Sequential(
            nn.Conv2d(num_channels, 512, 1), drop, nn.ReLU(), nn.Conv2d(512, num_clusters, 1))
        self.dust_bin = nn.Parameter(torch.tensor(1.0))

    def _mlp(self, seq: nn.Sequential, x: torch.Tensor, mask: Optional[torch.Tensor]):
        if mask is None:
            return seq(x)
        h = seq[0](x)
        h = h * mask.to(h.dtype) / (1.0 - self.p)
        return seq[3](seq[2](h))

    def forward(self, x: Tuple[torch.Tensor, torch.Tensor],
                drop_mask: Optional[torch.Tensor] = None, const_dtype=torch.float32):
        x, t = x
        mc = ms = None
        if drop_mask is not None:
            mc, ms = drop_mask[:, :512], drop_mask[:, 512:]
        f = self._mlp(self.cluster_features, x, mc).flatten(2)   # [B, l, n]
        p = self._mlp(self.score, x, ms).flatten(2)              # [B, m, n]
        t = self.token_features(t)                               # [B, g]
        p = get_matching_probs(p, self.dust_bin, 3, const_dtype=const_dtype)
        p = torch.exp(p)[:, :-1, :]
        if self.aggregate == "repeat":      # what upstream launches (two 8.4 MB/img temporaries)
            pr = p.unsqueeze(1).repeat(1, self.cluster_dim, 1, 1)
            fr = f.unsqueeze(2).repeat(1, 1, self.num_clusters, 1)
            agg = (fr * pr).sum(dim=-1)
        else:                               # same contraction without the temporaries
            agg = torch.einsum("bln,bmn->blm", f, p)
        out = torch.cat([F.normalize(t, p=2, dim=-1),
                         F.normalize(agg, p=2, dim=1).flatten(1)], dim=-1)
        return F.normalize(out, p=2, dim=-1)


def tokens_to_salad_input(patch_tokens: torch.Tensor, h: int, w: int) -> torch.Tensor:
    """[B, n, C] token-major -> the [B, C, h, w] permuted view upstream's backbone
    wrapper hands to SALAD.  No copy."""
    B, n, C = patch_tokens.shape
    assert n == h * w
    return patch_tokens.reshape(B, h, w, C).permute(0, 3, 1, 2)


# ----------------------------------------------------------------------------
# Geopose heads
# ----------------------------------------------------------------------------


class LatLonRegressorRef(nn.Module):
    """``Linear(D,512)-ReLU-[Dropout]-Linear(512,2)``.

    reference dinov2salad/dinov2salad_finetuning.py:28-32 (no dropout, D=8448);
    dinov2salad/dinov2salad_finetuning_2.py:118-123 (Dropout(0.3) variant).
    Parameter names ``regressor.{0,2|3}.*`` follow the reference's Sequential.
    """

    def __init__(self, in_dim=8448, hidden=512, dropout=0.0):
        super().__init__()
        layers = [nn.Linear(in_dim, hidden), nn.ReLU()]
        if dropout > 0:
            layers.append(nn.Dropout(dropout))
        layers.append(nn.Linear(hidden, 2))
        self.regressor = nn.Sequential(*layers)

    def forward(self, d):
        return self.regressor(d)


class SinCosHeadRef(nn.Module):
    """``[Dropout] -> Linear(D,2)`` = [sin, cos].

    reference angle_prediction/dinov2salad/dino_v2_gemini.py:104-114 (raw output),
    angle_prediction/swin/swin_angle_finetuning_sin_cos.py:56-62
    (``F.normalize(out, dim=1, eps=1e-6)`` applied when ``unit_norm``).
    """

    def __init__(self, in_dim=8448, unit_norm=False):
        super().__init__()
        self.head = nn.Linear(in_dim, 2)
        self.unit_norm = unit_norm

    def forward(self, d):
        out = self.head(d)
        if self.unit_norm:
            out = F.normalize(out, dim=1, p=2, eps=1e-6)
        return out


class ScalarAngleHeadRef(nn.Module):
    """``Linear(D,1)`` -> ``squeeze(-1) % 360``: the scalar-angle head of
    reference angle_prediction/dinov2salad/dino_angle_finetuning.py:37-47."""

    def __init__(self, in_dim=8448):
        super().__init__()
        self.head = nn.Linear(in_dim, 1)

    def forward(self, d):
        angle = self.head(d).squeeze(-1)
        return angle % 360


def mse_loss(pred, target):
    """reference dinov2salad/dinov2salad_finetuning.py:96 (nn.MSELoss, mean over B*2)."""
    return F.mse_loss(pred, target)


def huber_loss(pred, target, delta=1.0):
    """reference dinov2salad/dinov2salad_finetuning_2.py:154 (nn.HuberLoss(delta))."""
    return F.huber_loss(pred, target, delta=delta)


def sincos_to_deg(sin_cos):
    """reference angle_prediction/dinov2salad/dino_v2_gemini.py:135-141."""
    return torch.rad2deg(torch.atan2(sin_cos[:, 0], sin_cos[:, 1])) % 360.0


def maae_loss_deg(pred_deg, target_deg):
    """reference angle_prediction/dinov2salad/dino_v2_gemini.py:117-127."""
    t = target_deg % 360
    p = pred_deg % 360
    d = torch.abs(p - t)
    return torch.mean(torch.minimum(d, 360.0 - d))


def maae_loss_scalar(pred, target):
    """reference angle_prediction/dinov2salad/dino_angle_finetuning.py:50-52
    (``abs(p - t) % 360`` *then* min: the modulo is applied to the difference)."""
    d = torch.abs(pred - target) % 360
    return torch.mean(torch.minimum(d, 360 - d))


def sincos_atan2_maae(sin_cos_pred, target_deg):
    """Head output -> degrees -> MAAE: dino_v2_gemini.py:243-246 composition."""
    return maae_loss_deg(sincos_to_deg(sin_cos_pred), target_deg)


def acos_angular_loss(pred_unit, target_sincos):
    """reference angle_prediction/swin/swin_angle_finetuning_sin_cos.py:65-69."""
    c = (pred_unit * target_sincos).sum(dim=1)
    c = torch.clamp(c, -0.999999, 0.999999)
    return torch.mean(torch.rad2deg(torch.acos(c)))


def angle_error_atan2(pred, target):
    """reference angle_prediction/swin/swin_angle_finetuning_sin_cos.py:72-76."""
    pa = torch.atan2(pred[:, 0], pred[:, 1])
    ta = torch.atan2(target[:, 0], target[:, 1])
    d = torch.rad2deg(torch.abs(pa - ta)) % 360
    return torch.mean(torch.minimum(d, 360 - d))


def maae_sincos_metric(pred_sc, target_sc):
    """reference angle_prediction/swin/swin_angle_finetuning_gemini.py:131-146."""
    p = (torch.rad2deg(torch.atan2(pred_sc[:, 0], pred_sc[:, 1])) + 360) % 360
    t = (torch.rad2deg(torch.atan2(target_sc[:, 0], target_sc[:, 1])) + 360) % 360
    d = torch.abs(p - t)
    return torch.mean(torch.minimum(d, 360 - d))


def deg_to_sincos(deg):
    """reference angle_prediction/swin/swin_angle_finetuning_sin_cos.py:45-47."""
    r = torch.deg2rad(deg % 360)
    return torch.stack([torch.sin(r), torch.cos(r)], dim=1)


def latlon_validation_score(pred, target):
    """reference dinov2salad/dinov2salad_validation.py:101:
    0.5 * (sum dlat^2 + sum dlon^2) / N."""
    return 0.5 * ((pred[:, 0] - target[:, 0]) ** 2).sum() / pred.shape[0] + \
        0.5 * ((pred[:, 1] - target[:, 1]) ** 2).sum() / pred.shape[0]


LATLON_LOSSES = ("mse", "huber")
ANGLE_LOSSES = ("atan2_maae", "acos", "mse", "scalar_maae")


def geopose_loss(desc, latlon_head: LatLonRegressorRef, angle_head,
                 latlon_target, angle_target_deg, latlon_loss="mse", angle_loss="atan2_maae",
                 head_keep_mask=None, head_dropout_p=0.0, reg_keep_mask=None, reg_dropout_p=0.0):
    """One fine-tuning objective over a descriptor batch: lat/lon + angle head.

    Returns (total, latlon_loss, angle_loss, latlon_pred, sincos_pred).  The total
    is the plain sum; the reference trains the two tasks in separate scripts, so
    summing is this repo's (documented) choice for BASELINE.json config 3.

    ``head_keep_mask`` [B, D] in {0,1} with ``head_dropout_p``: the ``nn.Dropout`` in front of
    the angle head (dino_v2_gemini.py:104,111) evaluated with an injected mask (inverted
    dropout: kept elements are scaled by 1/(1-p)); ``reg_keep_mask`` [B, hidden] likewise for
    the regressor's Dropout (dinov2salad_finetuning_2.py:121).
    """
    if reg_keep_mask is None:
        ll = latlon_head(desc)
    else:
        reg = latlon_head.regressor
        z = torch.relu(reg[0](desc)) * reg_keep_mask / (1.0 - reg_dropout_p)
        ll = reg[-1](z)
    if latlon_loss == "mse":
        l1 = mse_loss(ll, latlon_target)
    elif latlon_loss == "huber":
        l1 = huber_loss(ll, latlon_target, 1.0)
    else:
        raise ValueError(latlon_loss)
    hd = desc if head_keep_mask is None else desc * head_keep_mask / (1.0 - head_dropout_p)
    sc = angle_head.head(hd)
    if angle_loss == "scalar_maae":
        pred = sc.squeeze(-1) % 360                      # dino_angle_finetuning.py:46-47
        l2 = maae_loss_scalar(pred, angle_target_deg)
        return l1 + l2, l1, l2, ll, pred
    if angle_loss == "atan2_maae":
        l2 = sincos_atan2_maae(sc, angle_target_deg)
    elif angle_loss == "acos":
        l2 = acos_angular_loss(F.normalize(sc, dim=1, p=2, eps=1e-6), deg_to_sincos(angle_target_deg))
    elif angle_loss == "mse":
        l2 = mse_loss(sc, deg_to_sincos(angle_target_deg))
    else:
        raise ValueError(angle_loss)
    return l1 + l2, l1, l2, ll, sc


# ----------------------------------------------------------------------------
# kNN (no reference code)
# ----------------------------------------------------------------------------


def knn_topk_ref(q: torch.Tensor, g: torch.Tensor, k: int = 25, g_offset: int = 0,
                 chunk: int = 65536):
    """topk(Q @ G.T) with fp32 accumulation on the given (already bf16-rounded)
    rows.  Ties broken towards the lower gallery index.  Returns (score fp32
    [Q,k], idx int64 [Q,k]) sorted by descending score."""
    qf = q.float()
    best_s = torch.full((q.shape[0], 0), 0.0)
    best_i = torch.zeros((q.shape[0], 0), dtype=torch.int64)
    for s0 in range(0, g.shape[0], chunk):
        sc = qf @ g[s0:s0 + chunk].float().T
        idx = torch.arange(s0, s0 + sc.shape[1]).expand_as(sc) + g_offset
        best_s = torch.cat([best_s, sc], dim=1)
        best_i = torch.cat([best_i, idx], dim=1)
        best_s, best_i = merge_topk(best_s, best_i, k)
    return best_s, best_i


def merge_topk(scores: torch.Tensor, idx: torch.Tensor, k: int):
    """Select the k best (score desc, index asc) columns of each row."""
    kk = min(k, scores.shape[1])
    # stable lexicographic order: sort by index first, then stable by -score
    o1 = torch.argsort(idx, dim=1, stable=True)
    s1 = torch.gather(scores, 1, o1)
    i1 = torch.gather(idx, 1, o1)
    o2 = torch.argsort(-s1, dim=1, stable=True)
    s2 = torch.gather(s1, 1, o2)[:, :kk]
    i2 = torch.gather(i1, 1, o2)[:, :kk]
    return s2, i2


# ----------------------------------------------------------------------------
# Explicit backward of the Sinkhorn solver in plain
# loops of tensor ops: an independent check of what the CUDA kernel implements.
# ----------------------------------------------------------------------------


def sinkhorn_backward_explicit(S: torch.Tensor, dP: torch.Tensor, num_iters: int = 3,
                               dust_bin: float = 1.0):
    """S: [B,m,n] scores, dP: [B,m,n] gradient w.r.t. P = exp(logP)[:, :m].
    Returns dS [B,m,n].  Follows the reverse-unrolled recurrence A-3."""
    B, m, n = S.shape
    dt = S.dtype
    norm = -math.log(n + m)
    M = torch.cat([S, torch.full((B, 1, n), dust_bin, dtype=dt)], dim=1)
    log_a = torch.full((B, m + 1), norm, dtype=dt)
    log_a[:, m] += math.log(n - m)
    log_b = torch.full((B, n), norm, dtype=dt)
    us, vs = [], [torch.zeros(B, n, dtype=dt)]
    for _ in range(num_iters):
        u = log_a - torch.logsumexp(M + vs[-1].unsqueeze(1), dim=2)
        v = log_b - torch.logsumexp(M + u.unsqueeze(2), dim=1)
        us.append(u)
        vs.append(v)
    logP = M + us[-1].unsqueeze(2) + vs[-1].unsqueeze(1) - norm
    G = torch.zeros_like(M)
    G[:, :m] = dP * torch.exp(logP[:, :m])
    dM = G.clone()
    du = G.sum(2)
    dv = G.sum(1)
    for k in range(num_iters, 0, -1):
        u_k, v_k, v_km1 = us[k - 1], vs[k], vs[k - 1]
        Wv = torch.exp(M + u_k.unsqueeze(2) + v_k.unsqueeze(1) - log_b.unsqueeze(1))
        dM = dM - dv.unsqueeze(1) * Wv
        du = du - (dv.unsqueeze(1) * Wv).sum(2)
        Wu = torch.exp(M + v_km1.unsqueeze(1) + u_k.unsqueeze(2) - log_a.unsqueeze(2))
        dM = dM - du.unsqueeze(2) * Wu
        dv = -(du.unsqueeze(2) * Wu).sum(1)
        du = torch.zeros_like(du)
    return dM[:, :m]
