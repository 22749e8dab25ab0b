"""Generate the committed golden vectors.  Run HERE (the build container), where
/root/reference exists; the GPU box only reads the .npz files.

    python tests/golden/make_golden.py

1. heads_golden.npz -- outputs of the REFERENCE's own loss / metric functions on
   seeded inputs.  The reference scripts run a training job at import time, so the
   functions are lifted out of their source files by AST (only the named ``def``s
   are compiled; no other reference statement runs):
     angle_prediction/dinov2salad/dino_v2_gemini.py   angular_regression_loss, prediction_to_angle_deg
     angle_prediction/dinov2salad/dino_angle_finetuning.py   angular_regression_loss (scalar variant)
     angle_prediction/swin/swin_angle_finetuning_sin_cos.py   angular_loss, compute_angle_error
     angle_prediction/swin/swin_angle_finetuning_gemini.py    mean_absolute_angular_error
   plus torch's own nn.MSELoss / nn.HuberLoss(delta=1) which the reference
   instantiates directly (dinov2salad_finetuning.py:96, dinov2salad_finetuning_2.py:154),
   and the validation score expression of dinov2salad_validation.py:101 evaluated
   with numpy exactly as written there.
2. salad_golden.npz -- outputs of the CPU oracle (oracle/salad_ref.py) on seeded
   inputs: regression fixtures for the restated SALAD (parity unpinned upstream).
"""
import ast
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F  # noqa: F401  (used by extracted functions)

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)


def lift(path, names):
    """Compile only the named top-level functions of a reference script."""
    src = open(os.path.join(REF, path)).read()
    tree = ast.parse(src)
    keep = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in names]
    assert len(keep) == len(names), (path, names, [k.name for k in keep])
    mod = ast.Module(body=keep, type_ignores=[])
    ns = {"torch": torch, "F": F, "np": np}
    exec(compile(mod, path, "exec"), ns)
    return [ns[n] for n in names]


def heads_inputs():
    g = torch.Generator().manual_seed(20261018)
    B = 37
    sc = torch.randn(B, 2, generator=g)
    sc[3] = torch.tensor([0.0, -1.0])          # atan2 branch cut neighbourhood
    sc[4] = torch.tensor([1e-4, -2.0])
    sc[5] = torch.tensor([-1e-4, -2.0])
    tgt_deg = torch.rand(B, generator=g) * 750.0   # labels go up to 750 deg (SURVEY D-8)
    pred_deg = torch.randn(B, generator=g) * 400.0
    ll_pred = torch.randn(B, 2, generator=g) * 2.0
    ll_tgt = torch.randn(B, 2, generator=g)
    return sc, tgt_deg, pred_deg, ll_pred, ll_tgt


def make_heads():
    sc, tgt_deg, pred_deg, ll_pred, ll_tgt = heads_inputs()
    (maae_deg, to_deg) = lift("angle_prediction/dinov2salad/dino_v2_gemini.py",
                              ["angular_regression_loss", "prediction_to_angle_deg"])
    (maae_scalar,) = lift("angle_prediction/dinov2salad/dino_angle_finetuning.py",
                          ["angular_regression_loss"])
    (acos_loss, angle_err) = lift("angle_prediction/swin/swin_angle_finetuning_sin_cos.py",
                                  ["angular_loss", "compute_angle_error"])
    (maae_sc,) = lift("angle_prediction/swin/swin_angle_finetuning_gemini.py",
                      ["mean_absolute_angular_error"])
    out = {}
    out["deg_from_sincos"] = to_deg(sc).numpy()
    out["maae_deg"] = maae_deg(pred_deg, tgt_deg).numpy()
    out["maae_atan2"] = maae_deg(to_deg(sc), tgt_deg).numpy()
    out["maae_scalar"] = maae_scalar(pred_deg % 360, tgt_deg).numpy()
    r = torch.deg2rad(tgt_deg % 360)
    tgt_sc = torch.stack([torch.sin(r), torch.cos(r)], dim=1)
    unit = F.normalize(sc, dim=1, p=2, eps=1e-6)     # swin_angle_finetuning_sin_cos.py:62
    out["acos_loss"] = acos_loss(unit, tgt_sc).numpy()
    out["angle_err_atan2"] = angle_err(unit, tgt_sc).numpy()
    out["maae_sincos_metric"] = maae_sc(sc, tgt_sc).numpy()
    out["mse"] = torch.nn.MSELoss()(ll_pred, ll_tgt).numpy()
    out["huber"] = torch.nn.HuberLoss(delta=1.0)(ll_pred, ll_tgt).numpy()
    out["mse_sincos"] = torch.nn.MSELoss()(sc, tgt_sc).numpy()
    # dinov2salad_validation.py:96-101, evaluated with numpy as written there
    all_preds, all_targets = ll_pred.numpy(), ll_tgt.numpy()
    out["val_score"] = np.asarray(
        0.5 * (np.sum((all_preds[:, 0] - all_targets[:, 0]) ** 2)
               + np.sum((all_preds[:, 1] - all_targets[:, 1]) ** 2)) / len(all_preds))
    # gradients of the differentiable losses w.r.t. the head output (autograd through the
    # reference functions themselves)
    for name, fn in (("maae_atan2", lambda s: maae_deg(to_deg(s), tgt_deg)),
                     ("acos_loss", lambda s: acos_loss(F.normalize(s, dim=1, p=2, eps=1e-6), tgt_sc)),
                     ("mse_sincos", lambda s: torch.nn.MSELoss()(s, tgt_sc))):
        s = sc.clone().requires_grad_(True)
        fn(s).backward()
        out["grad_" + name] = s.grad.numpy()
    # scalar-angle head (dino_angle_finetuning.py:46-52): gradient w.r.t. the raw head output through `% 360`
    raw = pred_deg.clone().requires_grad_(True)
    maae_scalar(raw % 360, tgt_deg % 360).backward()
    out["maae_scalar_wrapped_targets"] = maae_scalar(pred_deg % 360, tgt_deg % 360).numpy()
    out["grad_maae_scalar"] = raw.grad.numpy()
    for name, fn in (("mse", torch.nn.MSELoss()), ("huber", torch.nn.HuberLoss(delta=1.0))):
        p = ll_pred.clone().requires_grad_(True)
        fn(p, ll_tgt).backward()
        out["grad_" + name] = p.grad.numpy()
    np.savez(os.path.join(HERE, "heads_golden.npz"), **out)
    print("heads_golden.npz:", sorted(out))


def salad_case(B, hw, seed):
    from oracle.salad_ref import SaladRef, tokens_to_salad_input
    torch.manual_seed(0)
    mod = SaladRef(dropout=0.0)
    g = torch.Generator().manual_seed(seed)
    n = hw * hw
    x = torch.randn(B, n, 768, generator=g)
    t = torch.randn(B, 768, generator=g)
    x.requires_grad_(True)
    t.requires_grad_(True)
    w = torch.randn(B, 8448, generator=g)
    out = mod((tokens_to_salad_input(x, hw, hw), t))
    (out * w).sum().backward()
    return {
        "desc": out.detach().numpy(),
        "dx_rows": x.grad[:, :4].numpy().copy(),          # first 4 tokens of each image
        "dt": t.grad.numpy(),
        "g_score3_w": mod.score[3].weight.grad.flatten().numpy()[:4096].copy(),
        "g_cluster0_b": mod.cluster_features[0].bias.grad.numpy(),
        "g_token2_w_norm": np.asarray(mod.token_features[2].weight.grad.norm().item()),
        "g_dust_bin": np.asarray(mod.dust_bin.grad.item()),
    }


def make_salad():
    out = {}
    for tag, (B, hw, seed) in {"n256": (2, 16, 1234), "n529": (2, 23, 4321)}.items():
        for k, v in salad_case(B, hw, seed).items():
            out[f"{tag}_{k}"] = v
    np.savez_compressed(os.path.join(HERE, "salad_golden.npz"), **out)
    print("salad_golden.npz:", sorted(out))


def metrics_inputs():
    g = torch.Generator().manual_seed(77)
    N = 203
    train = torch.stack([17.44 + 0.01 * torch.randn(500, generator=g), 78.35 + 0.02 * torch.randn(500, generator=g)], 1)
    tgt = torch.stack([17.44 + 0.01 * torch.randn(N, generator=g), 78.35 + 0.02 * torch.randn(N, generator=g)], 1)
    pred_scaled = torch.randn(N, 2, generator=g)
    cs = torch.randn(N, 2, generator=g)
    cs[0] = torch.tensor([-1.0, 0.0])
    cs[1] = torch.tensor([-2.0, -1e-5])
    true_deg = torch.rand(N, generator=g) * 360.0
    return train.double().numpy(), tgt.double().numpy(), pred_scaled.numpy(), cs, true_deg


def make_metrics():
    """dinov2salad_finetuning_2.py:202-215 (sklearn, targets scaled) and dinov2salad_validation.py:84-101 (targets in
    original units), and the MAAE loop body of angle_prediction/efficient_net/validation_script.py:180-202."""
    import math
    from sklearn.metrics import mean_absolute_error, mean_squared_error
    from sklearn.preprocessing import StandardScaler
    train, tgt, pred_scaled, cs, true_deg = metrics_inputs()
    scaler = StandardScaler()
    scaler.fit(train)
    out = {"scaler_mean": scaler.mean_, "scaler_scale": scaler.scale_}
    tgt_scaled = scaler.transform(tgt).astype(np.float32)
    preds_original = scaler.inverse_transform(pred_scaled)
    targets_original = scaler.inverse_transform(tgt_scaled)
    val_mse = mean_squared_error(targets_original, preds_original)
    out["f2_mse"] = np.asarray(val_mse)
    out["f2_rmse"] = np.asarray(math.sqrt(val_mse))
    out["f2_mae"] = np.asarray(mean_absolute_error(targets_original, preds_original))
    out["f2_mae_lat"] = np.asarray(mean_absolute_error(targets_original[:, 0], preds_original[:, 0]))
    out["f2_mae_lon"] = np.asarray(mean_absolute_error(targets_original[:, 1], preds_original[:, 1]))
    all_preds = np.vstack([scaler.inverse_transform(pred_scaled)])
    all_targets = np.vstack([tgt.astype(np.float32)])
    out["val_final_loss"] = np.asarray(
        0.5 * (np.sum((all_preds[:, 0] - all_targets[:, 0]) ** 2) + np.sum((all_preds[:, 1] - all_targets[:, 1]) ** 2))
        / len(all_preds))
    preds_cos_sin = cs
    pred_angle_rad = torch.atan2(preds_cos_sin[:, 1], preds_cos_sin[:, 0])
    pred_angle_deg = torch.rad2deg(pred_angle_rad)
    pred_angle_deg = (pred_angle_deg + 360.0) % 360.0
    diff = torch.abs(pred_angle_deg - true_deg)
    batch_angular_errors = torch.minimum(diff, 360.0 - diff)
    out["maae_cos_sin"] = np.asarray(torch.sum(batch_angular_errors).item() / cs.size(0))
    out["deg_cos_sin"] = pred_angle_deg.numpy()
    np.savez(os.path.join(HERE, "metrics_golden.npz"), **out)
    print("metrics_golden.npz:", sorted(out))


def make_seam():
    """HuggingFace Dinov2Model (random init): hidden_states[-1] is the input of its final LayerNorm and
    last_hidden_state the output (the backbone for config 1)."""
    from transformers import Dinov2Config, Dinov2Model
    torch.manual_seed(5)
    cfg = Dinov2Config(hidden_size=768, num_hidden_layers=2, num_attention_heads=12, mlp_ratio=4, patch_size=14,
                       image_size=56)
    model = Dinov2Model(cfg).eval()
    with torch.no_grad():
        model.layernorm.weight.copy_(1.0 + 0.1 * torch.randn(768))
        model.layernorm.bias.copy_(0.1 * torch.randn(768))
        o = model(pixel_values=torch.randn(2, 3, 56, 56), output_hidden_states=True)
    np.savez_compressed(os.path.join(HERE, "seam_golden.npz"), hidden=o.hidden_states[-1].numpy(),
                        out=o.last_hidden_state.numpy(), ln_w=model.layernorm.weight.detach().numpy(),
                        ln_b=model.layernorm.bias.detach().numpy(), eps=np.asarray(cfg.layer_norm_eps))
    print("seam_golden.npz:", tuple(o.last_hidden_state.shape), "eps", cfg.layer_norm_eps)


PREPROCESS_CASES = [(37, 53, 16, 16), (480, 640, 224, 224), (64, 64, 64, 64), (100, 300, 56, 56), (1000, 750, 64, 48),
                    (17, 9, 40, 33), (333, 517, 46, 46), (64, 500, 64, 64), (3, 2, 5, 7), (2000, 31, 8, 62)]


def preprocess_inputs(i):
    """Seeded uint8 RGB image [H, W, 3]: noise on top of a smooth ramp (both matter: the ramp exercises rounding,
    the noise every tap of the antialiasing filter)."""
    H, W, _, _ = PREPROCESS_CASES[i]
    rng = np.random.RandomState(100 + i)
    yy, xx = np.mgrid[0:H, 0:W]
    ramp = np.stack([(xx * 255.0 / max(W - 1, 1)), (yy * 255.0 / max(H - 1, 1)), ((xx + yy) % 256)], axis=2)
    noise = rng.randint(-40, 41, (H, W, 3))
    return np.clip(ramp + noise, 0, 255).astype(np.uint8)


def make_preprocess():
    """The reference's transform, dinov2salad/dinov2salad_finetuning.py:45-50, through torchvision + Pillow."""
    from PIL import Image
    from torchvision import transforms
    out = {}
    for i, (H, W, oh, ow) in enumerate(PREPROCESS_CASES):
        pil = Image.fromarray(preprocess_inputs(i), "RGB")
        tf = transforms.Compose([transforms.Resize((oh, ow)), transforms.ToTensor(),
                                 transforms.Normalize(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225])])
        out[f"u8_{i}"] = np.array(transforms.Resize((oh, ow))(pil))
        if oh * ow <= 64 * 64:
            out[f"f32_{i}"] = tf(pil).numpy()
        else:                                   # full-size float outputs are large: keep a strided sample
            out[f"f32s_{i}"] = tf(pil).numpy()[:, ::7, ::5].copy()
    np.savez_compressed(os.path.join(HERE, "preprocess_golden.npz"), **out)
    print("preprocess_golden.npz:", len(out), "arrays")


if __name__ == "__main__":
    torch.set_num_threads(8)
    only = sys.argv[1:]
    for name, fn in (("heads", make_heads), ("salad", make_salad), ("metrics", make_metrics), ("seam", make_seam),
                     ("preprocess", make_preprocess)):
        if not only or name in only:
            fn()
