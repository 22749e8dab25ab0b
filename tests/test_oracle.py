"""CPU tests pinning the oracle: against the reference's own functions (golden
vectors lifted from /root/reference by tests/golden/make_golden.py), against the
committed SALAD regression fixtures, and against known-answer properties."""
import math
import os

import numpy as np
import pytest
import torch

from oracle import salad_ref as O
from tests.golden.make_golden import heads_inputs


@pytest.fixture(scope="module")
def hg(golden_dir):
    return np.load(os.path.join(golden_dir, "heads_golden.npz"))


@pytest.fixture(scope="module")
def sg(golden_dir):
    return np.load(os.path.join(golden_dir, "salad_golden.npz"))


def _np(a):
    return a.detach().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)


def close(a, b, rtol=1e-6, atol=1e-6):
    np.testing.assert_allclose(_np(a), _np(b), rtol=rtol, atol=atol)


def test_heads_losses_match_reference_functions(hg):
    sc, tgt_deg, pred_deg, ll_pred, ll_tgt = heads_inputs()
    tgt_sc = O.deg_to_sincos(tgt_deg)
    unit = torch.nn.functional.normalize(sc, dim=1, p=2, eps=1e-6)
    close(O.sincos_to_deg(sc), hg["deg_from_sincos"], atol=1e-4)
    close(O.maae_loss_deg(pred_deg, tgt_deg), hg["maae_deg"])
    close(O.sincos_atan2_maae(sc, tgt_deg), hg["maae_atan2"])
    close(O.maae_loss_scalar(pred_deg % 360, tgt_deg), hg["maae_scalar"])
    close(O.acos_angular_loss(unit, tgt_sc), hg["acos_loss"])
    close(O.angle_error_atan2(unit, tgt_sc), hg["angle_err_atan2"])
    close(O.maae_sincos_metric(sc, tgt_sc), hg["maae_sincos_metric"])
    close(O.mse_loss(ll_pred, ll_tgt), hg["mse"])
    close(O.huber_loss(ll_pred, ll_tgt), hg["huber"])
    close(O.mse_loss(sc, tgt_sc), hg["mse_sincos"])
    close(O.latlon_validation_score(ll_pred, ll_tgt), hg["val_score"], rtol=1e-5)


def test_heads_loss_gradients_match_reference_functions(hg):
    sc, tgt_deg, _, ll_pred, ll_tgt = heads_inputs()
    tgt_sc = O.deg_to_sincos(tgt_deg)
    cases = {
        "maae_atan2": lambda s: O.sincos_atan2_maae(s, tgt_deg),
        "acos_loss": lambda s: O.acos_angular_loss(
            torch.nn.functional.normalize(s, dim=1, p=2, eps=1e-6), tgt_sc),
        "mse_sincos": lambda s: O.mse_loss(s, tgt_sc),
    }
    for name, fn in cases.items():
        s = sc.clone().requires_grad_(True)
        fn(s).backward()
        close(s.grad, hg["grad_" + name], rtol=1e-5, atol=1e-7)
    for name, fn in (("mse", O.mse_loss), ("huber", O.huber_loss)):
        p = ll_pred.clone().requires_grad_(True)
        fn(p, ll_tgt).backward()
        close(p.grad, hg["grad_" + name], rtol=1e-6, atol=1e-8)


def test_scalar_angle_head_oracle_matches_reference_functions(hg):
    """dino_angle_finetuning.py:37-52: Linear(D,1) -> % 360 -> |p - t| % 360 -> min(d, 360 - d)."""
    _, tgt_deg, pred_deg, _, _ = heads_inputs()
    head = O.ScalarAngleHeadRef(4)
    with torch.no_grad():
        head.head.weight.zero_()
        head.head.weight[0, 0] = 1.0
        head.head.bias.zero_()
    d = torch.zeros(pred_deg.shape[0], 4)
    d[:, 0] = pred_deg
    d.requires_grad_(True)
    out = head(d)
    assert out.shape == pred_deg.shape and out.min() >= 0 and out.max() < 360
    loss = O.maae_loss_scalar(out, tgt_deg % 360)
    close(loss, hg["maae_scalar_wrapped_targets"])
    loss.backward()
    close(d.grad[:, 0], hg["grad_maae_scalar"], rtol=1e-6, atol=1e-9)


def _salad_case(B, hw, seed, aggregate="repeat"):
    torch.manual_seed(0)
    mod = O.SaladRef(dropout=0.0, aggregate=aggregate)
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, hw * hw, 768, generator=g)
    t = torch.randn(B, 768, generator=g)
    w = torch.randn(B, 8448, generator=g)
    return mod, x, t, w


@pytest.mark.parametrize("tag,B,hw,seed", [("n256", 2, 16, 1234), ("n529", 2, 23, 4321)])
def test_salad_regression_fixture(sg, tag, B, hw, seed):
    mod, x, t, w = _salad_case(B, hw, seed)
    x.requires_grad_(True)
    t.requires_grad_(True)
    out = mod((O.tokens_to_salad_input(x, hw, hw), t))
    (out * w).sum().backward()
    close(out.detach(), sg[f"{tag}_desc"], rtol=1e-5, atol=1e-7)
    close(x.grad[:, :4], sg[f"{tag}_dx_rows"], rtol=1e-4, atol=1e-8)
    close(t.grad, sg[f"{tag}_dt"], rtol=1e-4, atol=1e-8)
    close(mod.cluster_features[0].bias.grad, sg[f"{tag}_g_cluster0_b"], rtol=1e-4, atol=1e-7)
    assert abs(float(mod.dust_bin.grad)) < 1e-6          # SURVEY A-3b: d/d dust_bin == 0


def test_salad_contract_and_known_answers():
    """Width 8448 (reference dinov2salad_finetuning.py:29,36), unit rows, parameter
    inventory, sqrt(1+m) pre-normalisation norm."""
    mod, x, t, _ = _salad_case(3, 16, 7)
    out = mod((O.tokens_to_salad_input(x, 16, 16), t))
    assert out.shape == (3, 8448)
    close(out.norm(dim=1), torch.ones(3), atol=1e-6)
    assert sum(p.numel() for p in mod.parameters()) == 1411009
    sd = mod.state_dict()
    assert sd["dust_bin"].shape == ()
    assert sd["cluster_features.0.weight"].shape == (512, 768, 1, 1)
    assert sd["cluster_features.3.weight"].shape == (128, 512, 1, 1)
    assert sd["score.3.weight"].shape == (64, 512, 1, 1)
    assert sd["token_features.2.weight"].shape == (256, 512)
    # token block has norm 1/sqrt(65), every cluster column too
    tok = out[:, :256].norm(dim=1)
    close(tok, torch.full((3,), 1 / math.sqrt(65.0)), rtol=1e-5)
    cl = out[:, 256:].reshape(3, 128, 64).norm(dim=1)
    close(cl, torch.full((3, 64), 1 / math.sqrt(65.0)), rtol=1e-5)


def test_einsum_equals_repeat_aggregation():
    mod, x, t, _ = _salad_case(2, 16, 11)
    a = mod((O.tokens_to_salad_input(x, 16, 16), t))
    mod.aggregate = "einsum"
    b = mod((O.tokens_to_salad_input(x, 16, 16), t))
    close(a.detach(), b.detach(), rtol=1e-5, atol=1e-7)


@pytest.mark.parametrize("n", [256, 529, 100])
def test_sinkhorn_marginals(n):
    g = torch.Generator().manual_seed(n)
    S = torch.randn(2, 64, n, generator=g, dtype=torch.float64) * 2
    logP = O.get_matching_probs(S, torch.tensor(1.0, dtype=torch.float64), 3, const_dtype=torch.float64)
    P = logP.exp()
    close(P.sum(1), torch.ones(2, n, dtype=torch.float64), atol=1e-12)   # columns exact after last v
    assert P.shape == (2, 65, n)
    assert (P[:, :64].sum(2) - 1).abs().max() < 0.35                     # cluster rows ~ 1
    assert (P[:, 64].sum(1) - (n - 64)).abs().max() < 0.35 * 64


def test_dustbin_value_is_irrelevant():
    g = torch.Generator().manual_seed(5)
    S = torch.randn(2, 64, 256, generator=g, dtype=torch.float64)
    a = O.get_matching_probs(S, torch.tensor(1.0, dtype=torch.float64), 3, const_dtype=torch.float64)
    b = O.get_matching_probs(S, torch.tensor(-7.3, dtype=torch.float64), 3, const_dtype=torch.float64)
    close(a[:, :64].exp(), b[:, :64].exp(), rtol=1e-12, atol=0)


def test_explicit_sinkhorn_backward_equals_autograd():
    g = torch.Generator().manual_seed(9)
    S = torch.randn(2, 64, 96, generator=g, dtype=torch.float64, requires_grad=True)
    dP = torch.randn(2, 64, 96, generator=g, dtype=torch.float64)
    P = O.get_matching_probs(S, torch.tensor(1.0, dtype=torch.float64), 3,
                             const_dtype=torch.float64).exp()[:, :64]
    (P * dP).sum().backward()
    dS = O.sinkhorn_backward_explicit(S.detach(), dP, 3)
    close(dS, S.grad, rtol=1e-9, atol=1e-13)


def test_gradcheck_small_salad_fp64():
    torch.manual_seed(1)
    mod = O.SaladRef(num_channels=16, num_clusters=4, cluster_dim=8, token_dim=8, dropout=0.0).double()
    x = torch.randn(2, 16, 3, 3, dtype=torch.float64, requires_grad=True)
    t = torch.randn(2, 16, dtype=torch.float64, requires_grad=True)
    assert torch.autograd.gradcheck(lambda a, b: mod((a, b), const_dtype=torch.float64), (x, t),
                                    eps=1e-6, atol=1e-5)


def test_drop_mask_injection_matches_dropout_scaling():
    mod, x, t, _ = _salad_case(2, 16, 3)
    mod.p = 0.3
    ones = torch.ones(2, 1024, 16, 16)
    a = mod((O.tokens_to_salad_input(x, 16, 16), t), drop_mask=ones * 0.7)   # mask*1/(1-p) == 1
    b = mod((O.tokens_to_salad_input(x, 16, 16), t))
    close(a.detach(), b.detach(), rtol=1e-5, atol=1e-7)


def test_knn_ref_ties_and_merge():
    g = torch.Generator().manual_seed(2)
    G = torch.randn(500, 64, generator=g).bfloat16()
    G[100] = G[7]                      # exact duplicate row -> tie, lower index must win
    Q = torch.randn(9, 64, generator=g).bfloat16()
    s, i = O.knn_topk_ref(Q, G, k=25, chunk=128)
    full = Q.float() @ G.float().T
    ts, ti = torch.topk(full, 25, dim=1)
    close(s, ts, rtol=0, atol=0)
    assert s.shape == (9, 25) and (s[:, :-1] >= s[:, 1:]).all()
    for r in range(9):
        assert set(i[r].tolist()) == set(ti[r].tolist()) or 100 in ti[r].tolist() or 7 in ti[r].tolist()
        pos = {int(v): k for k, v in enumerate(i[r].tolist())}
        if 7 in pos and 100 in pos:
            assert pos[7] < pos[100]
    # sharded merge == unsharded
    s1, i1 = O.knn_topk_ref(Q, G[:250], 25, g_offset=0)
    s2, i2 = O.knn_topk_ref(Q, G[250:], 25, g_offset=250)
    ms, mi = O.merge_topk(torch.cat([s1, s2], 1), torch.cat([i1, i2], 1), 25)
    assert torch.equal(mi, i) and torch.equal(ms, s)


# ------------------------------------------------------------------ rows next to the path
def test_seam_oracle_matches_huggingface_dinov2(golden_dir):
    from oracle import extras_ref as X
    z = np.load(os.path.join(golden_dir, "seam_golden.npz"))
    x_tok, t = X.seam_ref(torch.from_numpy(z["hidden"]), torch.from_numpy(z["ln_w"]), torch.from_numpy(z["ln_b"]),
                          float(z["eps"]))
    out = torch.from_numpy(z["out"])
    assert torch.allclose(x_tok, out[:, 1:], atol=1e-5) and torch.allclose(t, out[:, 0], atol=1e-5)


def test_metric_oracles_match_reference_expressions(golden_dir):
    from oracle import extras_ref as X
    from tests.golden.make_golden import heads_inputs, metrics_inputs
    z = np.load(os.path.join(golden_dir, "metrics_golden.npz"))
    hg = np.load(os.path.join(golden_dir, "heads_golden.npz"))
    train, tgt, pred_scaled, cs, true_deg = metrics_inputs()
    mean, scale = z["scaler_mean"], z["scaler_scale"]
    tgt_scaled = ((tgt - mean) / scale).astype(np.float32)
    m = X.latlon_metrics_ref(pred_scaled, tgt_scaled, mean, scale, targets_scaled=True)
    # the reference de-normalises in float32 (sklearn keeps the dtype of the float32 predictions) at magnitudes of
    # 17.4 / 78.4 degrees, so its own numbers carry ~1e-5 relative rounding noise; the oracle works in float64
    for key in ("mse", "rmse", "mae", "mae_lat", "mae_lon"):
        np.testing.assert_allclose(m[key], float(z["f2_" + key]), rtol=1e-4)
    m2 = X.latlon_metrics_ref(pred_scaled, tgt.astype(np.float32), mean, scale)
    np.testing.assert_allclose(m2["final_loss"], float(z["val_final_loss"]), rtol=1e-4)
    val, deg = X.maae_ref(cs, true_deg, "cos_sin")
    np.testing.assert_allclose(val, float(z["maae_cos_sin"]), rtol=1e-6)
    np.testing.assert_allclose(deg.numpy(), z["deg_cos_sin"], rtol=1e-6)
    sc, tgt_deg, *_ = heads_inputs()
    val2, _ = X.maae_ref(sc, tgt_deg, "sin_cos")
    np.testing.assert_allclose(val2, float(hg["maae_atan2"]), rtol=1e-6)
    idx = torch.tensor([[3, 1, 2], [0, 0, 0], [2, 3, 1]])
    got = X.recall_at_k_ref(idx, torch.tensor([7, 9, 5]), torch.tensor([1, 7, 5, 8]), (1, 2, 3))
    assert got == {1: 1 / 3, 2: 2 / 3, 3: 2 / 3}



def test_preprocess_oracle_matches_torchvision_pillow_golden(golden_dir):
    from oracle import extras_ref as X
    from tests.golden.make_golden import PREPROCESS_CASES, preprocess_inputs
    z = np.load(os.path.join(golden_dir, "preprocess_golden.npz"))
    for i, (H, W, oh, ow) in enumerate(PREPROCESS_CASES):
        if H * W > 400 * 400:
            continue                                  # the big cases run on the GPU side only (seconds matter here)
        img = preprocess_inputs(i)
        assert np.array_equal(X.pil_resize_ref(img, oh, ow), z[f"u8_{i}"])
        got = X.preprocess_ref(img, oh, ow).numpy()
        if f"f32_{i}" in z.files:
            assert np.array_equal(got, z[f"f32_{i}"])
        else:
            assert np.array_equal(got[:, ::7, ::5], z[f"f32s_{i}"])
