"""GPU parity tests for the rows next to the hot path: backbone->head seam, the reference's model
wrapper + checkpoint layout, and the validation metrics.  Same rules as test_gpu_parity.py: every CUDA path is called
through the C-ABI library and compared with the CPU oracle / the committed golden vectors."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import extras_ref as X  # noqa: E402
from oracle import salad_ref as O  # noqa: E402

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def vpr():
    import __graft_entry__ as g
    if not os.path.exists(os.path.join(g.PKG_DIR, "libvpr_b200.so")):
        g.build()
    import vpr_b200
    assert torch.cuda.is_available()
    return vpr_b200


def rel(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


# ------------------------------------------------------------------ seam
def test_seam_matches_huggingface_dinov2_final_layernorm(vpr):
    z = np.load(os.path.join(GOLD, "seam_golden.npz"))
    seam = vpr.BackboneSeam(768, eps=float(z["eps"]), out_dtype=torch.float32).cuda()
    with torch.no_grad():
        seam.norm.weight.copy_(torch.from_numpy(z["ln_w"]))
        seam.norm.bias.copy_(torch.from_numpy(z["ln_b"]))
    x_tok, t = seam(torch.from_numpy(z["hidden"]).cuda())
    out = torch.from_numpy(z["out"])
    assert x_tok.shape == (2, 16, 768) and t.shape == (2, 768)
    assert rel(x_tok, out[:, 1:]) < 1e-5 and rel(t, out[:, 0]) < 1e-5
    assert (x_tok.cpu() - out[:, 1:]).abs().max() < 2e-5


@pytest.mark.parametrize("d,skip,n,B", [(768, 0, 256, 3), (1024, 4, 37, 2), (1536, 0, 16, 2), (384, 1, 529, 1), (2048, 0, 5, 2),
                                        (8, 0, 1, 1)])
@pytest.mark.parametrize("in_dtype,out_dtype", [(torch.float32, torch.float32), (torch.float32, torch.bfloat16),
                                                (torch.bfloat16, torch.bfloat16), (torch.bfloat16, torch.float32)])
def test_seam_forward_backward_against_oracle(vpr, d, skip, n, B, in_dtype, out_dtype):
    g = torch.Generator().manual_seed(d + n)
    hidden = (torch.randn(B, 1 + skip + n, d, generator=g) * 2.0 + 0.3).to(in_dtype)
    w = 1.0 + 0.2 * torch.randn(d, generator=g)
    b = 0.2 * torch.randn(d, generator=g)
    gx = torch.randn(B, n, d, generator=g).to(out_dtype)
    gt = torch.randn(B, d, generator=g).to(out_dtype)
    seam = vpr.BackboneSeam(d, eps=1e-6, num_register_tokens=skip, out_dtype=out_dtype).cuda()
    with torch.no_grad():
        seam.norm.weight.copy_(w)
        seam.norm.bias.copy_(b)
    hg = hidden.cuda().requires_grad_(True)
    x_tok, t = seam(hg)
    assert x_tok.dtype == out_dtype and x_tok.is_contiguous()
    hr = hidden.float().requires_grad_(True)
    wr, br = w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    xr, tr = X.seam_ref(hr, wr, br, 1e-6, skip)
    if out_dtype == torch.float32:
        assert rel(x_tok, xr) < 1e-5 and rel(t, tr) < 1e-5
    else:        # correctly rounded fp32 result, up to a last-place flip where the fp32 values differ by 1e-7
        assert rel(x_tok, xr.bfloat16()) < 1e-3 and rel(t, tr.bfloat16()) < 1e-3
    (x_tok.float() * gx.cuda().float()).sum().backward(retain_graph=True)
    (t.float() * gt.cuda().float()).sum().backward()
    ((xr * gx.float()).sum() + (tr * gt.float()).sum()).backward()
    tol = 1e-4 if in_dtype == torch.float32 else 1e-2            # d_hidden is stored in the hidden state's dtype
    assert rel(hg.grad, hr.grad) < tol
    assert rel(seam.norm.weight.grad, wr.grad) < 1e-4
    assert rel(seam.norm.bias.grad, br.grad) < 1e-4
    if skip:
        assert hg.grad[:, 1:1 + skip].abs().max().item() == 0.0


def test_seam_partial_gradients_and_no_grad_path(vpr):
    torch.manual_seed(0)
    hidden = torch.randn(2, 1 + 9, 768)
    seam = vpr.BackboneSeam(768, out_dtype=torch.float32).cuda()
    hg = hidden.cuda().requires_grad_(True)
    x_tok, t = seam(hg)
    t.sum().backward()                                  # only the CLS branch has a gradient
    hr = hidden.clone().requires_grad_(True)
    xr, tr = X.seam_ref(hr, seam.norm.weight.detach().cpu(), seam.norm.bias.detach().cpu())
    tr.sum().backward()
    assert (hg.grad - hr.grad.cuda()).abs().max().item() < 1e-5   # sum over an LN output: gradient ~ 0 for unit gamma
    assert hg.grad[:, 1:].abs().max().item() == 0.0
    with torch.no_grad():
        x2, t2 = seam(hidden.cuda())
    assert torch.equal(x2, x_tok) and torch.equal(t2, t)
    with pytest.raises(ValueError):
        seam(torch.randn(2, 10, 512, device="cuda"))
    with pytest.raises(RuntimeError):
        seam(hidden)                                    # CPU tensor: no fallback


def test_seam_feeds_salad_head_like_the_upstream_wrapper(vpr):
    """hidden -> seam -> SaladHead  ==  oracle LN + split + reshape/permute -> SaladRef."""
    torch.manual_seed(0)
    ref = O.SaladRef(dropout=0.0).eval()
    head = vpr.SaladHead(dropout=0.0).cuda().eval()
    head.load_state_dict(ref.state_dict())
    g = torch.Generator().manual_seed(11)
    hidden = torch.randn(2, 1 + 256, 768, generator=g)
    seam = vpr.BackboneSeam(768, out_dtype=torch.float32).cuda()
    model = vpr.VPRModel(torch.nn.Identity(), head, seam).cuda().eval()
    with torch.no_grad():
        out = model(hidden.cuda())
        xr, tr = X.seam_ref(hidden, seam.norm.weight.cpu(), seam.norm.bias.cpu())
        want = ref((O.tokens_to_salad_input(xr, 16, 16), tr))
    assert out.shape == (2, 8448)
    assert rel(out, want) < 1e-4


# ------------------------------------------------------------------ reference model wrapper + checkpoints
class _FakeExtractor(torch.nn.Module):
    """stands in for the frozen hub model: images -> descriptor (a fixed random projection)"""

    def __init__(self):
        super().__init__()
        self.proj = torch.nn.Linear(32, 8448)

    def forward(self, x):
        return torch.nn.functional.normalize(self.proj(x), dim=1)


@pytest.mark.parametrize("loss,dropout", [("mse", 0.0), ("huber", 0.0), ("huber", 0.3)])
def test_regression_model_matches_reference_training_lines(vpr, loss, dropout):
    torch.manual_seed(1)
    fe = _FakeExtractor()
    model = vpr.DINOv2RegressionModel(fe, dropout_prob=dropout, loss=loss).cuda()
    assert all(not p.requires_grad for p in model.feature_extractor.parameters())
    last = 3 if dropout > 0 else 2
    assert set(k for k in model.state_dict() if k.startswith("regressor")) == {
        "regressor.0.weight", "regressor.0.bias", f"regressor.{last}.weight", f"regressor.{last}.bias"}
    ll = O.LatLonRegressorRef(8448, 512)
    ll.regressor[0].load_state_dict(model.regressor[0].state_dict())
    ll.regressor[2].load_state_dict(model.regressor[last].state_dict())
    x = torch.randn(24, 32)
    tg = torch.randn(24, 2)
    feats = fe.cpu()(x).detach()
    fe.cuda()
    model.eval()
    preds = model(x.cuda())
    want = ll(feats)
    assert rel(preds, want) < 1e-4
    # the reference's training lines, unchanged (dinov2salad_finetuning.py:115-124): model.train();
    # preds = model(inputs); loss = loss_fn(preds, targets); optimizer.zero_grad(); loss.backward()
    model.train()
    if dropout > 0:
        model.eval()                                   # parity without the random mask (the masked path is checked below)
    loss_fn = torch.nn.MSELoss() if loss == "mse" else torch.nn.HuberLoss(delta=1.0)
    preds_t = model(x.cuda())
    assert preds_t.requires_grad
    l3 = loss_fn(preds_t, tg.cuda())
    l3.backward()
    ll3 = O.LatLonRegressorRef(8448, 512)
    ll3.load_state_dict(ll.state_dict())
    r3 = (O.mse_loss if loss == "mse" else O.huber_loss)(ll3(feats), tg)
    r3.backward()
    assert abs(l3.item() - r3.item()) < 1e-4 * abs(r3.item())
    for pm, pr in ((model.regressor[0].weight, ll3.regressor[0].weight), (model.regressor[0].bias, ll3.regressor[0].bias),
                   (model.regressor[last].weight, ll3.regressor[2].weight), (model.regressor[last].bias, ll3.regressor[2].bias)):
        assert rel(pm.grad, pr.grad) < 1e-4
        pm.grad = None
    out = model.loss(x.cuda(), tg.cuda())
    out.backward()
    ref_loss = (O.mse_loss if loss == "mse" else O.huber_loss)(want, tg)
    ref_loss.backward()
    assert abs(out.item() - ref_loss.item()) < 1e-4 * abs(ref_loss.item())
    assert rel(model.regressor[0].weight.grad, ll.regressor[0].weight.grad) < 1e-4
    assert rel(model.regressor[last].bias.grad, ll.regressor[2].bias.grad) < 1e-4
    if dropout > 0:
        model.train()
        a = model.loss(x.cuda(), tg.cuda()).item()
        b = model.loss(x.cuda(), tg.cuda()).item()
        assert a != b                                   # a fresh mask every step
        # train-mode forward(): the backward must replay the mask the forward drew -- check against the oracle with
        # that exported mask
        for pm in model.regressor.parameters():
            pm.grad = None
        seed = (model.seed * 1000003 + model._step) & 0xFFFFFFFFFFFFFFFF
        pt = model(x.cuda())
        loss_fn(pt, tg.cuda()).backward()
        keep = vpr.dropout_keep_mask(seed, dropout, 24, 512).cpu().float()
        ll4 = O.LatLonRegressorRef(8448, 512)
        ll4.load_state_dict(ll.state_dict())
        z = torch.relu(ll4.regressor[0](feats)) * keep / (1 - dropout)
        p4 = ll4.regressor[2](z)
        assert rel(pt, p4) < 1e-4
        O.huber_loss(p4, tg).backward()
        assert rel(model.regressor[0].weight.grad, ll4.regressor[0].weight.grad) < 1e-4
        assert rel(model.regressor[last].weight.grad, ll4.regressor[2].weight.grad) < 1e-4


def test_reference_checkpoint_round_trip(vpr, tmp_path):
    torch.manual_seed(2)
    ref = O.SaladRef(dropout=0.3)
    base = vpr.VPRModel(torch.nn.Identity(), vpr.SaladHead())
    model = vpr.DINOv2RegressionModel(base)
    # a checkpoint as the reference writes it (dinov2salad_finetuning.py:130-135), with backbone entries we do not own
    sd = {f"feature_extractor.aggregator.{k}": v.clone() for k, v in ref.state_dict().items()}
    sd["feature_extractor.backbone.model.cls_token"] = torch.zeros(1, 1, 768)
    reg = O.LatLonRegressorRef(8448, 512)
    sd.update({f"regressor.{k}": v.clone() for k, v in reg.regressor.state_dict().items()})
    path = tmp_path / "checkpoint_7_.pth"
    torch.save({"epoch": 7, "model_state_dict": sd, "optimizer_state_dict": {}, "loss": torch.tensor(0.5)}, path)
    info = vpr.load_reference_checkpoint(model, str(path))
    assert info["epoch"] == 7 and info["skipped"] == ["feature_extractor.backbone.model.cls_token"]
    for k, v in ref.state_dict().items():
        assert torch.equal(model.feature_extractor.aggregator.state_dict()[k], v)
    assert torch.equal(model.regressor[0].weight, reg.regressor[0].weight)
    # and back: the file we write loads into the reference's module tree
    out = tmp_path / "checkpoint_8_.pth"
    vpr.save_reference_checkpoint(str(out), model, None, 8, 0.25)
    back = torch.load(out, weights_only=False)
    assert set(back) == {"epoch", "model_state_dict", "optimizer_state_dict", "loss"}
    agg = {k[len("feature_extractor.aggregator."):]: v for k, v in back["model_state_dict"].items()
           if k.startswith("feature_extractor.aggregator.")}
    ref2 = O.SaladRef()
    ref2.load_state_dict(agg)                           # strict
    # upstream dino_salad.ckpt layout (Lightning: {'state_dict': {'aggregator.*', 'backbone.*'}})
    head = vpr.SaladHead()
    vpr.load_upstream_salad_weights(head, {"state_dict": {f"aggregator.{k}": v for k, v in ref.state_dict().items()}})
    assert torch.equal(head.state_dict()["score.3.weight"], ref.state_dict()["score.3.weight"])
    with pytest.raises(KeyError):
        vpr.load_upstream_salad_weights(head, {"state_dict": {"backbone.x": torch.zeros(1)}})
    # end to end on the device with the loaded weights
    model = model.cuda().eval()
    g = torch.Generator().manual_seed(3)
    f = torch.randn(2, 256, 768, generator=g)
    t = torch.randn(2, 768, generator=g)
    preds = model((f.cuda(), t.cuda()))
    want = reg(ref.eval()((O.tokens_to_salad_input(f, 16, 16), t)))
    assert rel(preds, want) < 1e-4


# ------------------------------------------------------------------ validation metrics
def test_latlon_metrics_match_sklearn_and_validation_script_golden(vpr):
    from tests.golden.make_golden import metrics_inputs
    z = np.load(os.path.join(GOLD, "metrics_golden.npz"))
    train, tgt, pred_scaled, _, _ = metrics_inputs()
    scaler = vpr.LatLonScaler.fit(train)
    np.testing.assert_allclose(scaler.mean_, z["scaler_mean"], rtol=1e-12)
    np.testing.assert_allclose(scaler.scale_, z["scaler_scale"], rtol=1e-12)
    pred = torch.from_numpy(pred_scaled).cuda()
    # dinov2salad_finetuning_2.py:202-215: targets are z-scored too
    tgt_scaled = torch.from_numpy(scaler.transform(tgt).astype(np.float32)).cuda()
    m = vpr.latlon_metrics(pred, tgt_scaled, scaler, targets_scaled=True)
    # 1e-4: the reference's own float32 de-normalisation noise (see tests/test_oracle.py); 1e-9 against the fp64 oracle
    for key in ("mse", "rmse", "mae", "mae_lat", "mae_lon"):
        np.testing.assert_allclose(m[key], float(z["f2_" + key]), rtol=1e-4)
    want_scaled = X.latlon_metrics_ref(pred_scaled, tgt_scaled.cpu().numpy(), scaler.mean_, scaler.scale_, targets_scaled=True)
    for key, v in want_scaled.items():
        np.testing.assert_allclose(m[key], v, rtol=1e-9)
    # dinov2salad_validation.py:84-101: targets in original units (float32, as the loader delivers them)
    m2 = vpr.latlon_metrics(pred, torch.from_numpy(tgt.astype(np.float32)).cuda(), scaler, targets_scaled=False)
    np.testing.assert_allclose(m2["final_loss"], float(z["val_final_loss"]), rtol=1e-4)
    want = X.latlon_metrics_ref(pred_scaled, tgt.astype(np.float32), scaler.mean_, scaler.scale_)
    for key, v in want.items():
        np.testing.assert_allclose(m2[key], v, rtol=1e-9)
    # no scaler: plain sums; and the empty batch
    m3 = vpr.latlon_metrics(pred, tgt_scaled)
    np.testing.assert_allclose(m3["mse"], X.latlon_metrics_ref(pred_scaled, tgt_scaled.cpu().numpy())["mse"], rtol=1e-9)
    assert vpr.latlon_metrics(pred[:0], tgt_scaled[:0]) == {"n": 0}
    big = torch.randn(300001, 2, device="cuda")
    mb = vpr.latlon_metrics(big, torch.zeros_like(big))
    np.testing.assert_allclose(mb["mse"], big.double().pow(2).mean().item(), rtol=1e-10)


def test_maae_matches_reference_golden_both_orders(vpr):
    from tests.golden.make_golden import heads_inputs, metrics_inputs
    z = np.load(os.path.join(GOLD, "metrics_golden.npz"))
    hg = np.load(os.path.join(GOLD, "heads_golden.npz"))
    _, _, _, cs, true_deg = metrics_inputs()
    val, deg = vpr.maae(cs.cuda(), true_deg.cuda(), order="cos_sin")
    np.testing.assert_allclose(deg.cpu().numpy(), z["deg_cos_sin"], rtol=1e-5, atol=2e-4)
    np.testing.assert_allclose(val, float(z["maae_cos_sin"]), rtol=1e-5)
    sc, tgt_deg, *_ = heads_inputs()                      # dino_v2_gemini.py functions (targets up to 750 degrees)
    val2, deg2 = vpr.maae(sc.cuda(), tgt_deg.cuda(), order="sin_cos")
    np.testing.assert_allclose(deg2.cpu().numpy(), hg["deg_from_sincos"], rtol=1e-5, atol=2e-4)
    np.testing.assert_allclose(val2, float(hg["maae_atan2"]), rtol=1e-5)
    want, _ = X.maae_ref(sc, tgt_deg, "sin_cos")
    np.testing.assert_allclose(val2, want, rtol=1e-5)


def test_recall_at_k_on_knn_output(vpr):
    g = torch.Generator().manual_seed(9)
    Q, G, D, k = 300, 5000, 256, 25
    gal = torch.nn.functional.normalize(torch.randn(G, D, generator=g), dim=1).bfloat16()
    gl = torch.randint(0, 400, (G,), generator=g)
    pick = torch.randint(0, G, (Q,), generator=g)
    q = torch.nn.functional.normalize(gal[pick].float() + 0.25 * torch.randn(Q, D, generator=g), dim=1).bfloat16()
    ql = gl[pick].clone()
    ql[::7] = 100000                                       # labels absent from the gallery: never a hit
    _, idx = vpr.knn_topk(q.cuda(), gal.cuda(), k)
    got = vpr.recall_at_k(idx, ql.cuda(), gl.cuda(), ks=(1, 5, 10, 25))
    want = X.recall_at_k_ref(idx.cpu(), ql, gl, (1, 5, 10, 25))
    assert got == want
    assert got[1] <= got[5] <= got[10] <= got[25] <= 1.0 - 1.0 / 7 + 0.01 and got[1] > 0.5
    with pytest.raises(ValueError):
        vpr.recall_at_k(idx, ql.cuda(), gl.cuda(), ks=())
    # cut-offs beyond the list length are refused by the host API ...
    with pytest.raises(ValueError):
        vpr.recall_at_k(idx[:, :10].contiguous(), ql.cuda(), gl.cuda(), ks=(1, 5, 10, 25))
    # ... and at the C boundary a query without any matching neighbour is a miss at every cut-off, including cut-offs
    # larger than the list (the first-hit rank sentinel is "never", not k)
    from vpr_b200 import _lib as L
    lib = vpr.lib()
    idx10 = idx[:, :10].contiguous()
    ks = torch.tensor([1, 10, 25], dtype=torch.int32, device="cuda")
    hits = torch.zeros(3, dtype=torch.int64, device="cuda")
    glc, qlc = gl.cuda(), ql.cuda()
    L.check(lib.vpr_recall_at_k(L.ptr(idx10), L.ptr(glc), L.ptr(qlc), Q, 10, G, L.ptr(ks), 3, L.ptr(hits),
                                L.stream_ptr(idx.device)), "vpr_recall_at_k")
    want10 = X.recall_at_k_ref(idx10.cpu(), ql, gl, (1, 10))
    h = hits.tolist()
    assert h[0] == round(want10[1] * Q) and h[1] == round(want10[10] * Q) and h[2] == h[1] < Q



# ------------------------------------------------------------------ preprocessing
def test_preprocess_is_bit_identical_to_torchvision_pillow_golden(vpr):
    from tests.golden.make_golden import PREPROCESS_CASES, preprocess_inputs
    z = np.load(os.path.join(GOLD, "preprocess_golden.npz"))
    for i, (H, W, oh, ow) in enumerate(PREPROCESS_CASES):
        img = preprocess_inputs(i)
        pre = vpr.GpuPreprocess((oh, ow))
        out, u8 = pre([torch.from_numpy(img)], return_uint8=True)
        assert out.shape == (1, 3, oh, ow) and out.dtype == torch.float32
        assert np.array_equal(u8[0].cpu().numpy(), z[f"u8_{i}"]), (H, W, oh, ow)        # Pillow's resampler, bit for bit
        got = out[0].cpu().numpy()
        if f"f32_{i}" in z.files:
            assert np.array_equal(got, z[f"f32_{i}"])                                  # ToTensor + Normalize, bit for bit
        else:
            assert np.array_equal(got[:, ::7, ::5], z[f"f32s_{i}"])
        want = X.preprocess_ref(img, oh, ow)
        assert torch.equal(out[0].cpu(), want)


def test_preprocess_batch_of_mixed_sizes_and_bf16(vpr):
    rng = np.random.RandomState(5)
    imgs = [torch.from_numpy(rng.randint(0, 256, (h, w, 3)).astype(np.uint8)) for h, w in ((300, 400), (50, 20), (224, 224))]
    pre = vpr.GpuPreprocess((224, 224), out_dtype=torch.bfloat16)
    out = pre([imgs[0].cuda(), imgs[1], imgs[2].pin_memory()])
    assert out.shape == (3, 3, 224, 224) and out.dtype == torch.bfloat16
    for i, im in enumerate(imgs):
        want = X.preprocess_ref(im.numpy(), 224, 224)
        assert torch.equal(out[i].cpu(), want.bfloat16())
    with pytest.raises(ValueError):
        pre([torch.zeros(4, 4, 4, dtype=torch.uint8)])
    with pytest.raises(ValueError):
        pre([torch.zeros(4, 4, 3)])


def test_preprocess_same_size_batch_in_one_call(vpr):
    rng = np.random.RandomState(6)
    imgs = [torch.from_numpy(rng.randint(0, 256, (120, 200, 3)).astype(np.uint8)) for _ in range(5)]
    pre = vpr.GpuPreprocess((56, 56))
    out, u8 = pre(imgs, return_uint8=True)                      # list of equal sizes -> stacked batch path
    out4 = pre(torch.stack(imgs).cuda())                        # [B, H, W, 3] tensor
    assert torch.equal(out, out4)
    for i, im in enumerate(imgs):
        assert np.array_equal(u8[i].cpu().numpy(), X.pil_resize_ref(im.numpy(), 56, 56))
        assert torch.equal(out[i].cpu(), X.preprocess_ref(im.numpy(), 56, 56))
        one = pre([im])
        assert torch.equal(one[0], out[i])


# ------------------------------------------------------------------ BASELINE configs[0] end to end
def test_config1_images_to_descriptor_pipeline_against_cpu_oracle(vpr):
    """BASELINE configs[0]: DINOv2 ViT-B/14 + SALAD forward at 224x224, batch 8, random-init weights.  GPU pipeline:
    GpuPreprocess -> HuggingFace Dinov2Model blocks (the backbone is outside the path; fp32 here) -> BackboneSeam ->
    SaladHead -> regressor.  CPU oracle: torchvision-equivalent preprocessing restatement -> the same HF model incl. its
    final LayerNorm -> SaladRef -> LatLonRegressorRef."""
    transformers = pytest.importorskip("transformers")
    torch.manual_seed(0)
    cfg = transformers.Dinov2Config(hidden_size=768, num_hidden_layers=12, num_attention_heads=12, mlp_ratio=4, patch_size=14,
                                    image_size=224)
    backbone = transformers.Dinov2Model(cfg).eval()
    ref_head = O.SaladRef(dropout=0.3).eval()
    ref_reg = O.LatLonRegressorRef(8448, 512).eval()
    rng = np.random.RandomState(12)
    raw = [rng.randint(0, 256, (300, 400, 3)).astype(np.uint8) for _ in range(8)]
    # ---- CPU oracle
    with torch.no_grad():
        x_cpu = torch.stack([X.preprocess_ref(im, 224, 224) for im in raw])
        hs = backbone(pixel_values=x_cpu).last_hidden_state                       # final LayerNorm applied by HF
        f = hs[:, 1:].reshape(8, 16, 16, 768).permute(0, 3, 1, 2)
        want_desc = ref_head((f, hs[:, 0]))
        want_pred = ref_reg(want_desc)
    # ---- GPU pipeline
    class Blocks(torch.nn.Module):                        # the backbone without its final LayerNorm
        def __init__(self, m):
            super().__init__()
            self.m = m

        def forward(self, px):
            return self.m(pixel_values=px, output_hidden_states=True).hidden_states[-1]

    seam = vpr.BackboneSeam(768, eps=cfg.layer_norm_eps, out_dtype=torch.float32)
    seam.norm.load_state_dict(backbone.layernorm.state_dict())
    head = vpr.SaladHead(dropout=0.3)
    head.load_state_dict(ref_head.state_dict())
    model = vpr.DINOv2RegressionModel(vpr.VPRModel(Blocks(backbone), head, seam)).cuda().eval()
    model.regressor.load_state_dict(ref_reg.regressor.state_dict())
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        pre = vpr.GpuPreprocess((224, 224))
        px = pre([torch.from_numpy(im) for im in raw])
        assert torch.equal(px.cpu(), x_cpu)                                          # bit-identical inputs
        desc = model.features(px)
        pred = model(px)
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev
    assert desc.shape == (8, 8448)
    assert rel(desc, want_desc) < 1e-3                      # fp32 backbone on two devices: summation-order noise only
    assert rel(pred, want_pred) < 1e-3
