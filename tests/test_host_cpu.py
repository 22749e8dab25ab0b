"""CPU-side tests: the C-ABI library loads and exports everything include/vpr_b200.h declares,
the host-side module mirrors the upstream interface, there is no CPU fallback, and the
multi-process plumbing works under gloo (world_size 2)."""
import ctypes as C
import os
import re
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as g
    g.build()
    import vpr_b200
    return vpr_b200


def test_library_exports_every_declared_symbol(built):
    hdr = open(os.path.join(ROOT, "include", "vpr_b200.h")).read()
    declared = sorted(set(re.findall(r"VPR_API [^;]*?\b(vpr_[a-z0-9_]+)\s*\(", hdr)))
    assert len(declared) >= 16
    handle = C.CDLL(built.LIB_PATH)
    for name in declared:
        assert hasattr(handle, name), f"{name} declared in include/vpr_b200.h but not exported"
    from vpr_b200 import _lib
    assert sorted(_lib.EXPORTS) == declared, "ctypes binding and header disagree"
    assert built.lib().vpr_abi_version() == _lib.ABI_VERSION


def test_buffer_size_queries_need_no_gpu(built):
    from vpr_b200 import _lib as L
    lib = built.lib()
    s = L.SaladShape(256, 256, 768, 512, 64, 128, 256, 3)
    save_bf16 = lib.vpr_salad_save_bytes(C.byref(s), L.VPR_BF16)
    save_f32 = lib.vpr_salad_save_bytes(C.byref(s), L.VPR_F32)
    T = 256 * 256
    assert save_bf16 >= T * 1024 * 2 + T * 128 * 2 + T * 64 * 4          # H + F + S at least
    assert save_f32 > save_bf16
    assert lib.vpr_salad_workspace_bytes(C.byref(s), L.VPR_BF16) >= T * 1024 * 2
    assert lib.vpr_knn_workspace_bytes(65536, 1250000, 8448, 25) > 65536 * 25 * 8
    hs = L.HeadsShape(256, 8448, 512, 0, 1, 1.0, 1.0, 256, 0.0, 0)
    assert lib.vpr_heads_workspace_bytes(C.byref(hs), L.VPR_BF16) >= 520 * 8448 * 2


def test_salad_head_mirrors_upstream_interface(built):
    m = built.SaladHead()       # hub entry passes num_channels=768; same default here
    sd = m.state_dict()
    expect = {
        "dust_bin": (), "token_features.0.weight": (512, 768), "token_features.0.bias": (512,),
        "token_features.2.weight": (256, 512), "token_features.2.bias": (256,),
        "cluster_features.0.weight": (512, 768, 1, 1), "cluster_features.0.bias": (512,),
        "cluster_features.3.weight": (128, 512, 1, 1), "cluster_features.3.bias": (128,),
        "score.0.weight": (512, 768, 1, 1), "score.0.bias": (512,),
        "score.3.weight": (64, 512, 1, 1), "score.3.bias": (64,),
    }
    assert {k: tuple(v.shape) for k, v in sd.items()} == expect
    assert sum(p.numel() for p in m.parameters()) == 1411009
    assert m.out_dim == 8448                      # reference dinov2salad_finetuning.py:29
    from oracle.salad_ref import SaladRef
    m.load_state_dict(SaladRef().state_dict())    # interchangeable checkpoints


def test_reference_checkpoint_prefix_loads(built):
    """Reference checkpoints keep SALAD under feature_extractor.aggregator.* next to regressor.*
    (dinov2salad_finetuning.py:24,28-32,132)."""
    from oracle.salad_ref import LatLonRegressorRef, SaladRef
    ck = {"feature_extractor.aggregator." + k: v for k, v in SaladRef().state_dict().items()}
    ck.update({k: v for k, v in LatLonRegressorRef().state_dict().items()})
    head = built.SaladHead()
    heads = built.GeoposeHeads()
    pre = "feature_extractor.aggregator."
    head.load_state_dict({k[len(pre):]: v for k, v in ck.items() if k.startswith(pre)})
    missing = heads.load_state_dict({k: v for k, v in ck.items() if k.startswith("regressor.")}, strict=False)
    assert set(missing.missing_keys) == {"head.weight", "head.bias"} and not missing.unexpected_keys


def test_no_cpu_fallback(built):
    m = built.SaladHead()
    x = torch.randn(2, 768, 16, 16)
    t = torch.randn(2, 768)
    with pytest.raises(RuntimeError, match="CUDA"):
        m((x, t))
    with pytest.raises(RuntimeError, match="CUDA"):
        built.knn_topk(torch.randn(4, 64).bfloat16(), torch.randn(9, 64).bfloat16(), 3)
    with pytest.raises(RuntimeError, match="CUDA"):
        built.GeoposeHeads()(torch.randn(2, 8448), torch.randn(2, 2), torch.rand(2))


def test_next_rows_have_no_cpu_fallback_either(built):
    with pytest.raises(RuntimeError, match="CUDA"):
        built.BackboneSeam()(torch.randn(2, 17, 768))
    with pytest.raises(RuntimeError, match="CUDA"):
        built.latlon_metrics(torch.randn(4, 2), torch.randn(4, 2))
    with pytest.raises(RuntimeError, match="CUDA"):
        built.maae(torch.randn(4, 2), torch.rand(4))
    with pytest.raises(RuntimeError, match="CUDA"):
        built.recall_at_k(torch.zeros(4, 3, dtype=torch.int64), torch.zeros(4), torch.zeros(9))
    model = built.DINOv2RegressionModel(torch.nn.Identity()).eval()
    with pytest.raises(RuntimeError, match="CUDA"):
        model(torch.randn(2, 8448))
    with pytest.raises(RuntimeError, match="CUDA"):
        built.GpuPreprocess(device="cpu")([torch.zeros(8, 8, 3, dtype=torch.uint8)])
    assert built.lib().vpr_preprocess_workspace_bytes(480, 640, 224, 224) >= 480 * 224 * 3


def test_graphed_steps_and_numa_helper_without_a_gpu(built):
    head, heads = built.SaladHead(), built.GeoposeHeads()
    with pytest.raises(RuntimeError, match="CUDA"):
        built.GraphedSaladStep(head, batch=2, n_tokens=256)
    with pytest.raises(RuntimeError, match="CUDA"):
        built.GraphedFinetuneStep(head, heads, batch=2, n_tokens=256)
    assert built.bind_to_gpu_numa_node(0) is None            # no device, no sysfs topology: a no-op, never an exception


def test_scaler_mirrors_sklearn_and_joblib(built, tmp_path):
    import joblib
    from sklearn.preprocessing import StandardScaler
    rng = np.random.default_rng(0)
    latlon = np.stack([17.44 + 0.01 * rng.standard_normal(400), 78.35 + 0.02 * rng.standard_normal(400)], 1)
    sk = StandardScaler().fit(latlon)                      # dinov2salad_finetuning.py:79-81
    mine = built.LatLonScaler.fit(latlon)
    np.testing.assert_allclose(mine.mean_, sk.mean_, rtol=1e-13)
    np.testing.assert_allclose(mine.scale_, sk.scale_, rtol=1e-13)
    z = rng.standard_normal((5, 2))
    np.testing.assert_allclose(mine.inverse_transform(z), sk.inverse_transform(z), rtol=1e-13)
    np.testing.assert_allclose(mine.transform(latlon[:5]), sk.transform(latlon[:5]), rtol=1e-9, atol=1e-9)
    path = tmp_path / "latlon_scaler_x.pkl"
    joblib.dump(sk, path)                                  # dinov2salad_finetuning.py:110-111
    loaded = built.LatLonScaler.from_joblib(str(path))     # dinov2salad_validation.py:73
    np.testing.assert_array_equal(loaded.mean_, sk.mean_)
    np.testing.assert_allclose(loaded.to_sklearn().inverse_transform(z), sk.inverse_transform(z), rtol=1e-13)
    t = torch.from_numpy(z)
    np.testing.assert_allclose(mine.inverse_transform(t).numpy(), sk.inverse_transform(z), rtol=1e-13)


def test_reference_model_wrapper_and_checkpoint_layout(built, tmp_path):
    """dinov2salad_finetuning.py:21-37,130-135 / dinov2salad_validation.py:68-69 key layout, on CPU."""
    from oracle.salad_ref import LatLonRegressorRef, SaladRef
    torch.manual_seed(2)
    ref, reg = SaladRef(), LatLonRegressorRef(8448, 512)
    model = built.DINOv2RegressionModel(built.VPRModel(torch.nn.Identity(), built.SaladHead()))
    assert all(not p.requires_grad for p in model.feature_extractor.parameters())
    assert all(p.requires_grad for p in model.regressor.parameters())
    keys = set(model.state_dict())
    assert {"regressor.0.weight", "regressor.0.bias", "regressor.2.weight", "regressor.2.bias",
            "feature_extractor.aggregator.dust_bin", "feature_extractor.aggregator.score.3.weight"} <= keys
    assert {"regressor.0.weight", "regressor.3.weight"} <= set(
        built.DINOv2RegressionModel(torch.nn.Identity(), dropout_prob=0.3).state_dict())   # finetuning_2.py:118-123
    sd = {f"feature_extractor.aggregator.{k}": v.clone() for k, v in ref.state_dict().items()}
    sd["feature_extractor.backbone.model.cls_token"] = torch.zeros(1, 1, 768)
    sd.update({f"regressor.{k}": v.clone() for k, v in reg.regressor.state_dict().items()})
    path = tmp_path / "checkpoint_7_.pth"
    torch.save({"epoch": 7, "model_state_dict": sd, "optimizer_state_dict": {}, "loss": torch.tensor(0.5)}, path)
    info = built.load_reference_checkpoint(model, str(path))
    assert info["epoch"] == 7 and info["skipped"] == ["feature_extractor.backbone.model.cls_token"]
    for k, v in ref.state_dict().items():
        assert torch.equal(model.feature_extractor.aggregator.state_dict()[k], v)
    assert torch.equal(model.regressor[2].bias, reg.regressor[2].bias)
    bad = dict(sd)
    del bad["regressor.2.bias"]
    with pytest.raises(RuntimeError):
        built.load_reference_checkpoint(model, {"model_state_dict": bad})
    out = tmp_path / "checkpoint_8_.pth"
    built.save_reference_checkpoint(str(out), model, None, 8, 0.25)
    back = torch.load(out, weights_only=False)
    assert set(back) == {"epoch", "model_state_dict", "optimizer_state_dict", "loss"} and back["epoch"] == 8
    SaladRef().load_state_dict({k[len("feature_extractor.aggregator."):]: v for k, v in back["model_state_dict"].items()
                                if k.startswith("feature_extractor.aggregator.")})
    head = built.SaladHead()
    flat = {f"aggregator.{k}": (v.flatten(1) if v.dim() == 4 else v) for k, v in ref.state_dict().items()}
    built.load_upstream_salad_weights(head, {"state_dict": flat})      # Linear-shaped conv weights are accepted too
    assert torch.equal(head.state_dict()["score.3.weight"], ref.state_dict()["score.3.weight"])
    with pytest.raises(KeyError):
        built.load_upstream_salad_weights(head, {"state_dict": {"backbone.x": torch.zeros(1)}})


def test_single_thread_roles_issue_from_uniform_registers(built):
    """Regression guard for DESIGN.md 4.1: a TMA / tcgen05 instruction issued from inside `if (lane == 0) { loop }` gets
    wrapped by the compiler in an election loop (ELECT, R2UR.BROADCAST x 5-7, branch) because its operands must be uniform
    registers -- ~100 serial issue slots per k-block, which used to make the MMA issuer co-critical with the tensor pipe
    (0.414 -> 0.391 ms/step, kNN 55 k -> 60 k queries/s when they went).  R2UR.BROADCAST appears nowhere else, so: none
    of it in the kernels that run by default.  (sink_agg_fwd.cu is opt-in and keeps its single issuing lane.)"""
    import shutil
    import subprocess
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    obj_dir = os.path.join(ROOT, "visual-place-recognition-and-geopose-estimation_b200", "build")
    for name in ("gemm", "knn", "dh_fused", "fs_fwd", "agg_fwd", "agg_bwd"):
        obj = os.path.join(obj_dir, name + ".o")
        if not os.path.exists(obj):
            pytest.skip("object files not kept on this box (prebuilt library)")
        sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True, check=True).stdout
        assert "UTCHMMA" in sass, name                       # the tensor-core path is there at all
        assert sass.count("R2UR.BROADCAST") == 0, f"{name}.cu: a single-thread role issues from a divergent region again"


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "visual-place-recognition-and-geopose-estimation_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert "oracle" not in src, f"{f} mentions the oracle"


def test_tokens_from_map_is_a_view_for_channels_last(built):
    x = torch.randn(2, 16 * 16, 768)
    f = x.view(2, 16, 16, 768).permute(0, 3, 1, 2)       # upstream's backbone output view
    tok = built.SaladHead.tokens_from_map(f)
    assert tok.data_ptr() == x.data_ptr() and tok.shape == (2, 256, 768)
    f2 = torch.randn(2, 768, 16, 16)                      # genuinely NCHW: copied, same values
    tok2 = built.SaladHead.tokens_from_map(f2)
    assert torch.equal(tok2, f2.flatten(2).transpose(1, 2))


def test_shard_range_partitions_exactly(built):
    for total in (0, 1, 7, 256, 2048, 10_000_000):
        for world in (1, 2, 3, 8):
            spans = [built.shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1


def _gloo_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    import vpr_b200
    r, _, w = vpr_b200.init_distributed("gloo")
    torch.manual_seed(0)
    lin = torch.nn.Sequential(torch.nn.Linear(300, 200), torch.nn.Linear(200, 7))
    for i, p in enumerate(lin.parameters()):
        p.grad = torch.full_like(p, float(rank + 1) * (i + 1))
    fin = vpr_b200.allreduce_grads(lin.parameters(), bucket_bytes=64 << 10, async_op=True)   # several buckets
    fin()
    ok = all(torch.allclose(p.grad, torch.full_like(p, 3.0 * (i + 1))) for i, p in enumerate(lin.parameters()))
    # gradients that are back-to-back views of one buffer (what SaladHead's backward produces) are reduced in
    # place by a single collective: same storage afterwards, summed values
    ps = list(lin.parameters())
    flat = torch.full((sum(p.numel() for p in ps),), float(rank + 1))
    off = 0
    for p in ps:
        p.grad = flat[off:off + p.numel()].view_as(p)
        off += p.numel()
    from vpr_b200.parallel import _shared_flat_view
    fv = _shared_flat_view([p.grad for p in ps])
    ok = ok and fv is not None and fv.data_ptr() == flat.data_ptr() and fv.numel() == flat.numel()
    vpr_b200.allreduce_grads(ps)
    ok = ok and bool((flat == 3.0).all()) and all(p.grad.data_ptr() >= flat.data_ptr() for p in ps)
    ok = ok and _shared_flat_view([ps[0].grad, ps[2].grad]) is None          # a gap -> bucketed path
    # sharded top-k merge: allgather of per-rank lists, merged with the oracle's merge on CPU
    from oracle.salad_ref import knn_topk_ref, merge_topk
    g = torch.Generator().manual_seed(1)
    G = torch.randn(1000, 32, generator=g).bfloat16()
    Q = torch.randn(5, 32, generator=g).bfloat16()
    b, e = vpr_b200.shard_range(1000, rank, world)
    s, i = knn_topk_ref(Q, G[b:e], 25, g_offset=b)
    all_s = [torch.empty_like(s) for _ in range(world)]
    all_i = [torch.empty_like(i) for _ in range(world)]
    dist.all_gather(all_s, s)
    dist.all_gather(all_i, i)
    ms, mi = merge_topk(torch.cat(all_s, 1), torch.cat(all_i, 1), 25)
    fs, fi = knn_topk_ref(Q, G, 25)
    ok = ok and torch.equal(mi, fi) and torch.equal(ms, fs)
    q.put((rank, bool(ok)))
    dist.barrier()
    dist.destroy_process_group()


def test_gloo_world2_allreduce_and_sharded_merge(built):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29000 + os.getpid() % 2000
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(60)
    assert sorted(res) == [(0, True), (1, True)]


def test_prediction_csv_writers_reproduce_the_reference_files(built, tmp_path, golden_dir):
    """f3: the per-image CSV files of swin_transformer/validation_swin_2.py:205-218 and
    angle_prediction/efficient_net/validation_script.py:211-220.  Golden = the first rows of two files the reference
    repository itself checked in (swin_transformer/training_gemini_2_20250505_004059/validation_predictions.csv,
    angle_prediction/efficientnet/sincos/training_20250505_191535/validation_predictions.csv): re-writing their
    values must give the same bytes; pandas (what the reference calls) must agree on fresh values too."""
    import csv
    import pandas as pd
    want = open(os.path.join(golden_dir, "latlon_predictions_head.csv")).read()
    rows = list(csv.reader(want.splitlines()))[1:]
    names = [r[0] for r in rows]
    true = np.array([[float(r[1]), float(r[2])] for r in rows])
    pred = torch.tensor([[float(r[3]), float(r[4])] for r in rows], dtype=torch.float64)
    out = tmp_path / "val.csv"
    assert built.write_latlon_predictions_csv(out, names, true, pred) == len(rows)
    assert out.read_text() == want
    want_a = open(os.path.join(golden_dir, "angle_predictions_head.csv")).read()
    rows = list(csv.reader(want_a.splitlines()))[1:]
    out_a = tmp_path / "ang.csv"
    built.write_angle_predictions_csv(out_a, [r[0] for r in rows], [int(r[1]) for r in rows],
                                      torch.tensor([float(r[2]) for r in rows], dtype=torch.float32))
    assert out_a.read_text() == want_a
    # fresh values against pandas itself, including a name that needs quoting
    rng = np.random.default_rng(3)
    t2, p2 = rng.normal(2e5, 1e3, (9, 2)), rng.normal(2e5, 1e3, (9, 2)).astype(np.float32)
    nm = [f"img_{i:04d}.jpg" for i in range(8)] + ['we,ird "name".jpg']
    built.write_latlon_predictions_csv(out, nm, t2, p2)
    df = pd.DataFrame({"filename": nm, "true_latitude": t2[:, 0], "true_longitude": t2[:, 1],
                       "predicted_latitude": p2[:, 0].astype(np.float64), "predicted_longitude": p2[:, 1].astype(np.float64)})
    df["error_latitude"] = np.abs(df["true_latitude"] - df["predicted_latitude"])
    df["error_longitude"] = np.abs(df["true_longitude"] - df["predicted_longitude"])
    ref = tmp_path / "ref.csv"
    df.to_csv(ref, index=False, float_format="%.6f")
    assert out.read_text() == ref.read_text()
    ta, pa = rng.integers(0, 360, 9).tolist(), (rng.random(9) * 360).astype(np.float32)
    built.write_angle_predictions_csv(out_a, nm, ta, torch.from_numpy(pa))
    df = pd.DataFrame({"filename": nm, "true_angle": ta, "predicted_angle": [float(v) for v in pa]})
    df["angular_error"] = df.apply(lambda row: min(abs(row["predicted_angle"] - row["true_angle"]),
                                                   360 - abs(row["predicted_angle"] - row["true_angle"])), axis=1)
    df.to_csv(ref, index=False)
    assert out_a.read_text() == ref.read_text()
    with pytest.raises(ValueError):
        built.write_latlon_predictions_csv(out, nm[:3], t2, p2)


def test_heads_variants_mirror_the_reference_modules(built):
    """a9: the scalar-angle head is Linear(D,1) (dino_angle_finetuning.py:41), the sin/cos head Linear(D,2) behind a
    Dropout (dino_v2_gemini.py:104-106); state_dict keys as in the reference (`head.weight`, `head.bias`)."""
    h1 = built.GeoposeHeads(768, 64, angle_loss="scalar_maae")
    assert tuple(h1.head.weight.shape) == (1, 768)
    h2 = built.GeoposeHeads(768, 64, angle_loss="atan2_maae", head_dropout=0.1)
    assert tuple(h2.head.weight.shape) == (2, 768) and isinstance(h2.dropout, torch.nn.Dropout) and h2.dropout.p == 0.1
    assert {"head.weight", "head.bias"} <= set(h2.state_dict()) and not any(k.startswith("dropout") for k in h2.state_dict())
    with pytest.raises(ValueError):
        built.GeoposeHeads(768, 64, angle_loss="nope")


def test_decorrelate_dropout_gives_every_rank_its_own_stream(built):
    seeds = set()
    for r in range(8):
        m = built.SaladHead()
        built.decorrelate_dropout(m, r)
        built.decorrelate_dropout(m, r)            # idempotent
        seeds.add(m.seed)
    assert len(seeds) == 8
