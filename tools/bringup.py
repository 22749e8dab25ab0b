"""GPU bring-up diagnostics: runs one group of checks per process (a device trap in one group
must not hide the others) and prints error statistics instead of asserting.

    python tools/bringup.py all            # spawns every group in its own process
    python tools/bringup.py gemm_kk        # one group
"""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

GROUPS = ["gemm_kk", "gemm_kmn", "gemm_mnk", "gemm_mnmn", "gemm_misc", "gemm_f32", "sinkhorn", "salad_f32",
          "salad_bf16", "heads", "knn"]


def rel(a, b):
    import torch
    a, b = a.float(), b.float()
    return ((a - b).norm() / (b.norm() + 1e-30)).item(), (a - b).abs().max().item()


def run_gemm(a_mn, b_mn, M, N, K, batch=1, bn=0, k_splits=1, out_kind=0, bias=False, relu=False, mask=False,
             dtype="bf16", tag=""):
    import torch
    import vpr_b200
    from vpr_b200 import ops, _lib as L
    dt = torch.bfloat16 if dtype == "bf16" else torch.float32
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N * 3 + K)
    Af = torch.randn(batch, M, K, device="cuda", generator=g)
    Bf = torch.randn(batch, N, K, device="cuda", generator=g)
    A = Af.to(dt)
    B = Bf.to(dt)
    ref = torch.einsum("bmk,bnk->bmn", A.float(), B.float())
    Ast = A.transpose(1, 2).contiguous() if a_mn else A.contiguous()
    Bst = B.transpose(1, 2).contiguous() if b_mn else B.contiguous()
    bias_t = torch.randn(N, device="cuda", generator=g) if bias else None
    mask_t = torch.randn(batch, M, N, device="cuda", generator=g).to(dt) if mask else None
    if bias:
        ref = ref + bias_t
    if relu:
        ref = ref.clamp_min(0)
    if mask:
        ref = torch.where(mask_t.float() > 0, ref * 0.5, torch.zeros_like(ref))
    out = ops.gemm(Ast, Bst, M, N, K, a_mn=a_mn, b_mn=b_mn, batch=batch, a_bs=Ast.stride(0) if batch > 1 else 0,
                   b_bs=Bst.stride(0) if batch > 1 else 0, out_kind=out_kind, bias=bias_t, relu=relu, mask=mask_t,
                   mask_scale=0.5, k_splits=k_splits, bn=bn)
    torch.cuda.synchronize()
    r, mx = rel(out, ref)
    print(f"  gemm{tag} a_mn={int(a_mn)} b_mn={int(b_mn)} M={M} N={N} K={K} batch={batch} bn={bn} ks={k_splits} "
          f"out={out_kind} bias={int(bias)} relu={int(relu)} mask={int(mask)} {dtype}: rel={r:.3e} max={mx:.3e}",
          flush=True)
    return r


def group_gemm(a_mn, b_mn):
    for (M, N, K, bn) in [(128, 64, 64, 64), (128, 128, 64, 128), (128, 256, 64, 256), (128, 256, 128, 256),
                          (256, 512, 256, 0), (384, 192, 320, 0), (200, 100, 72, 0), (1024, 1024, 768, 0)]:
        run_gemm(a_mn, b_mn, M, N, K, bn=bn)
    run_gemm(a_mn, b_mn, 256, 128, 192, batch=3)
    run_gemm(a_mn, b_mn, 529, 64, 128, batch=2)


def group_gemm_misc():
    run_gemm(False, False, 256, 256, 512, out_kind=1, bias=True, relu=True)
    run_gemm(False, True, 300, 512, 128, out_kind=1, mask=True)
    run_gemm(True, True, 128, 512, 4096, out_kind=2, k_splits=16)
    run_gemm(True, True, 64, 512, 2048, out_kind=2, k_splits=8)
    run_gemm(True, True, 512, 768, 8192, out_kind=2, k_splits=12)
    run_gemm(False, False, 2, 512, 768, bias=True, relu=True, out_kind=1)
    run_gemm(True, True, 256, 512, 2)
    run_gemm(False, False, 4096, 1024, 768, out_kind=1, bias=True, relu=True)


def group_gemm_f32():
    for a_mn in (False, True):
        for b_mn in (False, True):
            run_gemm(a_mn, b_mn, 200, 100, 72, dtype="f32")
            run_gemm(a_mn, b_mn, 128, 64, 529, batch=2, dtype="f32")
    run_gemm(True, True, 128, 512, 4096, out_kind=2, k_splits=16, dtype="f32")
    run_gemm(False, True, 300, 512, 128, mask=True, bias=True, relu=True, dtype="f32")


def group_sinkhorn():
    import torch
    import vpr_b200
    from oracle import salad_ref as O
    for (B, n, m) in [(2, 256, 64), (3, 529, 64), (2, 100, 64), (2, 300, 32)]:
        g = torch.Generator().manual_seed(n)
        S = torch.randn(B, m, n, generator=g) * 2.0           # oracle layout [B, m, n]
        db = torch.tensor(1.0)
        logP = O.get_matching_probs(S.double(), db.double(), 3, const_dtype=torch.float64)
        Pref = logP.exp()[:, :m].transpose(1, 2)               # [B, n, m]
        St = S.transpose(1, 2).contiguous().cuda()
        P, U, V = vpr_b200.sinkhorn(St, db.cuda(), 3)
        torch.cuda.synchronize()
        r, mx = rel(P.cpu().double(), Pref)
        print(f"  sinkhorn fwd B={B} n={n} m={m}: rel={r:.3e} max={mx:.3e}", flush=True)
        dP = torch.randn(B, n, m, generator=g)
        Sd = S.double().requires_grad_(True)
        Pd = O.get_matching_probs(Sd, db.double(), 3, const_dtype=torch.float64).exp()[:, :m]
        (Pd * dP.double().transpose(1, 2)).sum().backward()
        dSref = Sd.grad.transpose(1, 2)
        dS, dbias = vpr_b200.sinkhorn_backward(St, db.cuda(), U, V, dP.cuda(), 3)
        torch.cuda.synchronize()
        r, mx = rel(dS.cpu().double(), dSref)
        r2, _ = rel(dbias.cpu().double(), dSref.sum((0, 1)))
        print(f"  sinkhorn bwd B={B} n={n} m={m}: rel={r:.3e} max={mx:.3e} dbias rel={r2:.3e}", flush=True)


def salad_compare(dtype, B, hw, seed=1234, dropout=0.0, train=False):
    import torch
    import vpr_b200
    from oracle import salad_ref as O
    torch.manual_seed(0)
    ref = O.SaladRef(dropout=dropout)
    mod = vpr_b200.SaladHead(dropout=dropout).cuda()
    mod.load_state_dict(ref.state_dict())
    ref.train(train)
    mod.train(train)
    g = torch.Generator().manual_seed(seed)
    n = hw * hw
    x = torch.randn(B, n, 768, generator=g)
    t = torch.randn(B, 768, generator=g)
    w = torch.randn(B, 8448, generator=g)
    dt = torch.bfloat16 if dtype == "bf16" else torch.float32
    if dtype == "bf16":   # oracle sees the same rounded inputs
        x = x.bfloat16().float()
        t = t.bfloat16().float()
    xg = x.cuda().to(dt).requires_grad_(True)
    tg = t.cuda().to(dt).requires_grad_(True)
    out = mod((O.tokens_to_salad_input(xg, hw, hw), tg))
    (out.float() * w.cuda()).sum().backward()
    torch.cuda.synchronize()
    mask = None
    if train and dropout > 0:
        keep = vpr_b200.dropout_keep_mask(mod.last_seed, dropout, B * n, 1024).cpu().float()
        mask = keep.reshape(B, hw, hw, 1024).permute(0, 3, 1, 2)
    xr = x.clone().requires_grad_(True)
    tr = t.clone().requires_grad_(True)
    oref = ref((O.tokens_to_salad_input(xr, hw, hw), tr), drop_mask=mask)
    (oref * w).sum().backward()
    print(f"  salad {dtype} B={B} n={n} dropout={dropout} train={train}", flush=True)
    r, mx = rel(out.detach().cpu(), oref.detach()); print(f"    desc        rel={r:.3e} max={mx:.3e}")
    r, mx = rel(xg.grad.cpu(), xr.grad); print(f"    dx          rel={r:.3e} max={mx:.3e}")
    r, mx = rel(tg.grad.cpu(), tr.grad); print(f"    dt          rel={r:.3e} max={mx:.3e}")
    for (name, p), (_, pr) in zip(mod.named_parameters(), ref.named_parameters()):
        r, mx = rel(p.grad.cpu(), pr.grad)
        print(f"    g {name:28s} rel={r:.3e} max={mx:.3e} |ref|={pr.grad.norm().item():.3e}")
    sys.stdout.flush()


def group_heads():
    import torch
    import vpr_b200
    from oracle import salad_ref as O
    for dtype in ("f32", "bf16"):
        for latlon, angle in (("mse", "atan2_maae"), ("huber", "acos"), ("mse", "mse"), ("mse", None)):
            torch.manual_seed(3)
            B, D = 48, 8448
            ll = O.LatLonRegressorRef(D, 512)
            ah = O.SinCosHeadRef(D)
            heads = vpr_b200.GeoposeHeads(D, 512, latlon_loss=latlon, angle_loss=angle).cuda()
            heads.regressor.load_state_dict(ll.regressor.state_dict())
            heads.head.load_state_dict(ah.head.state_dict())
            g = torch.Generator().manual_seed(5)
            desc = torch.nn.functional.normalize(torch.randn(B, D, generator=g), dim=1) * 8.0
            lt = torch.randn(B, 2, generator=g)
            at = torch.rand(B, generator=g) * 720.0
            dt = torch.bfloat16 if dtype == "bf16" else torch.float32
            if dtype == "bf16":
                desc = desc.bfloat16().float()
            dg = desc.cuda().to(dt).requires_grad_(True)
            total, losses, llp, scp = heads(dg, lt.cuda(), at.cuda() if angle else None)
            total.backward()
            torch.cuda.synchronize()
            dr = desc.clone().requires_grad_(True)
            if angle:
                tot_r, l1, l2, llr, scr = O.geopose_loss(dr, ll, ah, lt, at, latlon, angle)
            else:
                llr = ll(dr); tot_r = O.mse_loss(llr, lt); l1 = tot_r; l2 = torch.tensor(0.0); scr = None
            tot_r.backward()
            print(f"  heads {dtype} latlon={latlon} angle={angle}: loss {total.item():.6f} vs {tot_r.item():.6f} "
                  f"(l1 {losses[0].item():.6f}/{l1.item():.6f}, l2 {losses[1].item():.6f}/{float(l2):.6f})")
            r, mx = rel(llp.cpu(), llr.detach()); print(f"    latlon_pred rel={r:.3e}")
            if angle:
                r, mx = rel(scp.cpu(), scr.detach()); print(f"    sincos_pred rel={r:.3e}")
            r, mx = rel(dg.grad.cpu(), dr.grad); print(f"    d_desc      rel={r:.3e} max={mx:.3e}")
            pairs = [(heads.regressor[0].weight, ll.regressor[0].weight), (heads.regressor[0].bias, ll.regressor[0].bias),
                     (heads.regressor[2].weight, ll.regressor[2].weight), (heads.regressor[2].bias, ll.regressor[2].bias)]
            if angle:
                pairs += [(heads.head.weight, ah.head.weight), (heads.head.bias, ah.head.bias)]
            for i, (p, pr) in enumerate(pairs):
                r, mx = rel(p.grad.cpu(), pr.grad); print(f"    g[{i}]        rel={r:.3e} max={mx:.3e}")
            sys.stdout.flush()


def group_knn():
    import torch
    import vpr_b200
    from oracle import salad_ref as O
    for (Q, G, D, k) in [(64, 1000, 256, 25), (300, 5000, 8448, 25), (1024, 20000, 512, 10)]:
        g = torch.Generator().manual_seed(Q)
        q = torch.nn.functional.normalize(torch.randn(Q, D, generator=g), dim=1).bfloat16()
        ga = torch.nn.functional.normalize(torch.randn(G, D, generator=g), dim=1).bfloat16()
        s, i = vpr_b200.knn_topk(q.cuda(), ga.cuda(), k, g_offset=1000)
        torch.cuda.synchronize()
        rs, ri = O.knn_topk_ref(q, ga, k, g_offset=1000)
        same = (i.cpu() == ri).float().mean().item()
        sets = sum(set(a.tolist()) == set(b.tolist()) for a, b in zip(i.cpu(), ri)) / Q
        r, mx = rel(s.cpu(), rs)
        print(f"  knn Q={Q} G={G} D={D} k={k}: idx match {same:.4f} set match {sets:.4f} score rel={r:.3e} max={mx:.3e}",
              flush=True)


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which == "all":
        for gname in GROUPS:
            t0 = time.time()
            print(f"=== {gname} ===", flush=True)
            r = subprocess.run([sys.executable, os.path.abspath(__file__), gname], timeout=600)
            print(f"=== {gname}: exit {r.returncode} in {time.time() - t0:.1f}s ===", flush=True)
        return
    if which == "gemm_kk": group_gemm(False, False)
    elif which == "gemm_kmn": group_gemm(False, True)
    elif which == "gemm_mnk": group_gemm(True, False)
    elif which == "gemm_mnmn": group_gemm(True, True)
    elif which == "gemm_misc": group_gemm_misc()
    elif which == "gemm_f32": group_gemm_f32()
    elif which == "sinkhorn": group_sinkhorn()
    elif which == "salad_f32":
        salad_compare("f32", 2, 16)
        salad_compare("f32", 2, 23, seed=4321)
        salad_compare("f32", 3, 16, dropout=0.3, train=True)
    elif which == "salad_bf16":
        salad_compare("bf16", 2, 16)
        salad_compare("bf16", 4, 23, seed=4321)
        salad_compare("bf16", 3, 16, dropout=0.3, train=True)
    elif which == "heads": group_heads()
    elif which == "knn": group_knn()
    else:
        raise SystemExit(f"unknown group {which}")


if __name__ == "__main__":
    main()
