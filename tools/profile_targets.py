"""Small driver for `ncu --set full -k regex:...`: two eager SALAD fwd+bwd steps (B=256, n=256, bf16, dropout 0.3) and two
seam forward/backward launches, so that the second instance of every kernel is a warm one.
    ncu --set full --clock-control none --import-source on -k regex:"agg_|seam_" -o gpurun_out/prof_x python tools/profile_targets.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import vpr_b200  # noqa: E402


def main():
    B, n = 256, 256
    torch.manual_seed(0)
    head = vpr_b200.SaladHead(dropout=0.3).cuda().train()
    x = torch.randn(B, n, 768, device="cuda").bfloat16().requires_grad_(True)
    t = torch.randn(B, 768, device="cuda").bfloat16().requires_grad_(True)
    w = torch.randn(B, 8448, device="cuda").bfloat16()
    for _ in range(2):
        out = head((x, t))
        out.backward(w)
        x.grad = None
        t.grad = None
    seam = vpr_b200.BackboneSeam(768, out_dtype=torch.bfloat16).cuda()
    hidden = torch.randn(B, 1 + n, 768, device="cuda").bfloat16().requires_grad_(True)
    for _ in range(2):
        xt, tt = seam(hidden)
        torch.autograd.backward([xt, tt], [torch.ones_like(xt), torch.ones_like(tt)])
        hidden.grad = None
    # geopose heads on the descriptor batch (fused forward + backward call), twice
    heads = vpr_b200.GeoposeHeads(latlon_loss="mse", angle_loss="atan2_maae").cuda().train()
    desc = out.detach()
    ll = torch.randn(B, 2, device="cuda")
    ang = torch.rand(B, device="cuda") * 360
    for _ in range(2):
        d = desc.clone().requires_grad_(True)
        total, *_ = heads(d, ll, ang)
        total.backward()
    torch.cuda.synchronize()
    print("done")


if __name__ == "__main__":
    main()
