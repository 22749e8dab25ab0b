"""Seam kernel (LayerNorm + CLS/patch split + cast) against the HBM roofline.
    python tools/seam_bench.py [--B 256 --n 256 --d 768 --iters 20]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vpr_b200  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=256)
    ap.add_argument("--n", type=int, default=256)
    ap.add_argument("--d", type=int, default=768)
    ap.add_argument("--iters", type=int, default=40)
    ap.add_argument("--sets", type=int, default=4)
    a = ap.parse_args()
    peaks = {}
    pk = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peaks = json.load(open(pk))
    for in_dt, out_dt in ((torch.bfloat16, torch.bfloat16), (torch.float32, torch.bfloat16)):
        from vpr_b200 import _lib as L
        lib = L.lib()
        rows = a.B * (1 + a.n)
        seam = vpr_b200.BackboneSeam(a.d, out_dtype=out_dt).cuda()
        w, b = seam.norm.weight.detach(), seam.norm.bias.detach()
        st = L.stream_ptr(torch.device("cuda"))
        di, do = L.dtype_code(in_dt), L.dtype_code(out_dt)
        sets = []
        for _ in range(a.sets):       # rotating working sets (each launch's operands are not in L2 from the previous one)
            hidden = torch.randn(a.B, 1 + a.n, a.d, device="cuda").to(in_dt)
            x = torch.empty(a.B, a.n, a.d, dtype=out_dt, device="cuda")
            t = torch.empty(a.B, a.d, dtype=out_dt, device="cuda")
            sets.append((hidden, x, t, torch.empty(rows, 2, device="cuda"), torch.randn_like(x), torch.randn_like(t),
                         torch.empty_like(hidden)))
        dw, db = torch.empty(a.d, device="cuda"), torch.empty(a.d, device="cuda")

        def fwd(i):
            hidden, x, t, stats, gx, gt, dh = sets[i % a.sets]
            L.check(lib.vpr_seam_ln_split_fwd(L.ptr(hidden), L.ptr(w), L.ptr(b), 1e-6, L.ptr(x), L.ptr(t), L.ptr(stats), a.B, a.n,
                                              0, a.d, di, do, st), "fwd")

        def bwd(i):
            hidden, x, t, stats, gx, gt, dh = sets[i % a.sets]
            L.check(lib.vpr_seam_ln_split_bwd(L.ptr(hidden), L.ptr(w), L.ptr(stats), L.ptr(gx), L.ptr(gt), L.ptr(dh), L.ptr(dw),
                                              L.ptr(db), a.B, a.n, 0, a.d, di, do, st), "bwd")
        esz_in, esz_out = sets[0][0].element_size(), sets[0][1].element_size()
        for name, fn, byt in (("seam fwd", fwd, rows * a.d * (esz_in + esz_out) + rows * 8),
                              ("seam bwd", bwd, rows * a.d * (2 * esz_in + esz_out) + rows * 8)):
            for i in range(a.sets):
                fn(i)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(a.iters):
                fn(i)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / a.iters
            print(json.dumps({"kernel": name, "in": str(in_dt), "out": str(out_dt), "ms": ms, "GB": byt / 1e9,
                              "GB/s": byt / ms / 1e6, "frac_of_measured_hbm": byt / ms / 1e6 / peaks.get("hbm_gbs", 6549.8),
                              "timing": f"{a.iters} back-to-back launches over {a.sets} rotating working sets"}))


if __name__ == "__main__":
    main()
