"""Isolated timings of the GEMM shapes the SALAD step is made of (CUDA events, L2 flushed
between launches).  python tools/gemm_bench.py [B]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import vpr_b200  # noqa: E402
from vpr_b200 import _lib as L, ops  # noqa: E402


def timeit(fn, iters=10, warm=3):
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    ts = []
    for i in range(warm + iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        if i >= warm:
            ts.append(e0.elapsed_time(e1))
    return sum(ts) / len(ts)


def main():
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("B", nargs="?", type=int, default=256)
    ap.add_argument("--only", default="")          # comma list of: h,hdrop,dx,dw0,dh,f,cublas
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--warm", type=int, default=3)
    a = ap.parse_args()
    B = a.B
    only = set(x for x in a.only.split(",") if x)
    global timeit
    _t = timeit
    timeit = lambda fn: _t(fn, a.iters, a.warm)  # noqa: E731

    def want(k):
        return not only or k in only
    T = B * 256
    dev = "cuda"
    bf = torch.bfloat16
    X = torch.randn(T, 768, device=dev).to(bf)
    W0 = torch.randn(1024, 768, device=dev).to(bf)
    b0 = torch.randn(1024, device=dev)
    H = torch.empty(T, 1024, device=dev, dtype=bf)
    dH = torch.randn(T, 1024, device=dev).to(bf)
    dX = torch.empty(T, 768, device=dev, dtype=bf)
    dF = torch.randn(T, 128, device=dev).to(bf)
    W1c = torch.randn(128, 512, device=dev).to(bf)
    gW0 = torch.zeros(512, 768, device=dev)
    cs = torch.zeros(512, device=dev)
    rows = []

    def rec(name, ms, flops, bytes_):
        rows.append((name, ms, flops / ms / 1e9, bytes_ / ms / 1e6))
        print(f"{name:44s} {ms*1e3:8.1f} us  {flops/ms/1e9:8.1f} TFLOP/s  {bytes_/ms/1e6:8.1f} GB/s", flush=True)

    f_h = 2.0 * T * 1024 * 768
    if want("h"):
        ms = timeit(lambda: ops.gemm(X, W0, T, 1024, 768, out_kind=L.OUT_BF16, bias=b0, relu=True, out=H, ldo=1024, out_bs=0))
        rec("H = relu(X W0^T + b)            KK  N=1024 K=768", ms, f_h, (T * 768 + T * 1024) * 2)
    if want("hdrop"):
        ms = timeit(lambda: ops.gemm(X, W0, T, 1024, 768, out_kind=L.OUT_BF16, bias=b0, relu=True, out=H, ldo=1024, out_bs=0,
                                     dropout_p=0.3, dropout_seed=7))
        rec("H with fused dropout", ms, f_h, (T * 768 + T * 1024) * 2)
    if want("dx"):
        ms = timeit(lambda: ops.gemm(dH, W0, T, 768, 1024, b_mn=True, out_kind=L.OUT_BF16, out=dX, ldo=768, out_bs=0))
        rec("dX = dH W0                      K.MN N=768 K=1024", ms, 2.0 * T * 768 * 1024, (T * 1024 + T * 768) * 2)
    if want("dw0"):
        ms = timeit(lambda: ops.gemm(dH, X, 512, 768, T, a_mn=True, b_mn=True, out_kind=L.OUT_ATOMIC_F32, out=gW0, ldo=768,
                                     out_bs=0, k_splits=12))
        rec("dW0 = dH[:, :512]^T X           MN.MN split-K 12", ms, 2.0 * 512 * 768 * T, (T * 512 + T * 768) * 2)
    if want("dh"):
        ms = timeit(lambda: ops.gemm(dF, W1c, T, 512, 128, b_mn=True, out_kind=L.OUT_BF16, out=dH, ldo=1024, out_bs=0, mask=H,
                                     mask_scale=1.0 / 0.7, colsum=cs))
        rec("dH half = (dF W1c)*(H>0)        K.MN N=512 K=128", ms, 2.0 * T * 512 * 128, (T * 128 + 2 * T * 512) * 2)
    if want("epi"):           # K = 64: one k-block per tile, so the launch time is the epilogue (+ output write) alone
        X64 = X[:, :64].contiguous()
        W64 = W0[:, :64].contiguous()
        for nm, kw in (("epilogue only, relu+bias (pairs)", {}), ("epilogue only, relu+bias+dropout (single)", {"dropout_p": 0.3, "dropout_seed": 7})):
            ms = timeit(lambda: ops.gemm(X64, W64, T, 1024, 64, out_kind=L.OUT_BF16, bias=b0, relu=True, out=H, ldo=1024, out_bs=0, **kw))
            rec(nm + " K=64", ms, 2.0 * T * 1024 * 64, (T * 64 + T * 1024) * 2)
    if want("cublas"):        # library reference points for the same shapes (not on the product path)
        Wt = W0.t().contiguous()
        ms = timeit(lambda: torch.matmul(X, Wt, out=H))
        rec("cuBLAS  X @ W0^T (no epilogue)   N=1024 K=768", ms, f_h, (T * 768 + T * 1024) * 2)
        ms = timeit(lambda: torch.matmul(dH, W0, out=dX))
        rec("cuBLAS  dH @ W0                  N=768 K=1024", ms, 2.0 * T * 768 * 1024, (T * 1024 + T * 768) * 2)
        dHt = dH.t()
        g2 = torch.empty(1024, 768, device=dev, dtype=bf)
        ms = timeit(lambda: torch.matmul(dHt, X, out=g2))
        rec("cuBLAS  dH^T @ X (both halves)   M=1024 N=768 K=T", ms, 2.0 * 1024 * 768 * T, (T * 1024 + T * 768) * 2)
    if want("f"):
        F = torch.empty(T, 128, device=dev, dtype=bf)
        ms = timeit(lambda: ops.gemm(H, W1c, T, 128, 512, out_kind=L.OUT_BF16, out=F, ldo=128, out_bs=0))
        rec("F = H[:, :512] W1c^T            KK  N=128 K=512", ms, 2.0 * T * 128 * 512, (T * 512 + T * 128) * 2)


if __name__ == "__main__":
    main()
