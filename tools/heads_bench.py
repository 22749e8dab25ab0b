"""Sweep of the three GEMM shapes of the geopose-heads call (B=256, D=8448, hidden=512) over tile width and split-K,
through the vpr_gemm test hook: python tools/heads_bench.py [B]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import vpr_b200  # noqa: E402
from vpr_b200 import _lib as L, ops  # noqa: E402


def timed(fn, reps=20):
    """GPU time per call: `reps` calls captured in one CUDA graph (the host needs ~14 us to issue one GEMM through
    ctypes, longer than these kernels run), L2 left warm as in the real step."""
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            fn()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        for _ in range(reps):
            fn()
    graph.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        graph.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (3 * reps) * 1e3


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    D, hid, hp = 8448, 512, 520
    dev = "cuda"
    desc = torch.randn(B, D, device=dev).bfloat16()
    waug = torch.randn(hp, D, device=dev).bfloat16()
    zacc = torch.zeros(B, hp, device=dev)
    dz = torch.randn(B, hp, device=dev).bfloat16()
    ddesc = torch.empty(B, D, device=dev, dtype=torch.bfloat16)
    dw = torch.empty(hid, D, device=dev)
    print(f"B={B}")
    for N in (hid, hp):
        for bn in (64, 128, 256):
            for ks in (2, 4, 6, 8, 12, 16, 24, 33, 44, 66):
                def f():
                    ops.gemm(desc, waug, B, N, D, out_kind=L.OUT_ATOMIC_F32, out=zacc, ldo=hp, out_bs=0, k_splits=ks, bn=bn)
                try:
                    print(f"zacc N={N} bn={bn} ks={ks}: {timed(f):.1f} us")
                except RuntimeError as e:
                    print(f"zacc N={N} bn={bn} ks={ks}: {e}")
    for bn in (64, 128, 256):
        def f():
            ops.gemm(dz, waug, B, D, hp, b_mn=True, out_kind=L.OUT_BF16, out=ddesc, ldo=D, out_bs=0, bn=bn)
        print(f"d_desc bn={bn}: {timed(f):.1f} us")
    for bn in (128, 256):
        def f():
            ops.gemm(dz, desc, hid, D, B, a_mn=True, b_mn=True, out_kind=L.OUT_F32, out=dw, ldo=D, out_bs=0, bn=bn)
        print(f"dW0 bn={bn}: {timed(f):.1f} us")


if __name__ == "__main__":
    main()
