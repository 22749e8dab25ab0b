"""Stand-alone Sinkhorn forward/backward timing (CUDA events) and HBM fraction.  python tools/sinkhorn_bench.py [B] [n]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import vpr_b200  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    iters = int(sys.argv[3]) if len(sys.argv) > 3 else 5
    S = (torch.randn(B, n, 64, device="cuda") * 2).contiguous()
    db = torch.tensor(1.0, device="cuda")
    dP = torch.randn(B, n, 64, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    tf, tb = [], []
    for i in range(2 + iters):
        flush.zero_()
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record()
        P, U, V = vpr_b200.sinkhorn(S, db, 3, out_dtype=torch.bfloat16)
        e1.record()
        dS, _ = vpr_b200.sinkhorn_backward(S, db, U, V, dP, 3, out_dtype=torch.bfloat16)
        e2.record()
        torch.cuda.synchronize()
        if i >= 2:
            tf.append(e0.elapsed_time(e1))
            tb.append(e1.elapsed_time(e2))
    f, b = sum(tf) / len(tf), sum(tb) / len(tb)
    bf = B * n * 64 * (4 + 2) + B * 3 * (65 + n) * 4
    bb = B * n * 64 * (4 + 4 + 2) + B * 3 * (65 + n) * 4
    print(f"sinkhorn B={B} n={n}: fwd {f*1e3:.1f} us ({bf/f/1e6:.0f} GB/s)  bwd {b*1e3:.1f} us ({bb/b/1e6:.0f} GB/s)  [includes torch.empty/launch]")


if __name__ == "__main__":
    main()
