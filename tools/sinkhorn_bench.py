"""Stand-alone Sinkhorn forward/backward timing (CUDA events) and HBM fraction, register-tile kernels (n <= 256)
against the shared-memory-tile kernels (VPR_SINK_RT=0) in one process.
python tools/sinkhorn_bench.py [B] [n] [reps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import vpr_b200  # noqa: E402
from vpr_b200 import _lib as L  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
    dev = torch.device("cuda")
    S = (torch.randn(B, n, 64, device=dev) * 2).contiguous()
    db = torch.tensor([1.0], device=dev)
    dP = torch.randn(B, n, 64, device=dev)
    P = torch.empty(B, n, 64, dtype=torch.bfloat16, device=dev)
    dS = torch.empty(B, n, 64, dtype=torch.bfloat16, device=dev)
    U = torch.empty(B, 3, 65, device=dev)
    V = torch.empty(B, 3, n, device=dev)
    dbias = torch.zeros(64, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    lib, st = vpr_b200.lib(), L.stream_ptr(dev)
    bf = B * n * 64 * (4 + 2) + B * 3 * (65 + n) * 4
    bb = B * n * 64 * (4 + 4 + 2) + B * 3 * (65 + n) * 4

    def fwd():
        L.check(lib.vpr_sinkhorn_fwd(L.ptr(S), L.ptr(db), L.ptr(P), L.VPR_BF16, L.ptr(U), L.ptr(V), B, n, 64, 3, st), "fwd")

    def bwd():
        L.check(lib.vpr_sinkhorn_bwd(L.ptr(S), L.ptr(db), L.ptr(U), L.ptr(V), L.ptr(dP), L.ptr(dS), L.VPR_BF16, L.ptr(dbias), B, n, 64, 3, st),
                "bwd")

    def timed(fn, cold):
        ts = []
        for i in range(3 + reps):
            if cold:
                flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            if i >= 3:
                ts.append(e0.elapsed_time(e1))
        ts.sort()
        return ts[len(ts) // 2] * 1e3

    for rt in ("1", "0"):
        os.environ["VPR_SINK_RT"] = rt
        name = "register tile (sinkhorn_rt.cuh)" if rt == "1" and n <= 256 else "shared tile (sinkhorn_scaled.cuh)"
        fc, fw, bc, bw = timed(fwd, True), timed(fwd, False), timed(bwd, True), timed(bwd, False)
        print(f"B={B} n={n} {name}: fwd {fc:.1f} us cold ({bf/fc/1e3:.0f} GB/s) / {fw:.1f} us L2-warm; "
              f"bwd {bc:.1f} us cold ({bb/bc/1e3:.0f} GB/s) / {bw:.1f} us L2-warm   [median of {reps}, includes launch]")


if __name__ == "__main__":
    main()
