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

    # back to back over rotating working sets (no host launch gap in the figure, operands never L2-resident)
    nsets = 6
    sets = [((torch.randn(B, n, 64, device=dev) * 2).contiguous(), torch.randn(B, n, 64, device=dev),
             torch.empty(B, n, 64, dtype=torch.bfloat16, device=dev), torch.empty(B, n, 64, dtype=torch.bfloat16, device=dev),
             torch.empty(B, 3, 65, device=dev), torch.empty(B, 3, n, device=dev)) for _ in range(nsets)]

    def fwd_i(i):
        S_, dP_, P_, dS_, U_, V_ = sets[i % nsets]
        L.check(lib.vpr_sinkhorn_fwd(L.ptr(S_), L.ptr(db), L.ptr(P_), L.VPR_BF16, L.ptr(U_), L.ptr(V_), B, n, 64, 3, st), "fwd")

    def bwd_i(i):
        S_, dP_, P_, dS_, U_, V_ = sets[i % nsets]
        L.check(lib.vpr_sinkhorn_bwd(L.ptr(S_), L.ptr(db), L.ptr(U_), L.ptr(V_), L.ptr(dP_), L.ptr(dS_), L.VPR_BF16, L.ptr(dbias), B, n, 64,
                                     3, st), "bwd")

    def pipelined(fn):
        for i in range(nsets):
            fn(i)
        torch.cuda.synchronize()
        best = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(24):
                fn(i)
            e1.record()
            torch.cuda.synchronize()
            best.append(e0.elapsed_time(e1) / 24 * 1e3)
        best.sort()
        return best[len(best) // 2]

    for rt in ("1", "2", "0"):
        os.environ["VPR_SINK_RT"] = rt
        name = ("shared tile (sinkhorn_scaled.cuh)" if rt == "0" or n > 256 else
                "register tile, 512 threads" if rt == "1" else "register tile, 256 threads")
        fc, fw, bc, bw = timed(fwd, True), timed(fwd, False), timed(bwd, True), timed(bwd, False)
        fp, bp = pipelined(fwd_i), pipelined(bwd_i)
        print(f"B={B} n={n} {name}: back to back fwd {fp:.1f} us ({bf/fp/1e3:.0f} GB/s) bwd {bp:.1f} us ({bb/bp/1e3:.0f} GB/s) | "
              f"single launch incl. launch latency: fwd {fc:.1f} cold / {fw:.1f} warm, bwd {bc:.1f} cold / {bw:.1f} warm us", flush=True)


def sweep():
    """Forward time against the iteration count: slope = one iteration, intercept = launch + load + exp + store."""
    B, n = 256, 256
    dev = torch.device("cuda")
    lib, st = vpr_b200.lib(), L.stream_ptr(dev)
    db = torch.tensor([1.0], device=dev)
    nsets = 6
    sets = [((torch.randn(B, n, 64, device=dev) * 2).contiguous(), torch.empty(B, n, 64, dtype=torch.bfloat16, device=dev),
             torch.empty(B, 8, 65, device=dev), torch.empty(B, 8, n, device=dev)) for _ in range(nsets)]
    for rt in ("1", "2", "0"):
        os.environ["VPR_SINK_RT"] = rt
        out = []
        for iters in (1, 2, 3, 5, 8):
            def fn(i):
                S_, P_, U_, V_ = sets[i % nsets]
                L.check(lib.vpr_sinkhorn_fwd(L.ptr(S_), L.ptr(db), L.ptr(P_), L.VPR_BF16, L.ptr(U_), L.ptr(V_), B, n, 64, iters, st), "fwd")
            for i in range(nsets):
                fn(i)
            torch.cuda.synchronize()
            ts = []
            for _ in range(5):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for i in range(24):
                    fn(i)
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1) / 24 * 1e3)
            ts.sort()
            out.append(f"{iters} it: {ts[2]:.1f} us")
        print(f"VPR_SINK_RT={rt} forward, back to back: " + ", ".join(out), flush=True)


if __name__ == "__main__":
    if "--sweep" in sys.argv:
        sweep()
        sys.exit(0)
    main()
