"""Descriptor kNN throughput (BASELINE configs[4], one GPU's gallery shard): queries/s and the tensor-pipe
fraction of the fused GEMM + top-k kernel.   python tools/knn_bench.py [Q] [G] [D] [k]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import vpr_b200  # noqa: E402


def make_rows(rows, D, seed, device="cuda", chunk=65536):
    out = torch.empty(rows, D, dtype=torch.bfloat16, device=device)
    g = torch.Generator(device=device).manual_seed(seed)
    for r0 in range(0, rows, chunk):
        r1 = min(rows, r0 + chunk)
        x = torch.randn(r1 - r0, D, device=device, generator=g)
        out[r0:r1] = torch.nn.functional.normalize(x, dim=1).bfloat16()
    return out


def main():
    Q = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
    G = int(sys.argv[2]) if len(sys.argv) > 2 else 1_250_000
    D = int(sys.argv[3]) if len(sys.argv) > 3 else 8448
    k = int(sys.argv[4]) if len(sys.argv) > 4 else 25
    q = make_rows(Q, D, 555)
    g = make_rows(G, D, 777)
    for _ in range(2):
        vpr_b200.knn_topk(q, g, k)
    torch.cuda.synchronize()
    ts = []
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        s, i = vpr_b200.knn_topk(q, g, k)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = sum(ts) / len(ts)
    fl = 2.0 * Q * G * D
    print(f"kNN Q={Q} G={G} D={D} k={k}: {ms:.2f} ms  {Q / ms * 1e3:.0f} q/s  {fl / ms / 1e9:.1f} TFLOP/s")
    # spot check against a direct product on a slice of the queries
    ref = (q[:64].float() @ g.float().T) if G <= 200_000 else None
    if ref is not None:
        rs, ri = ref.topk(k, dim=1)
        match = sum(set(a.tolist()) == set(b.tolist()) for a, b in zip(i[:64].cpu(), ri.cpu())) / 64
        print(f"top-{k} set match on 64 queries: {match:.3f}")


if __name__ == "__main__":
    main()
