"""Per-kernel device times of one SALAD fwd+bwd step (CUPTI via torch.profiler; warm, pipelined --
complements the cold, serialised ncu launch list).  python tools/step_profile.py [B] [n]

Run it with VPR_PDL=0 for per-kernel durations: with programmatic dependent launches (the default) a kernel is
resident, and its CUPTI duration running, while it still waits for the previous launch to finish."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

import vpr_b200  # noqa: E402


def main():
    as_json = "--json" in sys.argv                 # one JSON line {"kernels": [{"name", "us", "launches_per_step"}, ...]} for bench.py
    argv = [a for a in sys.argv[1:] if a != "--json"]
    B = int(argv[0]) if len(argv) > 0 else 256
    n = int(argv[1]) if len(argv) > 1 else 256
    hw = int(round(n ** 0.5))
    torch.manual_seed(0)
    head = vpr_b200.SaladHead(dropout=0.3).cuda().train()
    x = torch.randn(B, n, 768, device="cuda").bfloat16()
    t = torch.randn(B, 768, device="cuda").bfloat16()
    w = torch.randn(B, 8448, device="cuda").bfloat16()

    def step():
        for p in head.parameters():
            p.grad = None
        xg = x.detach().requires_grad_(True)
        tg = t.detach().requires_grad_(True)
        out = head((xg.view(B, hw, hw, 768).permute(0, 3, 1, 2), tg))
        out.backward(w)

    for _ in range(5):
        step()
    torch.cuda.synchronize()
    import time
    t0 = time.perf_counter()
    for _ in range(50):
        step()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    if not as_json:
        print(f"host issue time {1e6 * (t1 - t0) / 50:.0f} us/step, wall incl. drain {1e6 * (t2 - t0) / 50:.0f} us/step")
    steps = 10
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(steps):
            step()
        torch.cuda.synchronize()
    rows = []
    for e in prof.key_averages():
        dt = getattr(e, "device_time_total", None)
        if dt is None:
            dt = e.cuda_time_total
        if dt > 0:
            rows.append((dt / steps, e.count / steps, e.key))
    rows.sort(reverse=True)
    tot = sum(r[0] for r in rows)
    if as_json:
        import json
        print(json.dumps({"B": B, "n": n, "pdl": os.environ.get("VPR_PDL", "1") != "0", "sum_us": tot,
                          "kernels": [{"name": key, "us": dt, "launches_per_step": cnt} for dt, cnt, key in rows]}))
        return
    print(f"B={B} n={n}: sum of kernel time per step {tot:.1f} us")
    for dt, cnt, key in rows:
        print(f"{dt:9.1f} us  x{cnt:4.1f}  {key[:150]}")
    # in-order trace of the last step
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    evs.sort(key=lambda e: e.time_range.start)
    per = len(evs) // steps
    print("--- last step, launch order ---")
    t0 = evs[-per].time_range.start
    for e in evs[-per:]:
        print(f"{e.time_range.start - t0:9.1f} +{e.time_range.end - e.time_range.start:8.1f} us  {e.name[:120]}")


if __name__ == "__main__":
    main()
