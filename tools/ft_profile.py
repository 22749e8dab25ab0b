"""Timeline of one fine-tuning step (BASELINE configs[2]: SALAD head + geopose heads + losses, one CUDA graph):
CUPTI kernel records of the last of a few replays.  1 GPU: python tools/ft_profile.py [B] [n];
N GPUs: torchrun --nproc-per-node N tools/ft_profile.py  (shows where the gradient collectives sit).
Run with VPR_PDL=0 for per-kernel durations (see tools/step_profile.py)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

import vpr_b200  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = local = 0
    if world > 1:
        rank, local, world = vpr_b200.init_distributed("nccl")
    torch.cuda.set_device(local)
    torch.manual_seed(0)
    head = vpr_b200.SaladHead(dropout=0.3).cuda().train()
    heads = vpr_b200.GeoposeHeads(latlon_loss="mse", angle_loss="atan2_maae").cuda().train()
    step = vpr_b200.GraphedFinetuneStep(head, heads, B, n, global_batch=B * world)
    g = torch.Generator(device="cuda").manual_seed(1 + rank)
    step(torch.randn(B, n, 768, device="cuda", generator=g).bfloat16(), torch.randn(B, 768, device="cuda", generator=g).bfloat16(),
         torch.randn(B, 2, device="cuda", generator=g), torch.rand(B, device="cuda", generator=g) * 360)
    for _ in range(5):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        step()
    e1.record()
    torch.cuda.synchronize()
    reps = 4
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(reps):
            step()
        torch.cuda.synchronize()
    if rank == 0:
        print(f"B={B} n={n} world={world}: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us/step, {step.launches_per_replay} launches/step")
        evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
        evs.sort(key=lambda e: e.time_range.start)
        per = len(evs) // reps
        t0 = evs[-per].time_range.start
        for e in evs[-per:]:
            print(f"{e.time_range.start - t0:9.1f} .. {e.time_range.end - t0:9.1f}  (+{e.time_range.end - e.time_range.start:7.1f} us)  {e.name[:100]}")
    if world > 1:
        if rank == 0:
            print("comm:", step.comm_kind, step.comm_placement, step.comm_note)
        torch.cuda.synchronize()
        dist.barrier()
        sys.stdout.flush()
        os._exit(0)             # skip communicator teardown: it can wait on the captured NCCL work the graph holds


if __name__ == "__main__":
    main()
