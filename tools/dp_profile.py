"""Timeline of one data-parallel SALAD step (run under torchrun, N >= 2): does the gradient all-reduce overlap the
dX GEMM?   torchrun --nproc-per-node 2 tools/dp_profile.py [reserve_sms]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

import vpr_b200  # noqa: E402


def main():
    reserve = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    rank, local, world = vpr_b200.init_distributed("nccl")
    torch.cuda.set_device(local)
    B, n = 256, 256
    torch.manual_seed(0)
    head = vpr_b200.SaladHead(dropout=0.3).cuda().train()
    vpr_b200.overlap_grad_allreduce(head, reserve_sms=reserve)
    step = vpr_b200.GraphedSaladStep(head, B, n)
    for _ in range(5):
        step()
    torch.cuda.synchronize()
    dist.barrier()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(4):
            step()
        torch.cuda.synchronize()
    if rank == 0:
        evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
        evs.sort(key=lambda e: e.time_range.start)
        per = len(evs) // 4
        t0 = evs[-per].time_range.start
        for e in evs[-per:]:
            print(f"{e.time_range.start - t0:9.1f} .. {e.time_range.end - t0:9.1f}  (+{e.time_range.end - e.time_range.start:7.1f} us)  {e.name[:90]}")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
