"""All-reduce microbenchmark (run under torchrun, N >= 2): the library's peer-memory kernel against ncclAllReduce on the
message sizes of the fine-tuning step, each timed as 20 calls captured in one CUDA graph.
    torchrun --nproc-per-node N tools/comm_bench.py            (env: VPR_COMM_MULTIMEM=0 for the P2P path, VPR_COMM_THREADS)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import vpr_b200  # noqa: E402
from vpr_b200.comm import PeerComm  # noqa: E402


def timed(fn, dev, reps=20):
    side = torch.cuda.Stream(dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
        for _ in range(3):
            fn()
    torch.cuda.current_stream(dev).wait_stream(side)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, capture_error_mode="thread_local"):
        for _ in range(reps):
            fn()
    dist.barrier()
    graph.replay()
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        graph.replay()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / (3 * reps) * 1e3], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item(), graph


def main():
    rank, local, world = vpr_b200.init_distributed("nccl")
    dev = torch.device("cuda", local)
    sizes = (5_754_832, 1_411_012, 4_343_816, 262_144, 16_384)
    comm = PeerComm(max(sizes), device=dev)
    if rank == 0:
        print(f"world={world} backend={comm.backend} multimem={bool(comm.mc_ptr)} threads={os.environ.get('VPR_COMM_THREADS', '1024')}")
    keep = []
    for n in sizes:
        comm.data.zero_()
        for ctas in (8, 16, 32, 64, 148):
            us, g = timed(lambda: comm.allreduce_(0, n, ctas=ctas), dev)
            keep.append(g)
            if rank == 0:
                print(f"n={n:9d} ({n * 4 / 1e6:6.2f} MB) peer ctas={ctas:3d}: {us:7.1f} us   busbw {2 * (world - 1) / world * n * 4 / us / 1e3:6.1f} GB/s")
        buf = torch.zeros(n, device=dev)
        us, g = timed(lambda: dist.all_reduce(buf), dev)
        keep.append(g)
        if rank == 0:
            print(f"n={n:9d} ({n * 4 / 1e6:6.2f} MB) nccl          : {us:7.1f} us   busbw {2 * (world - 1) / world * n * 4 / us / 1e3:6.1f} GB/s")
    torch.cuda.synchronize()
    dist.barrier()
    sys.stdout.flush()
    os._exit(0)


if __name__ == "__main__":
    main()
