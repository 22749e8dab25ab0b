"""N-rank check of the two sharded paths on real GPUs, launched with torchrun -- by
tests/test_gpu_multi.py when the box has >= 2 GPUs, or by hand:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 tests/multi_gpu_worker.py

1. Retrieval: every rank holds all queries and a contiguous gallery shard; knn_topk_sharded (per-rank fused top-k,
   NCCL allgather, merge kernel) must give exactly the single-GPU answer over the whole gallery.
2. Fine-tuning: DataParallelStep on a batch shard must leave every rank with the gradients AND the losses of the
   global batch (compared with one process running the whole batch), with and without the in-backward overlapped
   all-reduce; GraphedFinetuneStep likewise, with NCCL and with the library's own peer-memory all-reduce kernel.
3. The all-reduce kernel itself (vpr_allreduce_sum_f32 over comm.PeerComm) against the sum computed by NCCL, over
   message sizes, CTA counts, repeated calls (the flags carry a call counter) and CUDA-graph replays.
Prints one JSON line per check from rank 0; exit code 1 on any mismatch."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import vpr_b200  # noqa: E402


def rel(a, b):
    return ((a.float() - b.float()).norm() / (b.float().norm() + 1e-30)).item()


def main():
    rank, local, world = vpr_b200.init_distributed("nccl")
    dev = torch.device("cuda", local)
    ok = True

    # ---------------- retrieval ----------------
    Q, G, D, k = 512, 40000, 8448, 25
    g = torch.Generator(device=dev).manual_seed(777)                 # same data on every rank
    gal = torch.nn.functional.normalize(torch.randn(G, D, device=dev, generator=g), dim=1).bfloat16()
    qry = torch.nn.functional.normalize(torch.randn(Q, D, device=dev, generator=g), dim=1).bfloat16()
    lo, hi = vpr_b200.shard_range(G, rank, world)
    s_sh, i_sh = vpr_b200.knn_topk_sharded(qry, gal[lo:hi].contiguous(), k, lo)
    s_one, i_one = vpr_b200.knn_topk(qry, gal, k)
    same_idx = bool(torch.equal(i_sh, i_one))
    same_score = bool(torch.equal(s_sh, s_one))
    flag = torch.tensor([int(same_idx and same_score)], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(json.dumps({"check": "knn_topk_sharded == knn_topk over the whole gallery", "world": world, "Q": Q, "G": G,
                          "k": k, "idx_equal": same_idx, "score_equal": same_score, "all_ranks": bool(flag.item())}), flush=True)
    ok = ok and bool(flag.item())

    # ---------------- the all-reduce kernel over peer memory ----------------
    from vpr_b200.comm import PeerComm
    comm = PeerComm(4_343_816, device=dev)
    worst, bit_equal_across_ranks = 0.0, True
    for count, ctas in ((4, 1), (1024, 3), (1_411_012, 64), (4_343_816, 16), (4_343_816, 64), (999_996, 7)):
        for rep in range(3):
            gq = torch.Generator(device=dev).manual_seed(1000 * rep + rank)
            src = torch.randn(count, device=dev, generator=gq)
            want_sum = src.clone()
            dist.all_reduce(want_sum)
            comm.data[:count].copy_(src)
            comm.data[count:count + 4].fill_(7.0) if count + 4 <= comm.numel else None      # must stay untouched
            comm.allreduce_(0, count, ctas=ctas)
            torch.cuda.synchronize()
            worst = max(worst, rel(comm.data[:count], want_sum))
            if count + 4 <= comm.numel:
                worst = max(worst, float((comm.data[count:count + 4] != 7.0).any()))
            mine = comm.data[:count].clone()
            ref0 = mine.clone()
            dist.broadcast(ref0, 0)
            bit_equal_across_ranks = bit_equal_across_ranks and bool(torch.equal(mine, ref0))
    # an offset message, and 50 replays of a captured call (the launch arguments are frozen: the call counter must live
    # on the device)
    comm.data.zero_()
    comm.data[4096:4096 + 8192].fill_(float(rank + 1))
    comm.allreduce_(4096, 8192, ctas=5)
    torch.cuda.synchronize()
    off_ok = bool((comm.data[4096:4096 + 8192] == world * (world + 1) / 2).all()) and float(comm.data[:4096].abs().sum()) == 0.0
    side = torch.cuda.Stream(dev)
    with torch.cuda.stream(side):
        comm.data[:65536].fill_(1.0)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, capture_error_mode="thread_local"):
            comm.allreduce_(0, 65536, ctas=8)
        for _ in range(50):
            comm.data[:65536].fill_(1.0)
            graph.replay()
    torch.cuda.synchronize()
    graph_ok = bool((comm.data[:65536] == float(world)).all())
    e = torch.tensor([worst, 0.0 if (bit_equal_across_ranks and off_ok and graph_ok) else 1.0], device=dev)
    dist.all_reduce(e, op=dist.ReduceOp.MAX)
    good = e[0].item() < 1e-6 and e[1].item() == 0.0
    if rank == 0:
        print(json.dumps({"check": "vpr_allreduce_sum_f32 == ncclAllReduce(sum)", "world": world, "backend": comm.backend,
                          "multimem": bool(comm.mc_ptr), "setup_notes": comm.setup_errors, "max_rel_err": e[0].item(),
                          "bit_identical_on_all_ranks": bit_equal_across_ranks, "offset_message": off_ok,
                          "graph_replays": graph_ok, "ok": good}), flush=True)
    ok = ok and good
    del graph

    # ---------------- fine-tuning ----------------
    Bg, n = 8 * world, 256
    torch.manual_seed(0)
    salad = vpr_b200.SaladHead(dropout=0.0).to(dev).train()
    heads = vpr_b200.GeoposeHeads().to(dev).train()
    gg = torch.Generator(device=dev).manual_seed(5)
    x = torch.randn(Bg, n, 768, device=dev, generator=gg).bfloat16()
    t = torch.randn(Bg, 768, device=dev, generator=gg).bfloat16()
    ll = torch.randn(Bg, 2, device=dev, generator=gg)
    ang = torch.rand(Bg, device=dev, generator=gg) * 360

    def grads():
        return torch.cat([p.grad.flatten().float() for p in list(salad.parameters()) + list(heads.parameters())])

    # reference: the whole batch in this process, no collective
    for p in list(salad.parameters()) + list(heads.parameters()):
        p.grad = None
    desc = salad((x, t))
    d = desc.detach().requires_grad_(True)
    total, *_ = heads(d, ll, ang, global_B=Bg)
    total.backward()
    desc.backward(d.grad)
    want = grads().clone()
    want_loss = total.item()

    b0, b1 = vpr_b200.shard_range(Bg, rank, world)
    for overlap in (False, True):
        vpr_b200.overlap_grad_allreduce(salad, reserve_sms=32) if overlap else salad.set_grad_ready_hook(None)
        step = vpr_b200.DataParallelStep(salad, heads)
        loss, _ = step.step(x[b0:b1], t[b0:b1], ll[b0:b1], ang[b0:b1], global_B=Bg)
        torch.cuda.synchronize()
        got = grads()
        lsum = loss.clone()                               # already the global-batch loss on every rank
        err = rel(got, want)
        e = torch.tensor([err], device=dev)
        dist.all_reduce(e, op=dist.ReduceOp.MAX)
        good = e.item() < 2e-2 and abs(lsum.item() - want_loss) < 1e-3 * abs(want_loss)
        if rank == 0:
            print(json.dumps({"check": "DataParallelStep gradients == global-batch gradients", "world": world,
                              "overlapped_allreduce": overlap, "grad_elements": step.grad_elements(),
                              "max_rel_err_over_ranks": e.item(), "loss_returned": lsum.item(), "loss_single": want_loss,
                              "ok": good}), flush=True)
        ok = ok and good
    salad.set_grad_ready_hook(None)

    # ---------------- the benchmarked N>1 step: fwd+bwd + overlapped all-reduce replayed from one CUDA graph ----------------
    Bl = 64
    xs = torch.randn(Bl, n, 768, device=dev, generator=torch.Generator(device=dev).manual_seed(100 + rank)).bfloat16()
    ts = torch.randn(Bl, 768, device=dev, generator=torch.Generator(device=dev).manual_seed(200 + rank)).bfloat16()
    ds = torch.randn(Bl, 8448, device=dev, generator=torch.Generator(device=dev).manual_seed(300 + rank)).bfloat16()
    for p in salad.parameters():
        p.grad = None
    xe, te = xs.clone().requires_grad_(True), ts.clone().requires_grad_(True)
    salad((xe, te)).backward(ds)
    local = torch.cat([p.grad.flatten().float() for p in salad.parameters()])
    want_g = local.clone()
    dist.all_reduce(want_g)                                   # sum over ranks of the local gradients
    want_dx = xe.grad.clone()
    vpr_b200.overlap_grad_allreduce(salad, reserve_sms=32)
    gstep = vpr_b200.GraphedSaladStep(salad, batch=Bl, n_tokens=n, capture_error_mode=os.environ.get("CAPTURE_MODE", "thread_local"))
    for rep in range(2):                                      # the second replay must not accumulate
        out, dx, dt = gstep(xs, ts, ds)
    torch.cuda.synchronize()
    got_g = torch.cat([p.grad.flatten().float() for p in salad.parameters()])
    e = torch.tensor([rel(got_g, want_g), rel(dx, want_dx)], device=dev)
    dist.all_reduce(e, op=dist.ReduceOp.MAX)
    good = e[0].item() < 1e-5 and e[1].item() < 1e-5
    if rank == 0:
        print(json.dumps({"check": "GraphedSaladStep + in-graph NCCL all-reduce == all-reduce of eager local gradients",
                          "world": world, "launches_per_replay": gstep.launches_per_replay,
                          "param_grad_rel_err": e[0].item(), "dx_rel_err": e[1].item(), "ok": good}), flush=True)
    ok = ok and good
    salad.set_grad_ready_hook(None)
    # the same step with the library's own all-reduce kernel (gradient buffer in the symmetric allocation)
    pstep = vpr_b200.GraphedSaladStep(salad, batch=Bl, n_tokens=n, comm="peer")
    for rep in range(3):
        out, dx, dt = pstep(xs, ts, ds)
    torch.cuda.synchronize()
    got_g = torch.cat([p.grad.flatten().float() for p in salad.parameters()])
    e = torch.tensor([rel(got_g, want_g), rel(dx, want_dx)], device=dev)
    dist.all_reduce(e, op=dist.ReduceOp.MAX)
    good = e[0].item() < 1e-5 and e[1].item() < 1e-5
    if rank == 0:
        print(json.dumps({"check": "GraphedSaladStep + own peer-memory all-reduce == all-reduce of eager local gradients",
                          "world": world, "comm": pstep.comm_kind, "param_grad_rel_err": e[0].item(),
                          "dx_rel_err": e[1].item(), "ok": good}), flush=True)
    ok = ok and good
    salad.set_grad_ready_hook(None)
    del pstep

    # ---------------- configs[2] as one graph: SALAD + heads + losses with both all-reduces captured ----------------
    step = vpr_b200.DataParallelStep(salad, heads)
    step.step(x[b0:b1], t[b0:b1], ll[b0:b1], ang[b0:b1], global_B=Bg)          # eager, un-hooked: the reference gradients
    torch.cuda.synchronize()
    want_f = grads().clone()
    for kind in ("nccl", "peer"):
        salad.set_grad_ready_hook(None)
        fstep = vpr_b200.GraphedFinetuneStep(salad, heads, batch=b1 - b0, n_tokens=n, global_batch=Bg, comm=kind)
        for rep in range(3):
            tot, _ = fstep(x[b0:b1], t[b0:b1], ll[b0:b1], ang[b0:b1])
        torch.cuda.synchronize()
        e = torch.tensor([rel(grads(), want_f)], device=dev)
        dist.all_reduce(e, op=dist.ReduceOp.MAX)
        good = e.item() < 1e-5 and abs(tot.item() - want_loss) < 1e-3 * abs(want_loss)
        if rank == 0:
            print(json.dumps({"check": "GraphedFinetuneStep (in-graph all-reduces) == eager DataParallelStep", "world": world,
                              "comm": fstep.comm_kind, "comm_note": fstep.comm_note,
                              "launches_per_replay": fstep.launches_per_replay, "grad_rel_err": e.item(),
                              "loss_returned": tot.item(), "loss_single": want_loss, "ok": good}), flush=True)
        ok = ok and good
        salad.set_grad_ready_hook(None)
        del fstep
    del gstep                                                 # the graph holds captured NCCL work: drop it before teardown
    torch.cuda.synchronize()
    dist.barrier()
    sys.stdout.flush()
    os._exit(0 if ok else 1)                                  # skip communicator teardown (it can wait on graph resources)


if __name__ == "__main__":
    main()
