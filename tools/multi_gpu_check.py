"""N-rank check of the two sharded paths on real GPUs, launched with torchrun:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 tools/multi_gpu_check.py

1. Retrieval: every rank holds all queries and a contiguous gallery shard; knn_topk_sharded (per-rank fused top-k,
   NCCL allgather, merge kernel) must give exactly the single-GPU answer over the whole gallery.
2. Fine-tuning: DataParallelStep on a batch shard must leave every rank with the gradients of the global-batch loss
   (compared with one process running the whole batch), with and without the in-backward overlapped all-reduce.
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
        lsum = loss.clone()
        dist.all_reduce(lsum)
        err = rel(got, want)
        e = torch.tensor([err], device=dev)
        dist.all_reduce(e, op=dist.ReduceOp.MAX)
        good = e.item() < 2e-2 and abs(lsum.item() - want_loss) < 1e-3 * abs(want_loss)
        if rank == 0:
            print(json.dumps({"check": "DataParallelStep gradients == global-batch gradients", "world": world,
                              "overlapped_allreduce": overlap, "grad_elements": step.grad_elements(),
                              "max_rel_err_over_ranks": e.item(), "loss_sum": lsum.item(), "loss_single": want_loss,
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

    # ---------------- configs[2] as one graph: SALAD + heads + losses with both all-reduces captured ----------------
    step = vpr_b200.DataParallelStep(salad, heads)
    step.step(x[b0:b1], t[b0:b1], ll[b0:b1], ang[b0:b1], global_B=Bg)          # eager, un-hooked: the reference gradients
    torch.cuda.synchronize()
    want_f = grads().clone()
    fstep = vpr_b200.GraphedFinetuneStep(salad, heads, batch=b1 - b0, n_tokens=n, global_batch=Bg)
    for rep in range(2):
        tot, _ = fstep(x[b0:b1], t[b0:b1], ll[b0:b1], ang[b0:b1])
    torch.cuda.synchronize()
    e = torch.tensor([rel(grads(), want_f)], device=dev)
    dist.all_reduce(e, op=dist.ReduceOp.MAX)
    lsum = tot.clone()
    dist.all_reduce(lsum)
    good = e.item() < 1e-5 and abs(lsum.item() - want_loss) < 1e-3 * abs(want_loss)
    if rank == 0:
        print(json.dumps({"check": "GraphedFinetuneStep (in-graph all-reduces) == eager DataParallelStep", "world": world,
                          "launches_per_replay": fstep.launches_per_replay, "grad_rel_err": e.item(),
                          "loss_sum": lsum.item(), "loss_single": want_loss, "ok": good}), flush=True)
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
