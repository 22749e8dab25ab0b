"""SaladHead: drop-in for the SALAD aggregator the reference pulls in through
``torch.hub.load("serizba/salad", "dinov2_salad")`` (reference
dinov2salad/dinov2salad_finetuning.py:19,92; dinov2salad/dinov2salad_validation.py:65).

Same constructor arguments, parameter names/shapes and ``forward((patch_map, cls_token)) ->
[B, token_dim + cluster_dim*num_clusters]`` L2-normalised descriptor contract as upstream
``SALAD``, with an autograd backward.  All arithmetic runs in
libvpr_b200.so (hand-written sm_100a kernels); this file only marshals pointers.
"""
from __future__ import annotations

import ctypes as C
from typing import Tuple

import torch
import torch.nn as nn

from . import _lib as L

_PARAM_ORDER = (
    ("cluster_w0", "cluster_features.0.weight"), ("cluster_b0", "cluster_features.0.bias"),
    ("cluster_w1", "cluster_features.3.weight"), ("cluster_b1", "cluster_features.3.bias"),
    ("score_w0", "score.0.weight"), ("score_b0", "score.0.bias"),
    ("score_w1", "score.3.weight"), ("score_b1", "score.3.bias"),
    ("token_w0", "token_features.0.weight"), ("token_b0", "token_features.0.bias"),
    ("token_w1", "token_features.2.weight"), ("token_b1", "token_features.2.bias"),
    ("dust_bin", "dust_bin"),
)


def _params_struct(cls, tensors):
    s = cls()
    for (field, _), t in zip(_PARAM_ORDER, tensors):
        setattr(s, field, t.data_ptr())
    return s


def _salad_forward(x_tok, t, shape_tuple, dropout_p, seed, seed_dev, params):
    """One vpr_salad_fwd call (+ the device step-counter advance).  Returns (out, save)."""
    lib = L.lib()
    B, n, d, hidden, m, l, g, iters = shape_tuple
    dt = L.dtype_code(x_tok.dtype)
    shape = L.SaladShape(B, n, d, hidden, m, l, g, iters)
    dev = x_tok.device
    save = torch.empty(lib.vpr_salad_save_bytes(C.byref(shape), dt), dtype=torch.uint8, device=dev)
    out = torch.empty(B, g + l * m, dtype=x_tok.dtype, device=dev)
    ps = _params_struct(L.SaladParams, params)
    with L.on_device(x_tok):
        L.check(lib.vpr_salad_fwd(C.byref(ps), L.ptr(x_tok), L.ptr(t), L.ptr(out), L.ptr(save), None, 0,
                                  C.byref(shape), dt, float(dropout_p), int(seed), L.ptr(seed_dev), L.stream_ptr(dev)),
                "vpr_salad_fwd")
    if seed_dev is not None and dropout_p > 0:      # next call (or next graph replay) draws a fresh mask
        with L.on_device(x_tok):
            L.check(lib.vpr_counter_advance(L.ptr(seed_dev), L.stream_ptr(dev)), "vpr_counter_advance")
    return out, save


def _salad_backward(x_tok, t, save, params, shape_tuple, dropout_p, d_out, want_dx, want_dt, hook, flat=None):
    """vpr_salad_bwd, in one call or -- with a gradient-ready hook -- in its two phases around the hook.
    Returns (dx, dt, flat fp32 gradient buffer, per-parameter views of it)."""
    lib = L.lib()
    B, n, d, hidden, m, l, g, iters = shape_tuple
    dt = L.dtype_code(x_tok.dtype)
    shape = L.SaladShape(B, n, d, hidden, m, l, g, iters)
    dev = x_tok.device
    d_out = d_out.contiguous()
    if d_out.dtype != x_tok.dtype:
        d_out = d_out.to(x_tok.dtype)
    ws_bytes = lib.vpr_salad_workspace_bytes(C.byref(shape), dt)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    sizes = [p.numel() for p in params]
    if flat is None:
        flat = torch.empty(sum(sizes), dtype=torch.float32, device=dev)
    elif flat.numel() < sum(sizes) or flat.dtype != torch.float32 or flat.device != dev or not flat.is_contiguous():
        raise RuntimeError("flat gradient buffer must be a contiguous fp32 tensor on the input's device")
    grads = list(torch.split(flat[: sum(sizes)], sizes))
    gs = _params_struct(L.SaladGrads, grads)
    ps = _params_struct(L.SaladParams, params)
    dx = torch.empty_like(x_tok) if want_dx else None
    dtok = torch.empty_like(t) if want_dt else None

    def bwd(phase, reserve):
        with L.on_device(x_tok):
            L.check(lib.vpr_salad_bwd(C.byref(ps), L.ptr(x_tok), L.ptr(t), L.ptr(d_out), L.ptr(save), L.ptr(ws),
                                      ws_bytes, C.byref(gs), L.ptr(dx), L.ptr(dtok), C.byref(shape), dt,
                                      float(dropout_p), phase, reserve, L.stream_ptr(dev)),
                    "vpr_salad_bwd")

    if hook is None:
        bwd(L.BWD_ALL, 0)
    else:
        # data parallel: every parameter gradient first, hand the flat buffer to the hook (it starts the
        # all-reduce on the collective's stream), then the input-gradient GEMM on a few SMs less so that the
        # collective's kernel finds room and the two overlap; finally order later work after the collective
        fn, reserve = hook[0], hook[1]
        bwd(L.BWD_PARAMS, int(hook[2]) if len(hook) > 2 else 0)
        work = fn(flat)
        if dx is not None or dtok is not None:
            bwd(L.BWD_INPUTS, int(reserve))
        if work is not None:
            work.wait()
    return dx, dtok, flat, grads


class _SaladFn(torch.autograd.Function):
    """x_tok: [B, n, d] contiguous, t: [B, d]; params in _PARAM_ORDER (fp32, contiguous)."""

    @staticmethod
    def forward(ctx, x_tok, t, shape_tuple, dropout_p, seed, seed_dev, grad_hook, *params):
        out, save = _salad_forward(x_tok, t, shape_tuple, dropout_p, seed, seed_dev, params)
        ctx.shape_tuple = shape_tuple
        ctx.dropout_p = float(dropout_p)
        ctx.grad_hook = grad_hook
        ctx.save_for_backward(x_tok, t, save, *params)
        return out

    @staticmethod
    def backward(ctx, d_out):
        x_tok, t, save, *params = ctx.saved_tensors
        dx, dtok, _flat, grads = _salad_backward(x_tok, t, save, params, ctx.shape_tuple, ctx.dropout_p, d_out,
                                                 ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.grad_hook)
        pg = [gr.view_as(p) if ctx.needs_input_grad[7 + i] else None
              for i, (gr, p) in enumerate(zip(grads, params))]
        return (dx, dtok, None, None, None, None, None, *pg)


class SaladHead(nn.Module):
    """B200-native SALAD aggregation head.

    ``forward((f, t))``: ``f`` is the patch-token map, either ``[B, C, h, w]`` (any strides; the
    permuted channels-last view upstream's DINOv2 wrapper produces is consumed without a copy)
    or ``[B, n, C]``; ``t`` is the CLS token ``[B, C]``.  fp32 inputs run the fp32 parity path,
    bf16 inputs the tcgen05 path; the descriptor has the input dtype.
    """

    def __init__(self, num_channels: int = 768, num_clusters: int = 64, cluster_dim: int = 128,
                 token_dim: int = 256, dropout: float = 0.3, sinkhorn_iters: int = 3):
        super().__init__()
        self.num_channels = num_channels
        self.num_clusters = num_clusters
        self.cluster_dim = cluster_dim
        self.token_dim = token_dim
        self.dropout_p = float(dropout)
        self.sinkhorn_iters = sinkhorn_iters
        drop = nn.Dropout(dropout) if dropout > 0 else nn.Identity()
        # parameter containers with upstream's names; the modules themselves are never called
        self.token_features = nn.Sequential(nn.Linear(num_channels, 512), nn.ReLU(), nn.Linear(512, token_dim))
        self.cluster_features = nn.Sequential(nn.Conv2d(num_channels, 512, 1), drop, nn.ReLU(),
                                              nn.Conv2d(512, cluster_dim, 1))
        self.score = nn.Sequential(nn.Conv2d(num_channels, 512, 1), drop, nn.ReLU(),
                                   nn.Conv2d(512, num_clusters, 1))
        self.dust_bin = nn.Parameter(torch.tensor(1.0))
        self._step = 0
        self.seed = 0x5A1AD
        self._seed_dev = None          # device-side step counter (use_device_step_counter)
        self._grad_hook = None         # (callable(flat_grads) -> work | None, reserve_sms); see set_grad_ready_hook

    def use_device_step_counter(self, enabled: bool = True) -> None:
        """Keep the dropout step counter on the device: the kernels add it to the seed and a one-thread kernel
        advances it after every training forward.  Needed when the forward is captured in a CUDA graph (a host
        counter would be frozen into the graph and every replay would draw the same mask)."""
        self._seed_dev = torch.zeros(1, dtype=torch.int64, device=self.dust_bin.device) if enabled else None

    def set_grad_ready_hook(self, fn, reserve_sms: int = 32, reserve_sms_params: int = 0) -> None:
        """``fn(flat)`` is called inside backward as soon as ALL parameter gradients (one flat fp32 buffer, the
        ``p.grad`` views point into it) have been enqueued and before the input-gradient GEMM is launched; it may
        start an asynchronous collective on ``flat`` and return its work handle (``.wait()`` is called after the
        remaining launches).  ``reserve_sms`` SMs are kept free of the input-gradient GEMM for the collective's
        kernel; ``reserve_sms_params`` SMs are kept free of the persistent kernels that run BEFORE the hook (fused dH
        pass, first-layer weight-gradient GEMM) for a collective the caller started earlier (the geopose-head
        gradients').  ``fn=None`` removes the hook.  See ``parallel.overlap_grad_allreduce``."""
        self._grad_hook = None if fn is None else (fn, int(reserve_sms), int(reserve_sms_params))

    def current_seed(self) -> int:
        """Seed the NEXT training forward will use (synchronises when the counter lives on the device)."""
        base = (self.seed * 1000003 + self._step) & 0xFFFFFFFFFFFFFFFF
        if self._seed_dev is not None:
            base = (base + int(self._seed_dev.item())) & 0xFFFFFFFFFFFFFFFF
        return base

    @property
    def out_dim(self) -> int:
        return self.token_dim + self.cluster_dim * self.num_clusters

    def _param_list(self):
        # direct attribute access: walking named_parameters() costs tens of microseconds per call
        tf, cf, sc = self.token_features, self.cluster_features, self.score
        return [cf[0].weight, cf[0].bias, cf[3].weight, cf[3].bias, sc[0].weight, sc[0].bias, sc[3].weight, sc[3].bias,
                tf[0].weight, tf[0].bias, tf[2].weight, tf[2].bias, self.dust_bin]

    @staticmethod
    def tokens_from_map(f: torch.Tensor) -> torch.Tensor:
        if f.dim() == 3:
            return f.contiguous()
        if f.dim() != 4:
            raise ValueError(f"patch tokens must be [B,C,h,w] or [B,n,C], got {tuple(f.shape)}")
        B, Cc, h, w = f.shape
        return f.permute(0, 2, 3, 1).reshape(B, h * w, Cc).contiguous()   # no copy for channels-last views

    def _prepare(self, x: Tuple[torch.Tensor, torch.Tensor]):
        """Checks and per-call state shared by forward() and GraphedSaladStep: token-major input, shape tuple,
        dropout probability and seed for this call, contiguous fp32 master parameters."""
        f, t = x
        L.require_cuda(f, "patch tokens")
        L.require_cuda(t, "cls token")
        x_tok = self.tokens_from_map(f)
        B, n, d = x_tok.shape
        if d != self.num_channels:
            raise ValueError(f"expected {self.num_channels} channels, got {d}")
        if t.shape != (B, d):
            raise ValueError(f"cls token must be [{B},{d}], got {tuple(t.shape)}")
        if n <= self.num_clusters:
            raise ValueError(f"SALAD needs more patch tokens ({n}) than clusters ({self.num_clusters})")
        t = t.to(x_tok.dtype).contiguous()
        params = self._param_list()
        for p in params:
            if p.dtype != torch.float32 or not p.is_cuda:
                raise RuntimeError("SaladHead parameters must be fp32 CUDA tensors (master weights)")
        p_drop = self.dropout_p if self.training else 0.0
        seed = (self.seed * 1000003 + self._step) & 0xFFFFFFFFFFFFFFFF
        seed_dev = self._seed_dev
        if seed_dev is not None and seed_dev.device != x_tok.device:
            seed_dev = self._seed_dev = seed_dev.to(x_tok.device)
        if self.training and seed_dev is None:
            self._step += 1
        shape = (B, n, d, 512, self.num_clusters, self.cluster_dim, self.token_dim, self.sinkhorn_iters)
        self.last_seed = seed if seed_dev is None else None      # device counter: ask current_seed() before the call
        return x_tok, t, shape, p_drop, seed, seed_dev, [p.contiguous() for p in params]

    def forward(self, x: Tuple[torch.Tensor, torch.Tensor]) -> torch.Tensor:
        x_tok, t, shape, p_drop, seed, seed_dev, params = self._prepare(x)
        return _SaladFn.apply(x_tok, t, shape, p_drop, seed, seed_dev, self._grad_hook, *params)


class GraphedSaladStep:
    """SaladHead forward + backward for one fixed shape, captured once in a CUDA graph and replayed with a single
    launch.  The eager path spends about as long in Python / autograd / 16 kernel launches per step as the B200
    spends computing (~0.45 ms vs ~0.5 ms at batch 256); the replay costs a few microseconds of host time.

        step = GraphedSaladStep(head, batch=256, n_tokens=256)
        desc, dx, dt = step(x_tok, t, d_desc)       # parameter gradients land in p.grad (overwritten each call)

    ``x_tok`` is ``[B, n, C]`` (token-major, the memory behind upstream's ``[B, C, h, w]`` view), ``t`` ``[B, C]``,
    ``d_desc`` ``[B, out_dim]``.  Outputs are static tensors owned by the graph: consume them before the next call.
    Dropout masks still differ from call to call (SaladHead.use_device_step_counter)."""

    def __init__(self, head: "SaladHead", batch: int, n_tokens: int, dtype=torch.bfloat16, input_grads: bool = True,
                 warmup: int = 3, capture_error_mode: str = "thread_local", comm: str | None = None, group=None,
                 comm_ctas: int = 16):
        """``comm="peer"`` (under torchrun): the parameter gradients are summed over the ranks of ``group`` by the
        library's own all-reduce kernel over NVLink peer memory (``comm.PeerComm``) from inside the step -- the flat
        gradient buffer lives in the symmetric allocation and the exchange runs on a forked stream beside the
        input-gradient GEMM, which is launched ``comm_ctas`` CTAs short.  Default: whatever hook the head carries
        (``parallel.overlap_grad_allreduce`` installs the NCCL one)."""
        dev = head.dust_bin.device
        if dev.type != "cuda":
            raise RuntimeError("GraphedSaladStep needs the head on a CUDA device")
        self.head = head
        head.use_device_step_counter(True)
        self.input_grads = bool(input_grads)
        self.x = torch.zeros(batch, n_tokens, head.num_channels, dtype=dtype, device=dev)
        self.t = torch.zeros(batch, head.num_channels, dtype=dtype, device=dev)
        self.d_out = torch.zeros(batch, head.out_dim, dtype=dtype, device=dev)
        lib = L.lib()
        self._flat = None
        self.comm_kind = "hook" if head._grad_hook is not None else "none"
        if comm == "peer":
            import torch.distributed as dist
            if dist.is_initialized() and dist.get_world_size(group) > 1:
                from .comm import PeerComm
                pc = PeerComm(sum(p.numel() for p in head._param_list()), group=group, device=dev)
                comm_stream = torch.cuda.Stream(dev)
                self._comm, self._flat = pc, pc.data
                self.comm_kind = "peer/" + pc.backend + ("+multimem" if pc.mc_ptr else "")
                full = L.COMM_MAX_CTAS if not self.input_grads else int(comm_ctas)

                class _Join:
                    def wait(self_inner):
                        torch.cuda.current_stream(dev).wait_stream(comm_stream)

                def start(flat):
                    if not self.input_grads:                    # nothing left to hide behind: the whole machine
                        pc.allreduce_(0, pc.numel, ctas=min(full, 64))
                        return None
                    comm_stream.wait_stream(torch.cuda.current_stream(dev))
                    with torch.cuda.stream(comm_stream):
                        pc.allreduce_(0, pc.numel, ctas=full)
                    return _Join()

                head.set_grad_ready_hook(start, int(comm_ctas) if self.input_grads else 0, 0)
        # The captured step calls the library directly (no autograd engine inside the capture): the engine keys the
        # stream of every parameter's gradient-accumulation node to the stream that was current when that node was
        # first created, so an eager graph the caller still holds would drag the default stream into the capture.
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(max(1, warmup)):
                self._run()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        launches0 = lib.vpr_launch_count()
        self.graph = torch.cuda.CUDAGraph()
        # thread_local: other threads of the process (NCCL's watchdog polling the events of earlier eager collectives,
        # a data-loader thread copying to the device) may keep making CUDA calls while this thread captures
        with torch.no_grad(), torch.cuda.graph(self.graph, capture_error_mode=capture_error_mode):
            self.out, self.dx, self.dt, self.flat_grad, grads = self._run()
        self.launches_per_replay = int(lib.vpr_launch_count() - launches0)
        plist = head._param_list()
        self.param_grads = []
        self.params = []
        for p, g in zip(plist, grads):
            if p.requires_grad:
                self.params.append(p)
                self.param_grads.append(g.view_as(p))

    def _run(self):
        head = self.head
        x_tok, t, shape, p_drop, seed, seed_dev, params = head._prepare((self.x, self.t))
        out, save = _salad_forward(x_tok, t, shape, p_drop, seed, seed_dev, params)
        dx, dt, flat, grads = _salad_backward(x_tok, t, save, params, shape, p_drop, self.d_out, self.input_grads,
                                              self.input_grads, head._grad_hook, flat=self._flat)
        return out, dx, dt, flat, grads

    def __call__(self, x_tok=None, t=None, d_out=None):
        if x_tok is not None:
            self.x.copy_(x_tok.reshape(self.x.shape), non_blocking=True)
        if t is not None:
            self.t.copy_(t, non_blocking=True)
        if d_out is not None:
            self.d_out.copy_(d_out, non_blocking=True)
        self.graph.replay()
        for p, g in zip(self.params, self.param_grads):
            if p.grad is not g:
                p.grad = g
        return self.out, self.dx, self.dt


def dropout_keep_mask(seed: int, p: float, rows: int, cols: int, device="cuda") -> torch.Tensor:
    """The {0,1} keep mask the fused dropout epilogue used for ``seed`` over the [rows, cols] hidden
    activation (cols = 1024: cluster half then score half).  Test support."""
    out = torch.empty(rows, cols, dtype=torch.uint8, device=device)
    with L.on_device(out):
        L.check(L.lib().vpr_dropout_mask(int(seed), float(p), rows, cols, L.ptr(out), L.stream_ptr(out.device)),
                "vpr_dropout_mask")
    return out


def sinkhorn(S: torch.Tensor, dust_bin: torch.Tensor, iters: int = 3, out_dtype=torch.float32):
    """Stand-alone log-domain Sinkhorn (upstream get_matching_probs + exp, dustbin row dropped).
    S: [B, n, m] fp32 token-major -> (P [B,n,m], U [B,iters,m+1], V [B,iters,n])."""
    L.require_cuda(S, "S")
    S = S.contiguous().float()
    B, n, m = S.shape
    P = torch.empty(B, n, m, dtype=out_dtype, device=S.device)
    U = torch.empty(B, iters, m + 1, dtype=torch.float32, device=S.device)
    V = torch.empty(B, iters, n, dtype=torch.float32, device=S.device)
    db = dust_bin.detach().float().reshape(1).to(S.device)
    with L.on_device(S):
        L.check(L.lib().vpr_sinkhorn_fwd(L.ptr(S), L.ptr(db), L.ptr(P), L.dtype_code(out_dtype), L.ptr(U), L.ptr(V),
                                         B, n, m, iters, L.stream_ptr(S.device)), "vpr_sinkhorn_fwd")
    return P, U, V


def sinkhorn_backward(S, dust_bin, U, V, dP, iters: int = 3, out_dtype=torch.float32):
    S = S.contiguous().float()
    dP = dP.contiguous().float()
    B, n, m = S.shape
    dS = torch.empty(B, n, m, dtype=out_dtype, device=S.device)
    dbias = torch.zeros(m, dtype=torch.float32, device=S.device)
    db = dust_bin.detach().float().reshape(1).to(S.device)
    with L.on_device(S):
        L.check(L.lib().vpr_sinkhorn_bwd(L.ptr(S), L.ptr(db), L.ptr(U), L.ptr(V), L.ptr(dP), L.ptr(dS),
                                         L.dtype_code(out_dtype), L.ptr(dbias), B, n, m, iters,
                                         L.stream_ptr(S.device)), "vpr_sinkhorn_bwd")
    return dS, dbias
