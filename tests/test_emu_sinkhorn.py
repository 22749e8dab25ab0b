"""CPU runs of the register-tile Sinkhorn kernels' DEVICE SOURCE (csrc/sinkhorn_rt.cuh) against the fp64 oracle.

tests/emu/simt_emu.h compiles the kernels as host C++ (one OS thread per CUDA thread, pthread barriers for
__syncthreads and the warp shuffles, shared memory poisoned with NaNs before every block).  What this covers
without a GPU: the index arithmetic of both thread mappings, the reduce-scatter over the warp, the barrier
placement as far as values are concerned (a read of a slot nobody wrote yet is a NaN), the n < 256 masking,
the range-guard decision and the algebra of the merged first adjoint step.  What it cannot cover (bank
conflicts, register allocation, memory-model races that a CPU's stronger ordering hides) is left to the GPU
parity tests, which run the same shapes through the library.
"""
import ctypes
import math
import os
import subprocess

import pytest
import torch

from oracle import salad_ref as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "emu", "sinkhorn_rt_emu.cpp")


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("emu") / "libsinkrt_emu.so")
    r = subprocess.run(["g++", "-std=c++17", "-O1", "-pthread", "-shared", "-fPIC", "-Wno-unknown-pragmas", "-o", out, SRC],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    lib = ctypes.CDLL(out)
    lib.emu_sinkrt_fwd_bytes.restype = ctypes.c_size_t
    lib.emu_sinkrt_bwd_bytes.restype = ctypes.c_size_t
    return lib


def ptr(t):
    return ctypes.c_void_p(t.data_ptr())


def rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm()).item()


def potentials_fp64(S, dust, iters):
    """u, v after every iteration (natural-log units), the upstream update order."""
    B, m, n = S.shape
    M = torch.cat([S.double(), torch.full((B, 1, n), float(dust), dtype=torch.float64)], 1)
    norm = -math.log(n + m)
    log_a = torch.full((B, m + 1), norm, dtype=torch.float64)
    log_a[:, -1] += math.log(n - m)
    u = torch.zeros(B, m + 1, dtype=torch.float64)
    v = torch.zeros(B, n, dtype=torch.float64)
    us, vs = [], []
    for _ in range(iters):
        u = log_a - torch.logsumexp(M + v.unsqueeze(1), dim=2)
        v = norm - torch.logsumexp(M + u.unsqueeze(2), dim=1)
        us.append(u)
        vs.append(v)
    return torch.stack(us, 1), torch.stack(vs, 1)


def run_fwd(emu, St, db, iters, bf16=False, nt=512):
    B, n, _ = St.shape
    P = torch.full((B, n, 64), float("nan"), dtype=torch.bfloat16 if bf16 else torch.float32)
    U = torch.full((B, iters, 65), float("nan"))
    V = torch.full((B, iters, n), float("nan"))
    guards = emu.emu_sinkhorn_fwd_rt(ptr(St), ptr(db), ptr(P), int(bf16), ptr(U), ptr(V), B, n, iters, nt)
    return P, U, V, guards


@pytest.mark.parametrize("nt", [512, 256])
@pytest.mark.parametrize("B,n,scale,dust", [(2, 256, 2.0, 1.0), (2, 100, 2.0, 1.0), (1, 65, 2.0, -3.0), (2, 200, 6.0, 0.5),
                                             (1, 255, 1.0, 9.0)])
def test_register_tile_kernels_match_the_fp64_oracle(emu, B, n, scale, dust, nt):
    g = torch.Generator().manual_seed(n)
    S = torch.randn(B, 64, n, generator=g) * scale
    db = torch.tensor([dust])
    Sd = S.double().requires_grad_(True)
    Pd = O.get_matching_probs(Sd, db.double()[0], 3, const_dtype=torch.float64).exp()
    St = S.transpose(1, 2).contiguous()
    P, U, V, guards = run_fwd(emu, St, db, 3, nt=nt)
    assert guards == 0
    assert rel(P, Pd[:, :64].transpose(1, 2)) < 1e-5
    u64, v64 = potentials_fp64(S, dust, 3)
    assert (U.double() - u64).abs().max() < 2e-5 and (V.double() - v64).abs().max() < 2e-5
    dP = torch.randn(B, n, 64, generator=g)
    (Pd[:, :64] * dP.double().transpose(1, 2)).sum().backward()
    dS = torch.full((B, n, 64), float("nan"))
    d_bias = torch.zeros(64)
    guards = emu.emu_sinkhorn_bwd_rt3(ptr(St), ptr(db), ptr(U), ptr(V), ptr(dP), ptr(dS), 0, ptr(d_bias), B, n, nt)
    assert guards == 0
    assert rel(dS, Sd.grad.transpose(1, 2)) < 1e-5
    # explicit recurrence (SURVEY A-3) agrees too; a per-cluster score offset cannot change the plan
    assert rel(dS, O.sinkhorn_backward_explicit(S.double(), dP.double().transpose(1, 2), 3).transpose(1, 2)) < 1e-5
    assert d_bias.abs().max() < 1e-4


@pytest.mark.parametrize("nt", [512, 256])
def test_register_tile_forward_other_iteration_counts_and_bf16_outputs(emu, nt):
    g = torch.Generator().manual_seed(11)
    S = torch.randn(2, 64, 256, generator=g) * 2.0
    db = torch.tensor([1.0])
    St = S.transpose(1, 2).contiguous()
    for iters in (1, 5, 8):
        Pd = O.get_matching_probs(S.double(), db.double()[0], iters, const_dtype=torch.float64).exp()
        P, U, V, guards = run_fwd(emu, St, db, iters, nt=nt)
        assert guards == 0 and rel(P, Pd[:, :64].transpose(1, 2)) < 1e-5, iters
    Pd = O.get_matching_probs(S.double(), db.double()[0], 3, const_dtype=torch.float64).exp()[:, :64].transpose(1, 2)
    P32, U, V, _ = run_fwd(emu, St, db, 3, nt=nt)
    Pb, _, _, _ = run_fwd(emu, St, db, 3, bf16=True, nt=nt)
    assert torch.equal(Pb, P32.bfloat16())                       # same values, rounded once (round to nearest even)
    dP = torch.randn(2, 256, 64, generator=g)
    dS32 = torch.full((2, 256, 64), float("nan"))
    dSb = torch.full((2, 256, 64), float("nan"), dtype=torch.bfloat16)
    d_bias = torch.zeros(64)
    emu.emu_sinkhorn_bwd_rt3(ptr(St), ptr(db), ptr(U), ptr(V), ptr(dP), ptr(dS32), 0, ptr(d_bias), 2, 256, nt)
    emu.emu_sinkhorn_bwd_rt3(ptr(St), ptr(db), ptr(U), ptr(V), ptr(dP), ptr(dSb), 1, ptr(d_bias), 2, 256, nt)
    assert torch.isfinite(dS32).all() and rel(dSb, dS32) < 4e-3


@pytest.mark.parametrize("nt", [512, 256])
def test_register_tile_range_guard_asks_for_the_log_domain_body(emu, nt):
    """Image 1 spans several hundred nats: its factors leave fp32 range, the CTA must hand the image to the
    log-domain body (compiled out under the emulation: the entry point counts the request)."""
    g = torch.Generator().manual_seed(5)
    S = torch.randn(2, 64, 256, generator=g) * 2.0
    S[1] *= 40.0
    db = torch.tensor([1.0])
    St = S.transpose(1, 2).contiguous()
    P, U, V, guards = run_fwd(emu, St, db, 3, nt=nt)
    assert guards == 1
    Pd = O.get_matching_probs(S.double(), db.double()[0], 3, const_dtype=torch.float64).exp()
    assert rel(P[0], Pd[0, :64].transpose(0, 1)) < 1e-5 and torch.isnan(P[1]).all()
    # backward: potentials of the wide image from the fp64 recurrence; the same decision is taken there
    u64, v64 = potentials_fp64(S, 1.0, 3)
    dP = torch.randn(2, 256, 64, generator=g)
    dS = torch.full((2, 256, 64), float("nan"))
    U32, V32, d_bias = u64.float().contiguous(), v64.float().contiguous(), torch.zeros(64)
    guards = emu.emu_sinkhorn_bwd_rt3(ptr(St), ptr(db), ptr(U32), ptr(V32), ptr(dP), ptr(dS), 0, ptr(d_bias), 2, 256, nt)
    assert guards == 1 and torch.isfinite(dS[0]).all() and torch.isnan(dS[1]).all()


def test_register_tile_kernels_on_seeded_random_shapes(emu):
    """Eight seeded draws of (n, score scale, dustbin, thread count, images): forward and backward against the fp64 oracle,
    plus the marginal identity the solver's last update makes exact (token rows sum to 1 - dustbin share)."""
    rng = torch.Generator().manual_seed(20261018)
    for _ in range(8):
        n = int(torch.randint(65, 257, (1,), generator=rng))
        scale = float(torch.rand(1, generator=rng) * 5 + 0.2)
        dust = float(torch.randn(1, generator=rng) * 3)
        nt = 512 if int(torch.randint(0, 2, (1,), generator=rng)) else 256
        B = int(torch.randint(1, 4, (1,), generator=rng))
        S = torch.randn(B, 64, n, generator=rng) * scale
        db = torch.tensor([dust])
        Sd = S.double().requires_grad_(True)
        Pd = O.get_matching_probs(Sd, db.double()[0], 3, const_dtype=torch.float64).exp()
        St = S.transpose(1, 2).contiguous()
        P, U, V, guards = run_fwd(emu, St, db, 3, nt=nt)
        tag = (n, scale, dust, nt, B)
        assert guards == 0, tag
        assert rel(P, Pd[:, :64].transpose(1, 2)) < 1e-5, tag
        assert (P.double().sum(2) - (1 - Pd[:, 64])).abs().max() < 1e-5, tag      # token marginal
        dP = torch.randn(B, n, 64, generator=rng)
        (Pd[:, :64] * dP.double().transpose(1, 2)).sum().backward()
        dS = torch.full((B, n, 64), float("nan"))
        d_bias = torch.zeros(64)
        assert emu.emu_sinkhorn_bwd_rt3(ptr(St), ptr(db), ptr(U), ptr(V), ptr(dP), ptr(dS), 0, ptr(d_bias), B, n, nt) == 0, tag
        assert rel(dS, Sd.grad.transpose(1, 2)) < 2e-5, tag


@pytest.mark.parametrize("n", [256, 196, 100])
def test_plan_written_as_the_aggregation_mma_operand(emu, n):
    """The fused Sinkhorn + aggregation kernel never reads the plan back: the row owners write it, as bf16, straight into
    the SWIZZLE_128B layout the tcgen05 B operand expects (what a TMA load of the [tokens, 64] plan with a 64 x 128 box
    would have produced): 128-token chunks of 16 KB, one 128-byte row per token, 16-byte unit q at q ^ (row & 7)."""
    g = torch.Generator().manual_seed(n + 7)
    S = (torch.randn(2, 64, n, generator=g) * 2.0).transpose(1, 2).contiguous()
    db = torch.tensor([1.0])
    P = torch.full((2, n, 64), float("nan"), dtype=torch.bfloat16)
    U = torch.empty(2, 3, 65)
    V = torch.empty(2, 3, n)
    rows = (n + 127) // 128 * 128
    op = torch.full((2, rows * 128), 0xAB, dtype=torch.uint8)
    assert emu.emu_sinkhorn_fwd_rt_operand(ptr(S), ptr(db), ptr(P), ptr(U), ptr(V), ptr(op), 2, n, 3) == 0
    Pref, _, _, _ = run_fwd(emu, S, db, 3, bf16=True)
    assert torch.equal(P, Pref)
    want = torch.zeros(2, rows, 64, dtype=torch.bfloat16)
    want[:, :n] = Pref
    units = want.view(torch.uint8).reshape(2, rows // 128, 128, 8, 16)        # [image, chunk, row, unit, 16 bytes]
    r = torch.arange(128)
    q = torch.arange(8)
    phys = (q[None, :] ^ (r[:, None] & 7))                                      # unit q of row r lives at unit q ^ (r & 7)
    laid = torch.empty_like(units)
    laid[:, :, r[:, None], phys] = units[:, :, r[:, None], q[None, :]]
    assert torch.equal(op, laid.reshape(2, rows * 128))


def test_register_tile_shared_memory_budget_keeps_two_ctas_per_sm(emu):
    # 227 KB usable per SM, 1 KB reserved per CTA: two CTAs of either kernel must fit side by side
    assert 2 * (emu.emu_sinkrt_fwd_bytes() + 1024) <= 227 * 1024
    assert 2 * (emu.emu_sinkrt_bwd_bytes(3) + 1024) <= 227 * 1024
