// Sinkhorn forward/backward with the scaled score tile held in REGISTERS by its row owners, m = 64
// clusters, n <= 256 tokens (224 x 224 images: n = 256).
//
// Same factorisation as sinkhorn_scaled.cuh: the tile is exponentiated once,
//     E_jc = 2^(M_jc - mx_j),   mx_j = max(max_c M_jc, dustbin),
// and every logsumexp of an iteration is the log of a dot product of a tile row / column with an
// exponentiated potential vector.  What changes is who holds the tile and how often the CTA meets:
//
//   * 512 threads, two per token row: thread (row j, half h) keeps E_j,32h..32h+31 in 32 registers for
//     the whole launch.  Everything that runs ALONG a row (the v-update's sum over clusters, the final
//     plan, the backward's dv recurrences) reads no shared memory but the 64 potentials.
//   * Sums ALONG a column (the u-update, the backward's du recurrences) read a copy of the tile in
//     shared memory (pitch 68 floats: conflict-free for both the row owners' 16-byte stores and the
//     column owners' 16-byte loads).  Warp w owns clusters 4w..4w+3 for ALL rows: lane = row mod 32,
//     four accumulators, one reduce-scatter over the warp (6 shuffles) -- the column totals never go
//     through a cross-warp partial-sum array, there is no serial "warp 0 finishes the reduction"
//     section, and an iteration costs two block barriers (alpha ready, beta ready) instead of three
//     plus that section.  Warp 0 carries the dustbin column as a fifth accumulator.
//   * The shifts that keep the factors in fp32 range need no reduction: with B0 = max_j mx_j fixed
//     for the launch, alpha_c = 2^(u_c - (norm - B0)) = 1 / r_c and beta_j = 2^(v_j + mx_j - B0)
//     = 1 / s_j, where r_c / s_j are the column / row sums just computed: one MUFU.RCP each instead
//     of log2 -> max over the block -> exp2.  (The potentials themselves, which the backward and the
//     caller get in natural-log units, still take one lg2 per row / column per iteration.)
//   * Backward: the plan-weighted upstream gradient Gt = dP * E * bv is written into the shared tile
//     first; once a row's sum is known its owner turns the row into E*X - Gt in place, so the column
//     sums of the first adjoint step and of G are ONE pass.  The tile then becomes E for the
//     remaining column passes and for the final rank-2K assembly of dS, which keeps the column-owner
//     mapping of sinkhorn_scaled.cuh (coalesced dP re-read / dS store, per-column coefficients in
//     registers, score.3.bias gradient as a by-product).  7 block barriers instead of 12.
//
// Range guard as in sinkhorn_scaled.cuh: every row / column sum is checked; an image whose factors
// leave (1e-20, 1e20) re-runs, in the same CTA, through the element-wise log-domain body
// (sinkhorn_fast.cuh).  Rows >= n are held as zeros.
//
// This header also compiles as plain host C++ under tests/emu/simt_emu.h (VPR_SIMT_EMU): the CPU test
// suite runs these very kernels, one OS thread per CUDA thread, against an fp64 restatement of the algorithm.
#pragma once
#ifndef VPR_SIMT_EMU
#include "sinkhorn_scaled.cuh"
#define VPR_RT_SMEM(name) extern __shared__ __align__(16) float name[]
#endif

namespace vpr {

static constexpr int RT_NT = 512;          // threads per CTA, two per token row (the default); NT = 256: one per row
static constexpr int RT_ROWS = 256;        // token rows a CTA can hold
static constexpr int RT_PITCH = 68;        // floats per row of the shared tile
static constexpr int RT_VEC = 72;          // floats per per-cluster vector (64 clusters + dustbin entries, padded)
static constexpr float RT_LO = 1e-20f, RT_HI = 1e20f;

__host__ __device__ constexpr size_t sinkrt_fwd_floats() {
    return (size_t)RT_ROWS * RT_PITCH + 2 * RT_ROWS + RT_VEC + 32;
}
__host__ __device__ constexpr size_t sinkrt_bwd_floats(int iters) {
    return (size_t)RT_ROWS * RT_PITCH + (size_t)(2 + 2 * iters) * RT_ROWS + (size_t)2 * iters * RT_VEC + 16 + 32 +
           (RT_NT / 32) * 64;       // sized for the larger thread count
}

#ifndef VPR_SIMT_EMU
__device__ __forceinline__ float rt_rcp(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
#endif
__device__ __forceinline__ bool rt_in_range(float x) { return x > RT_LO && x < RT_HI; }
__device__ __forceinline__ float rt_warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float rt_warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Sum a[0..CW) over the 32 lanes of the warp (CW = 4 or 8).  Every lane returns the total of ONE of the values: the one
// with index rt_col_of_lane<CW>(lane)  (reduce-scatter: CW - 1 + log2(32 / CW) shuffles instead of 5 * CW).
template <int CW>
__device__ __forceinline__ int rt_col_of_lane(int lane) {
    return CW == 8 ? ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1) : ((lane >> 4) & 1) * 2 + ((lane >> 3) & 1);
}
template <int CW>
__device__ __forceinline__ float rt_colreduce(float (&a)[CW], int lane) {
    static_assert(CW == 4 || CW == 8, "a warp owns 4 or 8 clusters");
#pragma unroll
    for (int w = CW / 2, bit = 16; w >= 1; w >>= 1, bit >>= 1) {
        const bool up = (lane & bit) != 0;
#pragma unroll
        for (int i = 0; i < w; ++i) {
            const float keep = up ? a[i + w] : a[i];
            const float send = up ? a[i] : a[i + w];
            a[i] = keep + __shfl_xor_sync(0xffffffffu, send, bit);
        }
    }
    float t = a[0];
#pragma unroll
    for (int bit = 16 / CW; bit >= 1; bit >>= 1) t += __shfl_xor_sync(0xffffffffu, t, bit);
    return t;
}

// Column sums of the shared tile for the CW = 2048 / NT clusters this warp owns (CW*warp ..), over all RT_ROWS rows:
//   tot  = sum_j tile_jc * x_j   (UNIT: sum_j tile_jc)   for c = CW*warp + rt_col_of_lane<CW>(lane)
//   dtot = sum_j dcol_j * x_j    (warp 0 only: the dustbin column)
template <int NT, bool UNIT>
__device__ __forceinline__ void rt_col_sums(const float* tile, const float* x, const float* dcol, int warp, int lane,
                                            float& tot, float& dtot) {
    constexpr int CW = 2048 / NT;
    float a[CW], ad = 0.f;
#pragma unroll
    for (int c = 0; c < CW; ++c) a[c] = 0.f;
    const float* tp = tile + lane * RT_PITCH + CW * warp;
#pragma unroll
    for (int i = 0; i < RT_ROWS / 32; ++i) {
        const float xr = x[lane + 32 * i];
#pragma unroll
        for (int q = 0; q < CW / 4; ++q) {
            const float4 t = *reinterpret_cast<const float4*>(tp + i * 32 * RT_PITCH + 4 * q);
            if (UNIT) {
                a[4 * q] += t.x; a[4 * q + 1] += t.y; a[4 * q + 2] += t.z; a[4 * q + 3] += t.w;
            } else {
                a[4 * q] = fmaf(t.x, xr, a[4 * q]); a[4 * q + 1] = fmaf(t.y, xr, a[4 * q + 1]);
                a[4 * q + 2] = fmaf(t.z, xr, a[4 * q + 2]); a[4 * q + 3] = fmaf(t.w, xr, a[4 * q + 3]);
            }
        }
        if (warp == 0) ad = fmaf(dcol[lane + 32 * i], xr, ad);
    }
    tot = rt_colreduce<CW>(a, lane);
    dtot = 0.f;
    if (warp == 0) dtot = rt_warp_sum(ad);
}

// Load this thread's EW = 64 / (threads per row) scores of its row and turn them into E (in place).  Returns mx_j (base-2
// units) and the dustbin entry of the row; rows >= n give zeros.
template <int EW>
__device__ __forceinline__ void rt_load_row(const float* __restrict__ src, bool valid, float z2, float (&e)[EW], float& mx2,
                                            float& dj) {
#pragma unroll
    for (int i = 0; i < EW / 4; ++i) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(src) + i);
        e[4 * i] = v.x; e[4 * i + 1] = v.y; e[4 * i + 2] = v.z; e[4 * i + 3] = v.w;
    }
    float mx = e[0];
#pragma unroll
    for (int i = 1; i < EW; ++i) mx = fmaxf(mx, e[i]);
    if (EW == 32) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 16));
    mx2 = fmaxf(mx * VPR_L2E, z2);
    dj = valid ? fex2(z2 - mx2) : 0.f;
#pragma unroll
    for (int i = 0; i < EW; ++i) e[i] = valid ? fex2(fmaf(e[i], VPR_L2E, -mx2)) : 0.f;
}

template <int EW>
__device__ __forceinline__ void rt_store_row(float* trow, const float (&e)[EW]) {
#pragma unroll
    for (int i = 0; i < EW / 4; ++i)
        *reinterpret_cast<float4*>(trow + 4 * i) = make_float4(e[4 * i], e[4 * i + 1], e[4 * i + 2], e[4 * i + 3]);
}

// sum_c e_c * y_c over this thread's EW clusters (y: their entries in shared memory), completed over the row
template <int EW>
__device__ __forceinline__ float rt_row_dot(const float (&e)[EW], const float* yh) {
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
    for (int i = 0; i < EW / 4; ++i) {
        const float4 y = *reinterpret_cast<const float4*>(yh + 4 * i);
        s0 = fmaf(e[4 * i], y.x, s0); s1 = fmaf(e[4 * i + 1], y.y, s1);
        s2 = fmaf(e[4 * i + 2], y.z, s2); s3 = fmaf(e[4 * i + 3], y.w, s3);
    }
    float s = (s0 + s1) + (s2 + s3);
    if (EW == 32) s += __shfl_xor_sync(0xffffffffu, s, 16);
    return s;
}

// Where a thread sits in the forward: its token row, which part of the row it holds, and what it keeps in registers.
template <int NT>
struct RtFwdState {
    static constexpr int EW = 64 * RT_ROWS / NT;   // tile entries a thread holds (32 or 64)
    float e[EW];        // E_j,c for the thread's clusters
    float bj;           // beta_j = 2^(v_j + mx_j - B0) after the last iteration (0 for rows >= n)
    int row, half;
    bool valid;
};

// Everything of the forward up to the plan: load, exponentiate, iterate.  `sk_smem` holds sinkrt_fwd_floats() floats;
// on return alpha (sk_smem + RT_ROWS * RT_PITCH + 2 * RT_ROWS) holds the last iteration's 2^(u_c - (norm - B0)) and the
// tile region [0, RT_ROWS * RT_PITCH) is dead.  Returns true (block-uniform) if the factors left fp32 range and the
// image has to go through the log-domain body instead.  Ends with a block barrier.
template <int NT>
__device__ __forceinline__ bool rt_fwd_iterate(float* sk_smem, const float* __restrict__ S, const float* __restrict__ dust_bin,
                                               float* __restrict__ U, float* __restrict__ V, int n, int iters, SinkConst2 k,
                                               RtFwdState<NT>& st) {
    static_assert(NT == 512 || NT == 256, "two threads or one thread per token row");
    constexpr int EW = RtFwdState<NT>::EW;
    constexpr int CW = 2048 / NT;                  // clusters a warp owns in the column passes (4 or 8)
    constexpr int NW = NT / 32;
    float* tile = sk_smem;                         // [RT_ROWS][RT_PITCH]  E
    float* beta = tile + RT_ROWS * RT_PITCH;       // [RT_ROWS]  2^(v_j + mx_j - B0)
    float* dcol = beta + RT_ROWS;                  // [RT_ROWS]  dustbin column of E
    float* alpha = dcol + RT_ROWS;                 // [RT_VEC]   2^(u_c - (norm - B0)); [64]: dustbin row
    float* misc = alpha + RT_VEC;                  // [0..15] per-warp max of mx, [16] guard flag
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int row = EW == 32 ? warp * 16 + (lane & 15) : tid, half = EW == 32 ? lane >> 4 : 0;
    const bool valid = row < n;
    st.row = row; st.half = half; st.valid = valid;
    float (&e)[EW] = st.e;
    const float z2 = __ldg(dust_bin) * VPR_L2E;
    float mx2, dj;
    rt_load_row<EW>(S + ((size_t)b * n + (valid ? row : n - 1)) * 64 + half * EW, valid, z2, e, mx2, dj);
    rt_store_row<EW>(tile + row * RT_PITCH + half * EW, e);
    if (half == 0) dcol[row] = dj;
    {
        const float wm = rt_warp_max(valid ? mx2 : -INFINITY);
        if (lane == 0) misc[warp] = wm;
        if (tid == 0) misc[16] = 0.f;
    }
    __syncthreads();
    float B0 = misc[0];
#pragma unroll
    for (int w = 1; w < NW; ++w) B0 = fmaxf(B0, misc[w]);
    float bj = valid ? fex2(mx2 - B0) : 0.f;                     // beta_j with v = 0
    if (half == 0) beta[row] = bj;
    const float kappa = fex2(k.log_a_m2 - k.norm2);              // a_dustbin / a_cluster = n - m
    const int cl = rt_col_of_lane<CW>(lane);
    const bool writer = (lane & (32 / CW - 1)) == 0;
    bool bad = false;
    __syncthreads();

    for (int it = 0; it < iters; ++it) {
        // ---- u_c = log_a_c - LSE_j(M_jc + v_j) = log_a_c - (B0 + log2 r_c),  r_c = sum_j E_jc beta_j ----
        float tot, dtot;
        rt_col_sums<NT, false>(tile, beta, dcol, warp, lane, tot, dtot);
        bad |= !rt_in_range(tot);
        float* Ub = U + ((size_t)b * iters + it) * 65;
        if (writer) {
            alpha[CW * warp + cl] = rt_rcp(tot);
            Ub[CW * warp + cl] = (k.norm2 - B0 - flg2(tot)) * VPR_LN2;
        }
        if (tid == 0) {
            bad |= !rt_in_range(dtot);
            alpha[64] = kappa * rt_rcp(dtot);
            Ub[64] = (k.log_a_m2 - B0 - flg2(dtot)) * VPR_LN2;
        }
        __syncthreads();
        // ---- v_j = log_b - LSE_c(M_jc + u_c) = B0 - mx_j - log2 s_j,  s_j = sum_c E_jc alpha_c ----
        float s = rt_row_dot<EW>(e, alpha + half * EW);
        s = fmaf(dj, alpha[64], s);
        if (!valid) s = 1.f;
        bad |= !rt_in_range(s);
        bj = valid ? rt_rcp(s) : 0.f;
        if (half == 0) {
            beta[row] = bj;
            if (valid) V[((size_t)b * iters + it) * n + row] = (B0 - mx2 - flg2(s)) * VPR_LN2;
        }
        if (bad) misc[16] = 1.f;
        __syncthreads();
    }
    st.bj = bj;
    return misc[16] != 0.f;
}

// Byte offset of the 16-byte unit `q` (clusters 8q .. 8q+7) of token `j` inside the bf16 plan laid out as the
// aggregation MMA's B operand: 128-token chunks of 16 KB, one 128-byte row per token, SWIZZLE_128B (unit ^= row & 7) --
// the bytes a TMA load of the [tokens, 64] bf16 plan with a 64 x 128 box would have produced.
__device__ __forceinline__ uint32_t rt_plan_operand_offset(int j, int q) {
    const int r = j & 127;
    return (uint32_t)((j >> 7) * 16384 + r * 128 + ((q ^ (r & 7)) << 4));
}

// P_jc = 2^(M_jc + u_c + v_j - norm) = E_jc * alpha_c * beta_j: this thread's part of its row to global memory
// (rows < n) and, OPERAND, also as bf16 into the shared-memory MMA operand `op` (rows >= n as zeros up to `op_rows`).
template <typename TP, int NT, bool OPERAND>
__device__ __forceinline__ void rt_fwd_plan(const float* alpha, const RtFwdState<NT>& st, TP* __restrict__ P, int n,
                                            uint8_t* op, int op_rows) {
    constexpr int EW = RtFwdState<NT>::EW;
    static_assert(!OPERAND || sizeof(TP) == 2, "the MMA operand is bf16");
    const int b = blockIdx.x;
    const float bj = st.bj;
    TP* Pr = P + ((size_t)b * n + st.row) * 64 + st.half * EW;
    const float* ah = alpha + st.half * EW;
    if (st.valid || (OPERAND && st.row < op_rows)) {
#pragma unroll
        for (int i = 0; i < EW / 8; ++i) {
            const float4 x = *reinterpret_cast<const float4*>(ah + 8 * i);
            const float4 y = *reinterpret_cast<const float4*>(ah + 8 * i + 4);
            const float p0 = st.e[8 * i] * bj * x.x, p1 = st.e[8 * i + 1] * bj * x.y, p2 = st.e[8 * i + 2] * bj * x.z,
                        p3 = st.e[8 * i + 3] * bj * x.w, p4 = st.e[8 * i + 4] * bj * y.x, p5 = st.e[8 * i + 5] * bj * y.y,
                        p6 = st.e[8 * i + 6] * bj * y.z, p7 = st.e[8 * i + 7] * bj * y.w;
            if constexpr (sizeof(TP) == 2) {
                __nv_bfloat162 q0 = __floats2bfloat162_rn(p0, p1), q1 = __floats2bfloat162_rn(p2, p3),
                               q2 = __floats2bfloat162_rn(p4, p5), q3 = __floats2bfloat162_rn(p6, p7);
                uint4 pk;
                pk.x = *reinterpret_cast<uint32_t*>(&q0); pk.y = *reinterpret_cast<uint32_t*>(&q1);
                pk.z = *reinterpret_cast<uint32_t*>(&q2); pk.w = *reinterpret_cast<uint32_t*>(&q3);
                if (st.valid) *reinterpret_cast<uint4*>(Pr + 8 * i) = pk;
                if constexpr (OPERAND) {
                    if (!st.valid) pk.x = pk.y = pk.z = pk.w = 0u;     // (e and beta are zeros there anyway)
                    *reinterpret_cast<uint4*>(op + rt_plan_operand_offset(st.row, st.half * (EW / 8) + i)) = pk;
                }
            } else {
                *reinterpret_cast<float4*>(Pr + 8 * i) = make_float4(p0, p1, p2, p3);
                *reinterpret_cast<float4*>(Pr + 8 * i + 4) = make_float4(p4, p5, p6, p7);
            }
        }
    }
}

template <typename TP, int NT>
__global__ void __launch_bounds__(NT, 2) sinkhorn_fwd_rt(const float* __restrict__ S, const float* __restrict__ dust_bin,
                                                         TP* __restrict__ P, float* __restrict__ U, float* __restrict__ V,
                                                         int n, int iters, SinkConst2 k) {
    VPR_RT_SMEM(sk_smem);
    pdl_wait();
    pdl_launch_dependents();
    RtFwdState<NT> st;
    if (rt_fwd_iterate<NT>(sk_smem, S, dust_bin, U, V, n, iters, k, st)) {     // block-uniform: out of fp32 range -> exact path
#ifdef VPR_SIMT_EMU
        simt_emu::note_guard();
#else
        __syncthreads();
        sinkhorn_fwd_logdomain<TP, 64, NT>(sk_smem, S, dust_bin, P, U, V, n, iters, k);
#endif
        return;
    }
    rt_fwd_plan<TP, NT, false>(sk_smem + RT_ROWS * RT_PITCH + 2 * RT_ROWS, st, P, n, nullptr, 0);
}

// K: Sinkhorn iteration count (compile time: the per-iteration row factors live in registers).
template <typename TD, int K, int NT>
__global__ void __launch_bounds__(NT, 2) sinkhorn_bwd_rt(const float* __restrict__ S, const float* __restrict__ dust_bin,
                                                         const float* __restrict__ U, const float* __restrict__ V,
                                                         const float* __restrict__ dP, TD* __restrict__ dS,
                                                         float* __restrict__ d_bias, int n, SinkConst2 k) {
    static_assert(NT == 512 || NT == 256, "two threads or one thread per token row");
    constexpr int EW = 64 * RT_ROWS / NT, CW = 2048 / NT, NW = NT / 32;
    VPR_RT_SMEM(sk_smem);
    float* tile = sk_smem;                         // [RT_ROWS][RT_PITCH]  E*X - Gt first, E afterwards
    float* dcol = tile + RT_ROWS * RT_PITCH;       // [RT_ROWS]  dustbin column of E
    float* bvl = dcol + RT_ROWS;                   // [RT_ROWS]  bv^{K-1}_j
    float* xs = bvl + RT_ROWS;                     // [K][RT_ROWS]  X^k_j = dv^k_j * bv^k_j
    float* bus = xs + K * RT_ROWS;                 // [K][RT_ROWS]  bu^k_j
    float* aK = bus + K * RT_ROWS;                 // [K][RT_VEC]  a^k_c ; [64] dustbin under Wv ; [65] dustbin under Wu
    float* Y = aK + K * RT_VEC;                    // [K][RT_VEC]  du^k_c * a^k_c ; [64] dustbin
    float* Ak = Y + K * RT_VEC;                    // [16]  shift of iteration k
    float* misc = Ak + 16;                         // [0] guard flag
    float* red = misc + 32;                        // [NW][64]  score-bias partial sums
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int row = EW == 32 ? warp * 16 + (lane & 15) : tid, half = EW == 32 ? lane >> 4 : 0;
    const bool valid = row < n;
    const int rowc = valid ? row : n - 1;
    pdl_wait();
    pdl_launch_dependents();
    const float z2 = __ldg(dust_bin) * VPR_L2E;
    float e[EW], mx2, dj;
    rt_load_row<EW>(S + ((size_t)b * n + rowc) * 64 + half * EW, valid, z2, e, mx2, dj);
    float vk[K];
#pragma unroll
    for (int kk = 0; kk < K; ++kk) vk[kk] = __ldg(V + ((size_t)b * K + kk) * n + rowc) * VPR_L2E;
    // per-iteration cluster factors: warp kk handles iteration kk
    for (int kk = warp; kk < K; kk += NW) {
        const float* Ub = U + ((size_t)b * K + kk) * 65;
        const float u0 = __ldg(Ub + lane) * VPR_L2E, u1 = __ldg(Ub + 32 + lane) * VPR_L2E, um = __ldg(Ub + 64) * VPR_L2E;
        const float A = rt_warp_max(fmaxf(fmaxf(u0, u1), um));
        aK[kk * RT_VEC + lane] = fex2(u0 - A);
        aK[kk * RT_VEC + 32 + lane] = fex2(u1 - A);
        if (lane == 0) {
            aK[kk * RT_VEC + 64] = fex2(um - A);
            aK[kk * RT_VEC + 65] = fex2(um - A + (k.norm2 - k.log_a_m2));
            Ak[kk] = A;
        }
    }
    if (tid == 0) misc[0] = 0.f;
    if (half == 0) dcol[row] = dj;
    // the first four 16-byte pieces of this row's dP are requested before the barrier: their latency hides behind it
    const float4* dsrc = reinterpret_cast<const float4*>(dP + ((size_t)b * n + rowc) * 64 + half * EW);
    float4 d[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) d[u] = __ldg(dsrc + u);
    __syncthreads();                                             // aK, Ak
    // per-iteration token factors: Wv^k = E a^k_c bv^k_j, Wu^k = E a^k_c bu^k_j
    float bv[K], bu[K];
    {
        const float mj = mx2 - k.norm2;
        float vprev = 0.f;
        bool bad = false;
#pragma unroll
        for (int kk = 0; kk < K; ++kk) {
            bv[kk] = fex2(vk[kk] + mj + Ak[kk]);
            bu[kk] = fex2(vprev + mj + Ak[kk]);
            bad |= !(bv[kk] < RT_HI && bu[kk] < RT_HI);
            vprev = vk[kk];
            if (!valid) { bv[kk] = 0.f; bu[kk] = 0.f; }
        }
        if (bad && valid) misc[0] = 1.f;
        if (half == 0) {
#pragma unroll
            for (int kk = 0; kk < K; ++kk) bus[kk * RT_ROWS + row] = bu[kk];
            bvl[row] = bv[K - 1];
        }
    }
    // ---- G = dP * P (P = Wv^{K-1}):  Gt_jc = dP_jc E_jc bv_j into the tile,  dv_j = bv_j * sum_c dP_jc E_jc a_c ----
    float* trow = tile + row * RT_PITCH + half * EW;
    const float bvL = bv[K - 1];
    {
        const float* aL = aK + (K - 1) * RT_VEC + half * EW;
        float rs0 = 0.f, rs1 = 0.f;
#pragma unroll
        for (int h4 = 0; h4 < EW / 16; ++h4) {
            if (h4 > 0) {
#pragma unroll
                for (int u = 0; u < 4; ++u) d[u] = __ldg(dsrc + 4 * h4 + u);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = h4 * 4 + u;
                const float4 a = *reinterpret_cast<const float4*>(aL + 4 * i);
                const float t0 = d[u].x * e[4 * i], t1 = d[u].y * e[4 * i + 1], t2 = d[u].z * e[4 * i + 2],
                            t3 = d[u].w * e[4 * i + 3];
                rs0 = fmaf(t0, a.x, rs0); rs1 = fmaf(t1, a.y, rs1); rs0 = fmaf(t2, a.z, rs0); rs1 = fmaf(t3, a.w, rs1);
                *reinterpret_cast<float4*>(trow + 4 * i) = make_float4(t0 * bvL, t1 * bvL, t2 * bvL, t3 * bvL);
            }
        }
        float rs = rs0 + rs1;
        if (EW == 32) rs += __shfl_xor_sync(0xffffffffu, rs, 16);
        const float Xl = rs * bvL * bvL;                         // dv^{K-1}_j * bv^{K-1}_j
        // the row becomes E*X - Gt: the column sums of the first adjoint step and of G are one pass
#pragma unroll
        for (int i = 0; i < EW / 4; ++i) {
            float4 t = *reinterpret_cast<const float4*>(trow + 4 * i);
            t.x = fmaf(e[4 * i], Xl, -t.x); t.y = fmaf(e[4 * i + 1], Xl, -t.y);
            t.z = fmaf(e[4 * i + 2], Xl, -t.z); t.w = fmaf(e[4 * i + 3], Xl, -t.w);
            *reinterpret_cast<float4*>(trow + 4 * i) = t;
        }
        if (half == 0) xs[(K - 1) * RT_ROWS + row] = Xl;
    }
    __syncthreads();                                             // tile, xs[K-1], bus, bvl, guard flag
    if (misc[0] != 0.f) {
#ifdef VPR_SIMT_EMU
        simt_emu::note_guard();
#else
        __syncthreads();
        sinkhorn_bwd_logdomain<TD, 64, NT>(sk_smem, S, dust_bin, U, V, dP, dS, d_bias, n, K, k);
#endif
        return;
    }
    const int cl = rt_col_of_lane<CW>(lane);
    const bool writer = (lane & (32 / CW - 1)) == 0;
#pragma unroll
    for (int kk = K - 1; kk >= 0; --kk) {
        const float* ak = aK + kk * RT_VEC;
        // du^k_c = [k = K-1: sum_j G_jc] - sum_j dv^k_j Wv^k_jc = -a^k_c * sum_j (E_jc X^k_j [- Gt_jc])
        float tot, dtot;
        if (kk == K - 1) rt_col_sums<NT, true>(tile, xs + kk * RT_ROWS, dcol, warp, lane, tot, dtot);
        else rt_col_sums<NT, false>(tile, xs + kk * RT_ROWS, dcol, warp, lane, tot, dtot);
        if (writer) {
            const float ac = ak[CW * warp + cl];
            Y[kk * RT_VEC + CW * warp + cl] = -ac * ac * tot;
        }
        if (tid == 0) Y[kk * RT_VEC + 64] = -ak[64] * dtot * ak[65];
        __syncthreads();                                         // Y^k ready, every read of the tile done
        if (kk == K - 1) rt_store_row<EW>(trow, e);              // the tile is E from here on
        if (kk > 0) {
            // dv^{k-1}_j = - sum_c du^k_c Wu^k_jc = - bu^k_j * sum_c E_jc y^k_c
            float s = rt_row_dot<EW>(e, Y + kk * RT_VEC + half * EW);
            s = fmaf(dj, Y[kk * RT_VEC + 64], s);
            const float Xp = -bu[kk] * s * bv[kk - 1];
            if (half == 0) xs[(kk - 1) * RT_ROWS + row] = Xp;
            __syncthreads();                                     // xs[k-1] (and, the first time, the E tile)
        } else if (K == 1) {
            __syncthreads();
        }
    }
    // ---- dM_jc = G_jc - sum_k (dv^k_j Wv^k_jc + du^k_c Wu^k_jc)
    //            = E_jc * (dP_jc a^{K-1}_c bv^{K-1}_j - sum_k (a^k_c x^k_j + y^k_c bu^k_j)) ----
    // column-owner mapping: thread = (4 clusters cg, every (NT/16)-th row), per-cluster coefficients in registers,
    // dP / dS coalesced
    {
        constexpr int RSTEP = NT / 16, RPT = RT_ROWS / RSTEP;    // rows per thread: 8 or 16
        const int cg = tid & 15, jr = tid >> 4;
        float4 ar[K], yr[K];
#pragma unroll
        for (int kk = 0; kk < K; ++kk) {
            ar[kk] = *reinterpret_cast<const float4*>(aK + kk * RT_VEC + cg * 4);
            yr[kk] = *reinterpret_cast<const float4*>(Y + kk * RT_VEC + cg * 4);
        }
        const float4 a4 = ar[K - 1];
        const float* dPb = dP + (size_t)b * n * 64;
        TD* dSb = dS + (size_t)b * n * 64;
        float4 bacc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int h4 = 0; h4 < RPT / 4; ++h4) {
            float4 g[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int j = (h4 * 4 + u) * RSTEP + jr;
                g[u] = __ldg(reinterpret_cast<const float4*>(dPb + (size_t)(j < n ? j : n - 1) * 64) + cg);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int j = (h4 * 4 + u) * RSTEP + jr;
                if (j < n) {
                    const float4 e4 = *reinterpret_cast<const float4*>(tile + j * RT_PITCH + cg * 4);
                    const float bjl = bvl[j];
                    float c0 = 0.f, c1 = 0.f, c2 = 0.f, c3 = 0.f;
#pragma unroll
                    for (int kk = 0; kk < K; ++kk) {
                        const float xj = xs[kk * RT_ROWS + j], uj = bus[kk * RT_ROWS + j];
                        c0 = fmaf(ar[kk].x, xj, fmaf(yr[kk].x, uj, c0)); c1 = fmaf(ar[kk].y, xj, fmaf(yr[kk].y, uj, c1));
                        c2 = fmaf(ar[kk].z, xj, fmaf(yr[kk].z, uj, c2)); c3 = fmaf(ar[kk].w, xj, fmaf(yr[kk].w, uj, c3));
                    }
                    const float g0 = e4.x * (g[u].x * a4.x * bjl - c0), g1 = e4.y * (g[u].y * a4.y * bjl - c1);
                    const float g2 = e4.z * (g[u].z * a4.z * bjl - c2), g3 = e4.w * (g[u].w * a4.w * bjl - c3);
                    if constexpr (sizeof(TD) == 2) {
                        __nv_bfloat162 lo = __floats2bfloat162_rn(g0, g1), hi = __floats2bfloat162_rn(g2, g3);
                        uint2 pk;
                        pk.x = *reinterpret_cast<uint32_t*>(&lo);
                        pk.y = *reinterpret_cast<uint32_t*>(&hi);
                        *reinterpret_cast<uint2*>(dSb + (size_t)j * 64 + cg * 4) = pk;
                    } else {
                        *reinterpret_cast<float4*>(dSb + (size_t)j * 64 + cg * 4) = make_float4(g0, g1, g2, g3);
                    }
                    bacc.x += g0; bacc.y += g1; bacc.z += g2; bacc.w += g3;
                }
            }
        }
        if (d_bias != nullptr) {
            bacc.x += __shfl_xor_sync(0xffffffffu, bacc.x, 16); bacc.y += __shfl_xor_sync(0xffffffffu, bacc.y, 16);
            bacc.z += __shfl_xor_sync(0xffffffffu, bacc.z, 16); bacc.w += __shfl_xor_sync(0xffffffffu, bacc.w, 16);
            if (lane < 16) *reinterpret_cast<float4*>(red + warp * 64 + cg * 4) = bacc;
            __syncthreads();
            if (tid < 64) {
                float t = 0.f;
#pragma unroll
                for (int w = 0; w < NW; ++w) t += red[w * 64 + tid];
                atomicAdd(d_bias + tid, t);
            }
        }
    }
}

}  // namespace vpr
