// Sinkhorn forward/backward on a scaled (exponentiated once) score tile, m = 64 clusters.
//
// The potentials u, v stay in the (base-2) log domain exactly as upstream's log_otp_solver keeps
// them, but the [n, m] tile held in shared memory across all iterations is
//     E_jc = 2^(M_jc - mx_j),   mx_j = max(max_c M_jc, dustbin)          (0 < E <= 1)
// so that every logsumexp of an iteration becomes the log of a dot product of a tile row / column
// with a max-shifted exponentiated potential vector:
//     LSE_j(M_jc + v_j) = B + log2 sum_j E_jc * 2^(v_j + mx_j - B),   B = max_j (v_j + mx_j)
//     LSE_c(M_jc + u_c) = mx_j + A + log2 sum_c E_jc * 2^(u_c - A),   A = max_c u_c
// One exp2 per tile element in total (instead of one per element per pass); the passes are FMA
// streams over the tile (128-bit shared-memory loads, bank-conflict free at a pitch of 80 floats)
// with warp-shuffle + shared-memory reductions.  The backward uses the same
// factorisation: every per-iteration plan Wv^k / Wu^k is E_jc * a^k_c * b^k_j, so the adjoint
// recurrence is row / column dot products and the final dS assembly is a rank-2K update.
//
// Range guard: the shifted factors are exact as long as no row/column sum under- or overflows
// fp32 (score spreads beyond ~60 nats).  Every sum is checked; if one is out of range the CTA
// re-runs its image through the element-wise log-domain body (sinkhorn_fast.cuh), which has no
// such limit.  Results are identical to fp32 rounding either way.
//
// Launch: one CTA of 512 threads per image, 2 CTAs per SM at n = 256.  HBM traffic per image:
// forward reads S (n*m*4 B) and writes P + potentials; backward reads S, dP and writes dS.
#pragma once
#include "sinkhorn_fast.cuh"

namespace vpr {

static constexpr int SS_NT = SK_NT;       // 512 threads
static constexpr int SS_PITCH = 80;       // floats per token row of the tile (64 used)
static constexpr int SS_LDB = 8;          // 16-byte global loads in flight per thread
static constexpr int SS_RED = 68;         // pitch of the cross-warp partial-sum rows (64 clusters + dustbin)
static constexpr float SS_LO = 1e-20f, SS_HI = 1e20f;   // flushed-to-zero tile entries then cost < 1e-18 absolute

__host__ __device__ constexpr size_t sinkscaled_fwd_floats(int n) {
    return (size_t)n * SS_PITCH + 3 * (size_t)((n + 3) & ~3) + 2 * SS_RED + (SS_NT / 32) * SS_RED + 32 + 8;
}
__host__ __device__ constexpr size_t sinkscaled_bwd_floats(int n, int iters) {
    return (size_t)n * SS_PITCH + (2 + 3 * (size_t)iters) * (size_t)((n + 3) & ~3) + 2 * (size_t)iters * SS_RED + 16 +
           (SS_NT / 32) * SS_RED + 32 + 8;
}

__device__ __forceinline__ bool ss_in_range(float x) { return x > SS_LO && x < SS_HI; }

// Load one image's scores, build E / mx / dustbin column.  Returns this thread's max of mx_j.
__device__ __forceinline__ float ss_load_tile(const float* __restrict__ Sb, float z2, int n, float* E, float* mxs,
                                              float* dst) {
    const int tid = threadIdx.x;
    float bmax = -INFINITY;
    const int total = n * 16;
    // all global loads of a batch are issued before any of them is consumed: the tile arrives in one
    // memory round trip instead of one per 512 x 16 bytes
    for (int b0 = 0; b0 < total; b0 += SS_LDB * SS_NT) {
        float4 buf[SS_LDB];
#pragma unroll
        for (int u = 0; u < SS_LDB; ++u) {
            const int idx = b0 + u * SS_NT + tid;
            const int j = idx < total ? (idx >> 4) : (n - 1);
            buf[u] = __ldg(reinterpret_cast<const float4*>(Sb + j * 64) + (idx & 15));
        }
#pragma unroll
        for (int u = 0; u < SS_LDB; ++u) {
            const int idx = b0 + u * SS_NT + tid;
            if (b0 + u * SS_NT >= total) break;                      // block-uniform
            const bool valid = idx < total;
            const int j = valid ? (idx >> 4) : (n - 1);
            const int cg = idx & 15;
            float4 s4 = buf[u];
            s4.x *= VPR_L2E; s4.y *= VPR_L2E; s4.z *= VPR_L2E; s4.w *= VPR_L2E;
            float mx = fmaxf(fmaxf(s4.x, s4.y), fmaxf(s4.z, s4.w));
#pragma unroll
            for (int o = 1; o < 16; o <<= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            mx = fmaxf(mx, z2);
            if (valid) {
                float4 e4;
                e4.x = fex2(s4.x - mx); e4.y = fex2(s4.y - mx); e4.z = fex2(s4.z - mx); e4.w = fex2(s4.w - mx);
                *reinterpret_cast<float4*>(E + j * SS_PITCH + cg * 4) = e4;
                if (cg == 0) { mxs[j] = mx; dst[j] = fex2(z2 - mx); }
                bmax = fmaxf(bmax, mx);
            }
        }
    }
    return bmax;
}

// acc[c-chunk of this thread] = sum_j E_jc * x_j  (+ dustbin column), reduced over the CTA into
// red[warp][0..64].  Thread = (chunk of 4 clusters, 1/32 of the tokens).  Caller syncs afterwards.
__device__ __forceinline__ void ss_row_pass(const float* E, const float* dst, const float* x, float xscale, int n,
                                            float* red) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int cg = tid & 15, part = tid >> 4;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    float accd = 0.f;
#pragma unroll 4
    for (int j = part; j < n; j += SS_NT / 16) {
        const float xj = x[j];
        const float4 e4 = *reinterpret_cast<const float4*>(E + j * SS_PITCH + cg * 4);
        acc.x = fmaf(e4.x, xj, acc.x); acc.y = fmaf(e4.y, xj, acc.y);
        acc.z = fmaf(e4.z, xj, acc.z); acc.w = fmaf(e4.w, xj, acc.w);
        if (cg == 0) accd = fmaf(dst[j], xj, accd);
    }
    acc.x += __shfl_xor_sync(0xffffffffu, acc.x, 16); acc.y += __shfl_xor_sync(0xffffffffu, acc.y, 16);
    acc.z += __shfl_xor_sync(0xffffffffu, acc.z, 16); acc.w += __shfl_xor_sync(0xffffffffu, acc.w, 16);
    accd += __shfl_xor_sync(0xffffffffu, accd, 16);
    if (lane < 16) {
        acc.x *= xscale; acc.y *= xscale; acc.z *= xscale; acc.w *= xscale;
        *reinterpret_cast<float4*>(red + warp * SS_RED + cg * 4) = acc;
        if (lane == 0) red[warp * SS_RED + 64] = accd * xscale;
    }
}

// warp 0 after a row pass: totals for clusters lane, lane + 32 and (lane 0) the dustbin
__device__ __forceinline__ void ss_row_totals(const float* red, int lane, float& t0, float& t1, float& tm) {
    t0 = 0.f; t1 = 0.f; tm = 0.f;
#pragma unroll
    for (int w = 0; w < SS_NT / 32; ++w) {
        t0 += red[w * SS_RED + lane];
        t1 += red[w * SS_RED + 32 + lane];
        tm += red[w * SS_RED + 64];
    }
}

// s_j = sum_c E_jc * y_c + d_j * y_m for the tokens of this thread's quad (4 threads per token).
// f(j, s) is called by the quad's first lane.
template <typename F>
__device__ __forceinline__ void ss_col_pass(const float* E, const float* dst, const float* y, int n, F f) {
    const int tid = threadIdx.x, q = tid & 3, jt = tid >> 2;
    float4 yr[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) yr[i] = *reinterpret_cast<const float4*>(y + (q + 4 * i) * 4);
    const float ym = y[64];
    for (int j0 = 0; j0 < n; j0 += SS_NT / 4) {
        const int j = j0 + jt;
        const bool valid = j < n;
        const int jj = valid ? j : n - 1;
        const float* row = E + jj * SS_PITCH;
        float s = (q == 0) ? dst[jj] * ym : 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float4 e4 = *reinterpret_cast<const float4*>(row + (q + 4 * i) * 4);
            s = fmaf(e4.x, yr[i].x, s); s = fmaf(e4.y, yr[i].y, s);
            s = fmaf(e4.z, yr[i].z, s); s = fmaf(e4.w, yr[i].w, s);
        }
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        if (valid && q == 0) f(j, s);
    }
}

template <typename TP>
__global__ void __launch_bounds__(SS_NT, 2) sinkhorn_fwd_scaled(const float* __restrict__ S, const float* __restrict__ dust_bin,
                                                                TP* __restrict__ P, float* __restrict__ U,
                                                                float* __restrict__ V, int n, int iters, SinkConst2 k) {
    constexpr int NT = SS_NT, NW = SS_NT / 32;
    extern __shared__ __align__(16) float sk_smem[];
    float* E = sk_smem;                          // [n][SS_PITCH]
    float* mxs = E + n * SS_PITCH;       // [n]  per-token shift
    const int np = (n + 3) & ~3;                 // vectors padded so that everything stays 16-byte aligned
    float* dst = mxs + np;                       // [n]  dustbin column of E
    float* binv = dst + np;                      // [n]  2^(mx_j - B0) before the first pass, 1/s_j afterwards
    float* us = binv + np;                       // [SS_RED]
    float* alpha = us + SS_RED;                  // [SS_RED]  2^(u_c - A)
    float* red = alpha + SS_RED;                 // [NW][SS_RED]
    float* wred = red + NW * SS_RED;             // [32]
    float* scal = wred + 32;                     // [8]  0: A   7: guard flag
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    pdl_wait();
    pdl_launch_dependents();
    const float z2 = __ldg(dust_bin) * VPR_L2E;
    const float* Sb = S + (size_t)b * n * 64;
    if (tid == 0) scal[7] = 0.f;
    float bmax = ss_load_tile(Sb, z2, n, E, mxs, dst);
    bmax = warp_max(bmax);
    if (lane == 0) wred[warp] = bmax;
    __syncthreads();
    float Bcur = wred[0];
#pragma unroll
    for (int w = 1; w < NW; ++w) Bcur = fmaxf(Bcur, wred[w]);
    for (int j = tid; j < n; j += NT) binv[j] = fex2(mxs[j] - Bcur);       // beta_j with v = 0
    float bscale = 1.f;
    __syncthreads();

    for (int it = 0; it < iters; ++it) {
        // ---- u_c = log_a_c - LSE_j(M_jc + v_j) = log_a_c - (B + log2 sum_j E_jc beta_j) ----
        ss_row_pass(E, dst, binv, bscale, n, red);
        __syncthreads();
        if (warp == 0) {
            float r0, r1, rm;
            ss_row_totals(red, lane, r0, r1, rm);
            const bool bad = !(ss_in_range(r0) && ss_in_range(r1) && ss_in_range(rm));
            const float u0 = k.norm2 - (flg2(r0) + Bcur);
            const float u1 = k.norm2 - (flg2(r1) + Bcur);
            const float um = k.log_a_m2 - (flg2(rm) + Bcur);
            const float A = warp_max(fmaxf(fmaxf(u0, u1), um));
            us[lane] = u0; us[lane + 32] = u1;
            alpha[lane] = fex2(u0 - A); alpha[lane + 32] = fex2(u1 - A);
            float* Ub = U + ((size_t)b * iters + it) * 65;
            Ub[lane] = u0 * VPR_LN2; Ub[lane + 32] = u1 * VPR_LN2;
            if (lane == 0) { us[64] = um; alpha[64] = fex2(um - A); Ub[64] = um * VPR_LN2; scal[0] = A; }
            if (bad) scal[7] = 1.f;
        }
        __syncthreads();
        // ---- v_j = log_b - LSE_c(M_jc + u_c) = log_b - (mx_j + A + log2 sum_c E_jc alpha_c) ----
        const float A = scal[0];
        float smin = INFINITY;
        bool bad = false;
        float* Vb = V + ((size_t)b * iters + it) * n;
        ss_col_pass(E, dst, alpha, n, [&](int j, float s) {
            bad |= !ss_in_range(s);
            Vb[j] = (k.norm2 - (mxs[j] + flg2(s) + A)) * VPR_LN2;
            binv[j] = __frcp_rn(s);
            smin = fminf(smin, s);
        });
        smin = -warp_max(-smin);
        if (lane == 0) wred[warp] = smin;
        if (bad) scal[7] = 1.f;
        __syncthreads();
        smin = wred[0];
#pragma unroll
        for (int w = 1; w < NW; ++w) smin = fminf(smin, wred[w]);
        bscale = smin;                                   // beta_j = smin / s_j  (max_j beta_j = 1)
        Bcur = k.norm2 - A - flg2(smin);                 // = max_j (v_j + mx_j)
    }
    if (scal[7] != 0.f) {                                // block-uniform: out of fp32 range -> exact path
        __syncthreads();
        sinkhorn_fwd_logdomain<TP, 64>(sk_smem, S, dust_bin, P, U, V, n, iters, k);
        return;
    }
    // ---- P_jc = 2^(M_jc + u_c + v_j - norm) = E_jc * alpha_c / s_j ----
    TP* Pb = P + (size_t)b * n * 64;
    const int cg = tid & 15;
    const float4 a4 = *reinterpret_cast<const float4*>(alpha + cg * 4);
    for (int idx = tid; idx < n * 16; idx += NT) {
        const int j = idx >> 4;
        const float bj = binv[j];
        const float4 e4 = *reinterpret_cast<const float4*>(E + j * SS_PITCH + cg * 4);
        const float p0 = e4.x * a4.x * bj, p1 = e4.y * a4.y * bj, p2 = e4.z * a4.z * bj, p3 = e4.w * a4.w * bj;
        if constexpr (sizeof(TP) == 2) {
            __nv_bfloat162 lo = __floats2bfloat162_rn(p0, p1), hi = __floats2bfloat162_rn(p2, p3);
            uint2 pk;
            pk.x = *reinterpret_cast<uint32_t*>(&lo);
            pk.y = *reinterpret_cast<uint32_t*>(&hi);
            *reinterpret_cast<uint2*>(Pb + idx * 4) = pk;
        } else {
            *reinterpret_cast<float4*>(Pb + idx * 4) = make_float4(p0, p1, p2, p3);
        }
    }
}

// KIT > 0: iteration count known at compile time (coefficients of the final rank-2K update live in
// registers); KIT == 0: run-time count, coefficients re-read from shared memory.
template <typename TD, int KIT>
__global__ void __launch_bounds__(SS_NT, 2) sinkhorn_bwd_scaled(const float* __restrict__ S, const float* __restrict__ dust_bin,
                                                                const float* __restrict__ U, const float* __restrict__ V,
                                                                const float* __restrict__ dP, TD* __restrict__ dS,
                                                                float* __restrict__ d_bias, int n, int iters_rt,
                                                                SinkConst2 k) {
    constexpr int NT = SS_NT, NW = SS_NT / 32;
    const int iters = KIT > 0 ? KIT : iters_rt;
    pdl_wait();
    pdl_launch_dependents();
    extern __shared__ __align__(16) float sk_smem[];
    float* E = sk_smem;                                  // [n][SS_PITCH]
    float* mxs = E + n * SS_PITCH;               // [n]
    const int np = (n + 3) & ~3;
    float* dst = mxs + np;                               // [n]
    float* bv = dst + np;                                // [iters][np]  Wv^k = E * a^k_c * bv^k_j
    float* bu = bv + iters * np;                 // [iters][np]  Wu^k = E * a'^k_c * bu^k_j
    float* X = bu + iters * np;                  // [iters][np]  dv^k_j * bv^k_j
    float* aK = X + iters * np;                  // [iters][SS_RED]  a^k_c ; [64] dustbin of Wv ; [65] dustbin of Wu
    float* Y = aK + iters * SS_RED;              // [iters][SS_RED]  du^k_c * a'^k_c ; [64] dustbin
    float* Ak = Y + iters * SS_RED;              // [16]
    float* red = Ak + 16;                                // [NW][SS_RED]
    float* wred = red + NW * SS_RED;                     // [32]
    float* scal = wred + 32;                             // [8]  7: guard flag
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float z2 = __ldg(dust_bin) * VPR_L2E;
    const float* Sb = S + (size_t)b * n * 64;
    const float* dPb = dP + (size_t)b * n * 64;
    if (tid == 0) scal[7] = 0.f;
    ss_load_tile(Sb, z2, n, E, mxs, dst);
    // per-iteration cluster factors: warp k handles iteration k
    for (int kk = warp; kk < iters; kk += NW) {
        const float* Ub = U + ((size_t)b * iters + kk) * 65;
        const float u0 = __ldg(Ub + lane) * VPR_L2E, u1 = __ldg(Ub + 32 + lane) * VPR_L2E, um = __ldg(Ub + 64) * VPR_L2E;
        const float A = warp_max(fmaxf(fmaxf(u0, u1), um));
        aK[kk * SS_RED + lane] = fex2(u0 - A);
        aK[kk * SS_RED + 32 + lane] = fex2(u1 - A);
        if (lane == 0) {
            aK[kk * SS_RED + 64] = fex2(um - A);
            aK[kk * SS_RED + 65] = fex2(um - A + (k.norm2 - k.log_a_m2));
            Ak[kk] = A;
        }
    }
    __syncthreads();
    {   // per-iteration token factors
        bool bad = false;
        for (int j = tid; j < n; j += NT) {
            const float mj = mxs[j] - k.norm2;
            float vprev = 0.f;
            for (int kk = 0; kk < iters; ++kk) {
                const float vk = __ldg(V + ((size_t)b * iters + kk) * n + j) * VPR_L2E;
                const float fv = fex2(vk + mj + Ak[kk]), fu = fex2(vprev + mj + Ak[kk]);
                bad |= !(fv < SS_HI && fu < SS_HI);
                bv[kk * np + j] = fv;
                bu[kk * np + j] = fu;
                vprev = vk;
            }
        }
        if (bad) scal[7] = 1.f;
    }
    __syncthreads();
    if (scal[7] != 0.f) {
        __syncthreads();
        sinkhorn_bwd_logdomain<TD, 64>(sk_smem, S, dust_bin, U, V, dP, dS, d_bias, n, iters, k);
        return;
    }
    const int cg = tid & 15;
    const float* aLast = aK + (iters - 1) * SS_RED;
    const float* bvLast = bv + (iters - 1) * np;
    // ---- G = dP * P (P = Wv^K):  dv^K_j = sum_c G_jc,  du^K_c = sum_j G_jc ----
    float du0 = 0.f, du1 = 0.f;       // warp 0: du of clusters lane, lane + 32 for the current iteration
    {
        const float4 a4 = *reinterpret_cast<const float4*>(aLast + cg * 4);
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        const int total = n * 16;
        for (int b0 = 0; b0 < total; b0 += SS_LDB * NT) {
            float4 dbuf[SS_LDB];
#pragma unroll
            for (int u = 0; u < SS_LDB; ++u) {
                const int idx = b0 + u * NT + tid;
                const int j = idx < total ? (idx >> 4) : (n - 1);
                dbuf[u] = __ldg(reinterpret_cast<const float4*>(dPb + j * 64) + cg);
            }
#pragma unroll
            for (int u = 0; u < SS_LDB; ++u) {
                const int idx = b0 + u * NT + tid;
                if (b0 + u * NT >= total) break;                     // block-uniform
                const bool valid = idx < total;
                const int j = valid ? (idx >> 4) : (n - 1);
                const float4 d4 = dbuf[u];
                const float4 e4 = *reinterpret_cast<const float4*>(E + j * SS_PITCH + cg * 4);
                const float bj = bvLast[j];
                const float t0 = d4.x * e4.x, t1 = d4.y * e4.y, t2 = d4.z * e4.z, t3 = d4.w * e4.w;
                float rs = t0 * a4.x + t1 * a4.y + t2 * a4.z + t3 * a4.w;
#pragma unroll
                for (int o = 1; o < 16; o <<= 1) rs += __shfl_xor_sync(0xffffffffu, rs, o);
                if (valid) {
                    acc.x = fmaf(t0, bj, acc.x); acc.y = fmaf(t1, bj, acc.y);
                    acc.z = fmaf(t2, bj, acc.z); acc.w = fmaf(t3, bj, acc.w);
                    if (cg == 0) X[(iters - 1) * np + j] = rs * bj * bj;      // dv^K_j * bv^K_j
                }
            }
        }
        acc.x += __shfl_xor_sync(0xffffffffu, acc.x, 16); acc.y += __shfl_xor_sync(0xffffffffu, acc.y, 16);
        acc.z += __shfl_xor_sync(0xffffffffu, acc.z, 16); acc.w += __shfl_xor_sync(0xffffffffu, acc.w, 16);
        if (lane < 16) *reinterpret_cast<float4*>(red + warp * SS_RED + cg * 4) = acc;
        if (lane == 0) red[warp * SS_RED + 64] = 0.f;
        __syncthreads();
        if (warp == 0) {
            float t0, t1, tm;
            ss_row_totals(red, lane, t0, t1, tm);
            du0 = aLast[lane] * t0;
            du1 = aLast[lane + 32] * t1;
        }
        __syncthreads();
    }
    for (int kk = iters - 1; kk >= 0; --kk) {
        const float* ak = aK + kk * SS_RED;
        // du^k_c -= sum_j dv^k_j Wv^k_jc = a^k_c * sum_j E_jc x^k_j
        ss_row_pass(E, dst, X + kk * np, 1.f, n, red);
        __syncthreads();
        if (warp == 0) {
            float t0, t1, tm;
            ss_row_totals(red, lane, t0, t1, tm);
            du0 -= ak[lane] * t0;
            du1 -= ak[lane + 32] * t1;
            const float dum = -ak[64] * tm;
            Y[kk * SS_RED + lane] = du0 * ak[lane];
            Y[kk * SS_RED + 32 + lane] = du1 * ak[lane + 32];
            if (lane == 0) Y[kk * SS_RED + 64] = dum * ak[65];
            du0 = 0.f; du1 = 0.f;
        }
        __syncthreads();
        // dv^{k-1}_j = - sum_c du^k_c Wu^k_jc = - bu^k_j * sum_c E_jc y^k_c
        if (kk > 0) {
            const float* buk = bu + kk * np;
            const float* bvp = bv + (kk - 1) * np;
            float* Xp = X + (kk - 1) * np;
            ss_col_pass(E, dst, Y + kk * SS_RED, n, [&](int j, float s) { Xp[j] = -buk[j] * s * bvp[j]; });
            __syncthreads();
        }
    }
    // ---- dM_jc = G_jc - sum_k (dv^k_j Wv^k_jc + du^k_c Wu^k_jc)
    //            = E_jc * (dP_jc a^K_c bv^K_j - sum_k (a^k_c x^k_j + y^k_c bu^k_j)) ----
    TD* dSb = dS + (size_t)b * n * 64;
    float4 bacc = make_float4(0.f, 0.f, 0.f, 0.f);
    {
        constexpr int KR = KIT > 0 ? KIT : 1;
        float4 ar[KR], yr[KR];
        if (KIT > 0) {
#pragma unroll
            for (int kk = 0; kk < KR; ++kk) {
                ar[kk] = *reinterpret_cast<const float4*>(aK + kk * SS_RED + cg * 4);
                yr[kk] = *reinterpret_cast<const float4*>(Y + kk * SS_RED + cg * 4);
            }
        }
        const float4 a4 = *reinterpret_cast<const float4*>(aLast + cg * 4);
        const int total = n * 16;
        for (int b0 = 0; b0 < total; b0 += SS_LDB * NT) {
          float4 dbuf[SS_LDB];                                   // dP again (L2-resident by now), batched
#pragma unroll
          for (int u = 0; u < SS_LDB; ++u) {
              const int idx = b0 + u * NT + tid;
              dbuf[u] = __ldg(reinterpret_cast<const float4*>(dPb + (idx < total ? (idx >> 4) : (n - 1)) * 64) + cg);
          }
#pragma unroll
          for (int u = 0; u < SS_LDB; ++u) {
            const int idx = b0 + u * NT + tid;
            if (idx >= total) break;
            const int j = idx >> 4;
            const float4 d4 = dbuf[u];
            const float4 e4 = *reinterpret_cast<const float4*>(E + j * SS_PITCH + cg * 4);
            const float bj = bvLast[j];
            float c0 = 0.f, c1 = 0.f, c2 = 0.f, c3 = 0.f;
            if (KIT > 0) {
#pragma unroll
                for (int kk = 0; kk < KR; ++kk) {
                    const float xj = X[kk * np + j], uj = bu[kk * np + j];
                    c0 = fmaf(ar[kk].x, xj, fmaf(yr[kk].x, uj, c0)); c1 = fmaf(ar[kk].y, xj, fmaf(yr[kk].y, uj, c1));
                    c2 = fmaf(ar[kk].z, xj, fmaf(yr[kk].z, uj, c2)); c3 = fmaf(ar[kk].w, xj, fmaf(yr[kk].w, uj, c3));
                }
            } else {
                for (int kk = 0; kk < iters; ++kk) {
                    const float4 aa = *reinterpret_cast<const float4*>(aK + kk * SS_RED + cg * 4);
                    const float4 yy = *reinterpret_cast<const float4*>(Y + kk * SS_RED + cg * 4);
                    const float xj = X[kk * np + j], uj = bu[kk * np + j];
                    c0 = fmaf(aa.x, xj, fmaf(yy.x, uj, c0)); c1 = fmaf(aa.y, xj, fmaf(yy.y, uj, c1));
                    c2 = fmaf(aa.z, xj, fmaf(yy.z, uj, c2)); c3 = fmaf(aa.w, xj, fmaf(yy.w, uj, c3));
                }
            }
            const float g0 = e4.x * (d4.x * a4.x * bj - c0), g1 = e4.y * (d4.y * a4.y * bj - c1);
            const float g2 = e4.z * (d4.z * a4.z * bj - c2), g3 = e4.w * (d4.w * a4.w * bj - c3);
            if constexpr (sizeof(TD) == 2) {
                __nv_bfloat162 lo = __floats2bfloat162_rn(g0, g1), hi = __floats2bfloat162_rn(g2, g3);
                uint2 pk;
                pk.x = *reinterpret_cast<uint32_t*>(&lo);
                pk.y = *reinterpret_cast<uint32_t*>(&hi);
                *reinterpret_cast<uint2*>(dSb + idx * 4) = pk;
            } else {
                *reinterpret_cast<float4*>(dSb + idx * 4) = make_float4(g0, g1, g2, g3);
            }
            bacc.x += g0; bacc.y += g1; bacc.z += g2; bacc.w += g3;
          }
        }
    }
    if (d_bias != nullptr) {
        bacc.x += __shfl_xor_sync(0xffffffffu, bacc.x, 16); bacc.y += __shfl_xor_sync(0xffffffffu, bacc.y, 16);
        bacc.z += __shfl_xor_sync(0xffffffffu, bacc.z, 16); bacc.w += __shfl_xor_sync(0xffffffffu, bacc.w, 16);
        if (lane < 16) *reinterpret_cast<float4*>(red + warp * SS_RED + cg * 4) = bacc;
        __syncthreads();
        if (tid < 64) {
            float tot = 0.f;
#pragma unroll
            for (int w = 0; w < NW; ++w) tot += red[w * SS_RED + tid];
            atomicAdd(d_bias + tid, tot);
        }
    }
}

}  // namespace vpr
