// Element-wise log-domain Sinkhorn forward/backward for a compile-time cluster count (m = 64
// upstream): one exp2 per tile element per pass, nothing exponentiated outside a max-subtracted
// logsumexp, so it is safe for any score range.  These bodies are the guard path of the
// scaled-tile kernels in sinkhorn_scaled.cuh (and the only path when their larger tile pitch
// does not fit in shared memory).  Same algorithm as the generic kernels in salad.cu
// (SURVEY.md A-2 / A-3), restructured for issue-slot efficiency:
//   * 512 threads per image, everything in the base-2 log domain so that every exp / log is a
//     single MUFU (ex2.approx / lg2.approx) with fp32 range (max-subtracted logsumexp kept);
//   * row reductions (over tokens j): thread = (cluster c, 1/8 of the tokens), partials via smem;
//   * column reductions (over clusters c): two threads per token, 32 clusters each held in
//     registers together with the 32 potentials they need; the pair's halves are rotated by 16
//     banks so the [n][m+1] tile is read without bank conflicts;
//   * the score tile is the only O(n*m) shared-memory object; dP/dS stream through registers with
//     fully coalesced global accesses.
#pragma once
#include "common.cuh"

namespace vpr {

#define VPR_L2E 1.4426950408889634f
#define VPR_LN2 0.6931471805599453f

__device__ __forceinline__ float fex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float flg2(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

template <int NT>
__device__ __forceinline__ float block_sum_nt(float v, float* red) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float r = 0.f;
#pragma unroll
    for (int i = 0; i < NT / 32; ++i) r += red[i];
    return r;
}
template <int NT>
__device__ __forceinline__ float block_max_nt(float v, float* red) {
    v = warp_max(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float r = red[0];
#pragma unroll
    for (int i = 1; i < NT / 32; ++i) r = fmaxf(r, red[i]);
    return r;
}

struct SinkConst2 {     // base-2 log domain constants
    float norm2;        // -log2(n+m)
    float log_a_m2;     // norm2 + log2(n-m)
};

static constexpr int SK_NT = 512;

template <int M_>
__host__ __device__ constexpr size_t sinkfast_fwd_floats(int n) {
    return (size_t)n * (M_ + 1) + n + (M_ + 1) + SK_NT + 32;
}
template <int M_>
__host__ __device__ constexpr size_t sinkfast_bwd_floats(int n, int iters) {
    return (size_t)n * (M_ + 1) + (size_t)(iters + 1) * n + (size_t)iters * (M_ + 1) + (size_t)iters * n +
           (size_t)iters * (M_ + 1) + SK_NT + 32;
}

// NT: threads of the calling CTA (512 for the stand-alone kernels below and the shared-tile kernels; the register-tile
// kernels also call these bodies from 256-thread CTAs)
template <typename TP, int M_, int NT = 512>
__device__ __noinline__ void sinkhorn_fwd_logdomain(float* sm, const float* __restrict__ S,
                                                    const float* __restrict__ dust_bin, TP* __restrict__ P,
                                                    float* __restrict__ U, float* __restrict__ V, int n, int iters,
                                                    SinkConst2 k) {
    static_assert(M_ == 64, "column-pass register tiling assumes 64 clusters");
    constexpr int PITCH = M_ + 1, PARTS = NT / M_;
    static_assert(NT % M_ == 0 && NT <= 512, "thread count: a multiple of the cluster count, at most the 512 the scratch is sized for");
    float* Ms = sm;
    float* vs = Ms + (size_t)n * PITCH;
    float* us = vs + n;
    float* red = us + (M_ + 1);
    float* red2 = red + NT;
    const int b = blockIdx.x, tid = threadIdx.x;
    const float z2 = __ldg(dust_bin) * VPR_L2E;
    const float* Sb = S + (size_t)b * n * M_;
    for (int e = tid; e < n * M_; e += NT) Ms[(e / M_) * PITCH + (e % M_)] = __ldg(Sb + e) * VPR_L2E;
    for (int j = tid; j < n; j += NT) vs[j] = 0.f;
    __syncthreads();
    const int c = tid % M_, part = tid / M_;
    const int sub = tid & 1, jt = tid >> 1;
    int cidx[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) cidx[i] = sub * 32 + ((i + sub * 16) & 31);
    float lse_v2 = flg2((float)n);
    for (int it = 0; it < iters; ++it) {
        // ---- u_c = log_a - LSE_j(M_jc + v_j) ----
        float mx = -INFINITY;
        for (int j = part; j < n; j += PARTS) mx = fmaxf(mx, Ms[j * PITCH + c] + vs[j]);
        red[tid] = mx;
        __syncthreads();
#pragma unroll
        for (int q = 0; q < PARTS; ++q) mx = fmaxf(mx, red[q * M_ + c]);
        __syncthreads();
        float sum = 0.f;
        for (int j = part; j < n; j += PARTS) sum += fex2(Ms[j * PITCH + c] + vs[j] - mx);
        red[tid] = sum;
        __syncthreads();
        if (part == 0) {
            float tot = 0.f;
#pragma unroll
            for (int q = 0; q < PARTS; ++q) tot += red[q * M_ + c];
            us[c] = k.norm2 - (flg2(tot) + mx);
        }
        if (tid == 0) us[M_] = k.log_a_m2 - (z2 + lse_v2);
        __syncthreads();
        // ---- v_j = log_b - LSE_{c<=m}(M_jc + u_c) ----
        float ureg[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) ureg[i] = us[cidx[i]];
        const float um = z2 + us[M_];
        float vmax = -INFINITY;
        for (int j0 = 0; j0 < n; j0 += NT / 2) {
            const int j = j0 + jt;
            const bool valid = j < n;
            const float* row = Ms + (valid ? j : n - 1) * PITCH;
            float x[32];
            float cm = sub == 0 ? um : -INFINITY;
#pragma unroll
            for (int i = 0; i < 32; ++i) { x[i] = row[cidx[i]] + ureg[i]; cm = fmaxf(cm, x[i]); }
            cm = fmaxf(cm, __shfl_xor_sync(0xffffffffu, cm, 1));
            float cs = sub == 0 ? fex2(um - cm) : 0.f;
#pragma unroll
            for (int i = 0; i < 32; ++i) cs += fex2(x[i] - cm);
            cs += __shfl_xor_sync(0xffffffffu, cs, 1);
            const float vj = k.norm2 - (flg2(cs) + cm);
            if (valid) {
                if (sub == 0) vs[j] = vj;
                vmax = fmaxf(vmax, vj);
            }
        }
        vmax = block_max_nt<NT>(vmax, red2);       // (contains the barriers that publish vs)
        float vsum = 0.f;
        for (int j = tid; j < n; j += NT) vsum += fex2(vs[j] - vmax);
        vsum = block_sum_nt<NT>(vsum, red2);
        lse_v2 = flg2(vsum) + vmax;
        if (tid <= M_) U[((size_t)b * iters + it) * (M_ + 1) + tid] = us[tid] * VPR_LN2;
        for (int j = tid; j < n; j += NT) V[((size_t)b * iters + it) * n + j] = vs[j] * VPR_LN2;
        __syncthreads();
    }
    TP* Pb = P + (size_t)b * n * M_;
    for (int e = tid; e < n * M_; e += NT) {
        const int j = e / M_, cc = e % M_;
        const float p = fex2(Ms[j * PITCH + cc] + us[cc] + vs[j] - k.norm2);
        if constexpr (sizeof(TP) == 2) Pb[e] = __float2bfloat16_rn(p); else Pb[e] = p;
    }
}

template <typename TD, int M_, int NT = 512>
__device__ __noinline__ void sinkhorn_bwd_logdomain(float* sm, const float* __restrict__ S,
                                                    const float* __restrict__ dust_bin, const float* __restrict__ U,
                                                    const float* __restrict__ V, const float* __restrict__ dP,
                                                    TD* __restrict__ dS, float* __restrict__ d_bias, int n, int iters,
                                                    SinkConst2 k) {
    static_assert(M_ == 64, "column-pass register tiling assumes 64 clusters");
    constexpr int PITCH = M_ + 1, PARTS = NT / M_;
    static_assert(NT % M_ == 0 && NT <= 512, "thread count: a multiple of the cluster count, at most the 512 the scratch is sized for");
    float* Ms = sm;                                   // [n][PITCH]        scores * log2(e)
    float* vs = Ms + (size_t)n * PITCH;               // [iters+1][n]      v^0..v^K (log2 units)
    float* us = vs + (size_t)(iters + 1) * n;         // [iters][M_+1]     u^1..u^K (log2 units)
    float* DV = us + (size_t)iters * (M_ + 1);        // [iters][n]
    float* DU = DV + (size_t)iters * n;               // [iters][M_+1]
    float* red = DU + (size_t)iters * (M_ + 1);       // [NT]
    float* red2 = red + NT;                           // [32]
    const int b = blockIdx.x, tid = threadIdx.x;
    const float z2 = __ldg(dust_bin) * VPR_L2E;
    const float* Sb = S + (size_t)b * n * M_;
    const float* dPb = dP + (size_t)b * n * M_;
    for (int e = tid; e < n * M_; e += NT) Ms[(e / M_) * PITCH + (e % M_)] = __ldg(Sb + e) * VPR_L2E;
    for (int j = tid; j < n; j += NT) vs[j] = 0.f;
    for (int e = tid; e < iters * n; e += NT) { vs[n + e] = __ldg(V + (size_t)b * iters * n + e) * VPR_L2E; DV[e] = 0.f; }
    for (int e = tid; e < iters * (M_ + 1); e += NT) { us[e] = __ldg(U + (size_t)b * iters * (M_ + 1) + e) * VPR_L2E; DU[e] = 0.f; }
    __syncthreads();
    const int c = tid % M_, part = tid / M_;
    const int sub = tid & 1, jt = tid >> 1;
    const int lane = tid & 31;
    int cidx[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) cidx[i] = sub * 32 + ((i + sub * 16) & 31);
    const float* uK = us + (size_t)(iters - 1) * (M_ + 1);
    const float* vK = vs + (size_t)iters * n;
    float* duK = DU + (size_t)(iters - 1) * (M_ + 1);
    float* dvK = DV + (size_t)(iters - 1) * n;
    // G_jc = dP_jc * P_jc ; du_c = sum_j G, dv_j = sum_c G.  A warp holds 32 consecutive clusters of one
    // token per step, so the token sums are a warp reduction + one shared atomic per half row.
    {
        const float uc = uK[c] - k.norm2;
        float acc = 0.f;
#pragma unroll 4
        for (int j = part; j < n; j += PARTS) {
            const float g = __ldg(dPb + j * M_ + c) * fex2(Ms[j * PITCH + c] + uc + vK[j]);
            acc += g;
            const float rs = warp_sum(g);
            if (lane == 0) atomicAdd(dvK + j, rs);
        }
        red[tid] = acc;
        __syncthreads();
        if (part == 0) {
            float tot = 0.f;
#pragma unroll
            for (int q = 0; q < PARTS; ++q) tot += red[q * M_ + c];
            duK[c] = tot;
        }
        __syncthreads();
    }
    for (int it = iters - 1; it >= 0; --it) {
        const float* uk = us + (size_t)it * (M_ + 1);
        const float* vk = vs + (size_t)(it + 1) * n;
        const float* vkm1 = vs + (size_t)it * n;
        float* du = DU + (size_t)it * (M_ + 1);
        const float* dv = DV + (size_t)it * n;
        // du_c -= sum_j dv_j * Wv_jc,  Wv_jc = 2^(M_jc + u_c + v_j - log_b)
        {
            const float uc = uk[c] - k.norm2;
            float acc = 0.f;
            for (int j = part; j < n; j += PARTS) acc += dv[j] * fex2(Ms[j * PITCH + c] + uc + vk[j]);
            red[tid] = acc;
            // dustbin row: Wv_jm = 2^(z + u_m + v_j - log_b)
            const float um = z2 + uk[M_] - k.norm2;
            float a2 = 0.f;
            for (int j = tid; j < n; j += NT) a2 += dv[j] * fex2(um + vk[j]);
            a2 = block_sum_nt<NT>(a2, red2);          // barriers also publish red[]
            if (part == 0) {
                float tot = 0.f;
#pragma unroll
                for (int q = 0; q < PARTS; ++q) tot += red[q * M_ + c];
                du[c] -= tot;
            }
            if (tid == 0) du[M_] -= a2;
            __syncthreads();
        }
        // dv^{k-1}_j = - sum_{c<=m} du_c * Wu_jc,  Wu_jc = 2^(M_jc + v^{k-1}_j + u_c - log_a_c)
        if (it > 0) {
            float* dvn = DV + (size_t)(it - 1) * n;
            float ua[32], dur[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) { ua[i] = uk[cidx[i]] - k.norm2; dur[i] = du[cidx[i]]; }
            const float uam = z2 + uk[M_] - k.log_a_m2;
            const float dum = du[M_];
            for (int j0 = 0; j0 < n; j0 += NT / 2) {
                const int j = j0 + jt;
                const bool valid = j < n;
                const int jj = valid ? j : n - 1;
                const float* row = Ms + jj * PITCH;
                const float vj = vkm1[jj];
                float a3 = sub == 0 ? dum * fex2(uam + vj) : 0.f;
#pragma unroll
                for (int i = 0; i < 32; ++i) a3 += dur[i] * fex2(row[cidx[i]] + ua[i] + vj);
                a3 += __shfl_xor_sync(0xffffffffu, a3, 1);
                if (valid && sub == 0) dvn[j] = -a3;
            }
            __syncthreads();
        }
    }
    // dM_jc = G_jc - sum_k ( DV[k][j] * Wv^k_jc + DU[k][c] * Wu^k_jc )
    TD* dSb = dS + (size_t)b * n * M_;
    float bsum = 0.f;
#pragma unroll 2
    for (int e = tid; e < n * M_; e += NT) {      // e % M_ == c because NT % M_ == 0
        const int j = e / M_;
        const float mv = Ms[j * PITCH + c] - k.norm2;
        float g = __ldg(dPb + e) * fex2(mv + uK[c] + vK[j]);
        for (int it = 0; it < iters; ++it) {
            const float t = mv + us[(size_t)it * (M_ + 1) + c];
            g -= DV[(size_t)it * n + j] * fex2(t + vs[(size_t)(it + 1) * n + j]);
            g -= DU[(size_t)it * (M_ + 1) + c] * fex2(t + vs[(size_t)it * n + j]);
        }
        if constexpr (sizeof(TD) == 2) dSb[e] = __float2bfloat16_rn(g); else dSb[e] = g;
        bsum += g;
    }
    if (d_bias != nullptr) {
        __syncthreads();
        red[tid] = bsum;
        __syncthreads();
        if (part == 0) {
            float tot = 0.f;
#pragma unroll
            for (int q = 0; q < PARTS; ++q) tot += red[q * M_ + c];
            atomicAdd(d_bias + c, tot);
        }
    }
}

// stand-alone launches of the log-domain bodies (tiles too large for the scaled kernels' pitch)
template <typename TP, int M_>
__global__ void __launch_bounds__(SK_NT) sinkhorn_fwd_fast(const float* __restrict__ S, const float* __restrict__ dust_bin,
                                                           TP* __restrict__ P, float* __restrict__ U,
                                                           float* __restrict__ V, int n, int iters, SinkConst2 k) {
    extern __shared__ __align__(16) float sk_smem[];
    sinkhorn_fwd_logdomain<TP, M_>(sk_smem, S, dust_bin, P, U, V, n, iters, k);
}
template <typename TD, int M_>
__global__ void __launch_bounds__(SK_NT) sinkhorn_bwd_fast(const float* __restrict__ S, const float* __restrict__ dust_bin,
                                                           const float* __restrict__ U, const float* __restrict__ V,
                                                           const float* __restrict__ dP, TD* __restrict__ dS,
                                                           float* __restrict__ d_bias, int n, int iters, SinkConst2 k) {
    extern __shared__ __align__(16) float sk_smem[];
    sinkhorn_bwd_logdomain<TD, M_>(sk_smem, S, dust_bin, U, V, dP, dS, d_bias, n, iters, k);
}

}  // namespace vpr
