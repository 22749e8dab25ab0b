// SALAD aggregation head, forward and backward.
//
// Data layout in HBM (T = B*n patch tokens, everything token-major so that every GEMM has
// the tokens on the 128-lane M side of the tcgen05 tile):
//   X   [T, d]        input tokens (the memory behind upstream's permuted [B,d,h,w] view)
//   H   [T, 2*hidden] relu(dropout(X W0^T + b0)); columns [0,hidden) cluster MLP, [hidden,2h) score MLP
//   F   [T, l]        cluster features          S [T, m] fp32 scores
//   P   [T, m]        Sinkhorn assignment (dustbin row never materialised)
//   AGG [B, l, m]     fp32  sum_j F[j,l] P[j,c]
//   out [B, g + l*m]  normalize(cat(normalize(TF), normalize(AGG, over l).flatten))
#include "common.cuh"
#include "gemm.h"
#include "sinkhorn_rt.cuh"
#include "finalize_fast.cuh"
#include "../../include/vpr_b200.h"

#include <math.h>

namespace vpr {

// ---------------------------------------------------------------------------
// scalar load/store helpers over the two storage types
// ---------------------------------------------------------------------------
__device__ __forceinline__ float ldf(const float* p) { return *p; }
__device__ __forceinline__ float ldf(const bf16* p) { return __bfloat162float(*p); }
__device__ __forceinline__ void stf(float* p, float v) { *p = v; }
__device__ __forceinline__ void stf(bf16* p, float v) { *p = __float2bfloat16_rn(v); }

// block-wide reductions for 256-thread blocks (8 warps); red must hold >= 8 floats
__device__ __forceinline__ float block_sum256(float v, float* red) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float r = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) r += red[i];
    return r;
}
__device__ __forceinline__ float block_max256(float v, float* red) {
    v = warp_max(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float r = red[0];
#pragma unroll
    for (int i = 1; i < 8; ++i) r = fmaxf(r, red[i]);
    return r;
}

// ---------------------------------------------------------------------------
// parameter packing: fp32 master weights -> run dtype, first layers concatenated
// ---------------------------------------------------------------------------
struct PackSeg { const float* src; void* dst; long long n4; };      // n4 = number of 4-element groups
struct PackPlan { PackSeg seg[8]; long long total4; };

// fp32 master weights -> run dtype, four elements per thread (all tensor sizes are multiples of 8)
template <typename T>
__global__ void __launch_bounds__(256) pack_params_kernel(const PackPlan plan) {
    pdl_wait();
    pdl_launch_dependents();
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < plan.total4;
         i += (long long)gridDim.x * blockDim.x) {
        long long e = i;
        int sidx = 0;
#pragma unroll
        for (int k = 0; k < 7; ++k)
            if (sidx == k && e >= plan.seg[k].n4) { e -= plan.seg[k].n4; sidx = k + 1; }
        const float4 v = __ldg(reinterpret_cast<const float4*>(plan.seg[sidx].src) + e);
        if (sizeof(T) == 2 && sidx < 6) {
            __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
            uint2 pk;
            pk.x = *reinterpret_cast<uint32_t*>(&lo);
            pk.y = *reinterpret_cast<uint32_t*>(&hi);
            reinterpret_cast<uint2*>(plan.seg[sidx].dst)[e] = pk;
        } else {
            reinterpret_cast<float4*>(plan.seg[sidx].dst)[e] = v;      // fp32 mode, and the bias concat
        }
    }
}

// ---------------------------------------------------------------------------
// Sinkhorn forward.  One CTA (256 threads) per image; the [n, m] score tile stays in shared
// memory (pitch m+1 -> conflict-free in both reduction directions) for all iterations.
// The dustbin row is a constant z, so it is never stored: it enters each column
// logsumexp as one scalar per image.
// ---------------------------------------------------------------------------
struct SinkConst {
    float norm;      // -log(n+m)  (fp32 like upstream's torch.tensor(math.log(..)))
    float log_a_m;   // norm + log(n-m)
};

template <typename TP>
__global__ void __launch_bounds__(256) sinkhorn_fwd_kernel(const float* __restrict__ S, const float* __restrict__ dust_bin,
                                                           TP* __restrict__ P, float* __restrict__ U,
                                                           float* __restrict__ V, int n, int m, int iters,
                                                           SinkConst k) {
    extern __shared__ float sm[];
    const int pitch = m + 1;
    float* Ms = sm;                       // [n][pitch]
    float* vs = Ms + (size_t)n * pitch;   // [n]
    float* us = vs + n;                   // [m+1]
    float* red = us + (m + 1);            // [256] partials
    float* red2 = red + 256;              // [8]
    const int b = blockIdx.x;
    const int tid = threadIdx.x;
    const float z = __ldg(dust_bin);
    const float* Sb = S + (size_t)b * n * m;
    for (int e = tid; e < n * m; e += 256) {
        const int j = e / m, c = e - j * m;
        Ms[j * pitch + c] = __ldg(Sb + e);
    }
    for (int j = tid; j < n; j += 256) vs[j] = 0.f;
    __syncthreads();
    const int parts = 256 / m;            // m in {16,32,64,128}
    const int c = tid % m, part = tid / m;
    float lse_v = logf((float)n);         // logsumexp_j(v_j) with v = 0
    for (int it = 0; it < iters; ++it) {
        // ---- u_c = log_a - LSE_j(M_jc + v_j) ----
        float mx = -INFINITY;
        for (int j = part; j < n; j += parts) mx = fmaxf(mx, Ms[j * pitch + c] + vs[j]);
        red[tid] = mx;
        __syncthreads();
        mx = red[c];
        for (int q = 1; q < parts; ++q) mx = fmaxf(mx, red[q * m + c]);
        __syncthreads();
        float sum = 0.f;
        for (int j = part; j < n; j += parts) sum += expf(Ms[j * pitch + c] + vs[j] - mx);
        red[tid] = sum;
        __syncthreads();
        if (part == 0) {
            float tot = red[c];
            for (int q = 1; q < parts; ++q) tot += red[q * m + c];
            us[c] = k.norm - (logf(tot) + mx);
        }
        if (tid == 0) us[m] = k.log_a_m - (z + lse_v);
        __syncthreads();
        // ---- v_j = log_b - LSE_{c<=m}(M_jc + u_c) ----
        float vmax = -INFINITY;
        for (int j = tid; j < n; j += 256) {
            const float* row = Ms + j * pitch;
            float cmx = z + us[m];
            for (int cc = 0; cc < m; ++cc) cmx = fmaxf(cmx, row[cc] + us[cc]);
            float cs = expf(z + us[m] - cmx);
            for (int cc = 0; cc < m; ++cc) cs += expf(row[cc] + us[cc] - cmx);
            const float vj = k.norm - (logf(cs) + cmx);
            vs[j] = vj;
            vmax = fmaxf(vmax, vj);
        }
        vmax = block_max256(vmax, red2);
        float vsum = 0.f;
        for (int j = tid; j < n; j += 256) vsum += expf(vs[j] - vmax);
        vsum = block_sum256(vsum, red2);
        lse_v = logf(vsum) + vmax;
        // save potentials for the backward
        if (tid <= m) U[((size_t)b * iters + it) * (m + 1) + tid] = us[tid];
        for (int j = tid; j < n; j += 256) V[((size_t)b * iters + it) * n + j] = vs[j];
        __syncthreads();
    }
    TP* Pb = P + (size_t)b * n * m;
    for (int e = tid; e < n * m; e += 256) {
        const int j = e / m, cc = e - j * m;
        stf(Pb + e, expf(Ms[j * pitch + cc] + us[cc] + vs[j] - k.norm));
    }
}

// ---------------------------------------------------------------------------
// Sinkhorn backward: reverse-unrolled recurrence on the potentials'
// adjoints (du, dv); the per-element dM is assembled once at the end from the saved
// per-iteration adjoints, so only the score tile itself has to stay in shared memory.
// ---------------------------------------------------------------------------
template <typename TD>
__global__ void __launch_bounds__(256) sinkhorn_bwd_kernel(const float* __restrict__ S, const float* __restrict__ dust_bin,
                                                           const float* __restrict__ U, const float* __restrict__ V,
                                                           const float* __restrict__ dP, TD* __restrict__ dS,
                                                           float* __restrict__ d_bias, int n, int m, int iters,
                                                           SinkConst k) {
    extern __shared__ float sm[];
    const int pitch = m + 1;
    float* Ms = sm;                                // [n][pitch]
    float* vs = Ms + (size_t)n * pitch;            // [iters+1][n]   v^0 .. v^K
    float* us = vs + (size_t)(iters + 1) * n;      // [iters][m+1]   u^1 .. u^K
    float* DV = us + (size_t)iters * (m + 1);      // [iters][n]     dv used in the v-step of iteration k
    float* DU = DV + (size_t)iters * n;            // [iters][m+1]   du used in the u-step of iteration k
    float* red = DU + (size_t)iters * (m + 1);     // [256]
    const int b = blockIdx.x;
    const int tid = threadIdx.x;
    const float z = __ldg(dust_bin);
    const float* Sb = S + (size_t)b * n * m;
    const float* dPb = dP + (size_t)b * n * m;
    for (int e = tid; e < n * m; e += 256) {
        const int j = e / m, c = e - j * m;
        Ms[j * pitch + c] = __ldg(Sb + e);
    }
    for (int j = tid; j < n; j += 256) vs[j] = 0.f;
    for (int e = tid; e < iters * n; e += 256) vs[n + e] = __ldg(V + (size_t)b * iters * n + e);
    for (int e = tid; e < iters * (m + 1); e += 256) us[e] = __ldg(U + (size_t)b * iters * (m + 1) + e);
    __syncthreads();
    const int parts = 256 / m;
    const int c = tid % m, part = tid / m;
    const float* uK = us + (size_t)(iters - 1) * (m + 1);
    const float* vK = vs + (size_t)iters * n;
    // G_jc = dP_jc * P_jc, P_jc = exp(M_jc + uK_c + vK_j - norm);  du_c = sum_j G, dv_j = sum_c G
    {
        float acc = 0.f;
        for (int j = part; j < n; j += parts)
            acc += __ldg(dPb + j * m + c) * expf(Ms[j * pitch + c] + uK[c] + vK[j] - k.norm);
        red[tid] = acc;
        __syncthreads();
        if (part == 0) {
            float tot = red[c];
            for (int q = 1; q < parts; ++q) tot += red[q * m + c];
            DU[(size_t)(iters - 1) * (m + 1) + c] = tot;
        }
        if (tid == 0) DU[(size_t)(iters - 1) * (m + 1) + m] = 0.f;
        for (int j = tid; j < n; j += 256) {
            const float* row = Ms + j * pitch;
            float acc2 = 0.f;
            for (int cc = 0; cc < m; ++cc)
                acc2 += __ldg(dPb + j * m + cc) * expf(row[cc] + uK[cc] + vK[j] - k.norm);
            DV[(size_t)(iters - 1) * n + j] = acc2;
        }
        __syncthreads();
    }
    for (int it = iters - 1; it >= 0; --it) {
        const float* uk = us + (size_t)it * (m + 1);
        const float* vk = vs + (size_t)(it + 1) * n;     // v^k
        const float* vkm1 = vs + (size_t)it * n;         // v^{k-1}
        float* du = DU + (size_t)it * (m + 1);
        const float* dv = DV + (size_t)it * n;
        // du_c -= sum_j dv_j * Wv_jc,  Wv_jc = exp(M_jc + u_c + v_j - log_b)
        float acc = 0.f;
        for (int j = part; j < n; j += parts) acc += dv[j] * expf(Ms[j * pitch + c] + uk[c] + vk[j] - k.norm);
        red[tid] = acc;
        __syncthreads();
        if (part == 0) {
            float tot = red[c];
            for (int q = 1; q < parts; ++q) tot += red[q * m + c];
            du[c] -= tot;
        }
        __syncthreads();
        {   // dustbin row: Wv_jm = exp(z + u_m + v_j - log_b)
            float a2 = 0.f;
            for (int j = tid; j < n; j += 256) a2 += dv[j] * expf(z + uk[m] + vk[j] - k.norm);
            a2 = block_sum256(a2, red);
            if (tid == 0) du[m] -= a2;
            __syncthreads();
        }
        // dv^{k-1}_j = - sum_{c<=m} du_c * Wu_jc,  Wu_jc = exp(M_jc + v^{k-1}_j + u_c - log_a_c)
        if (it > 0) {
            float* dvn = DV + (size_t)(it - 1) * n;
            float* dun = DU + (size_t)(it - 1) * (m + 1);
            for (int j = tid; j < n; j += 256) {
                const float* row = Ms + j * pitch;
                float a3 = du[m] * expf(z + vkm1[j] + uk[m] - k.log_a_m);
                for (int cc = 0; cc < m; ++cc) a3 += du[cc] * expf(row[cc] + vkm1[j] + uk[cc] - k.norm);
                dvn[j] = -a3;
            }
            if (tid <= m) dun[tid] = 0.f;
            __syncthreads();
        }
    }
    // dM_jc = G_jc - sum_k ( DV[k][j] * Wv^k_jc + DU[k][c] * Wu^k_jc )
    TD* dSb = dS + (size_t)b * n * m;
    float bsum = 0.f;
    for (int e = tid; e < n * m; e += 256) {
        const int j = e / m, cc = e - j * m;   // cc == c because 256 % m == 0
        const float mv = Ms[j * pitch + cc];
        float g = __ldg(dPb + e) * expf(mv + uK[cc] + vK[j] - k.norm);
        for (int it = 0; it < iters; ++it) {
            const float u = us[(size_t)it * (m + 1) + cc];
            g -= DV[(size_t)it * n + j] * expf(mv + u + vs[(size_t)(it + 1) * n + j] - k.norm);
            g -= DU[(size_t)it * (m + 1) + cc] * expf(mv + u + vs[(size_t)it * n + j] - k.norm);
        }
        stf(dSb + e, g);
        bsum += g;
    }
    if (d_bias != nullptr) {
        __syncthreads();
        red[tid] = bsum;
        __syncthreads();
        if (part == 0) {
            float tot = red[c];
            for (int q = 1; q < parts; ++q) tot += red[q * m + c];
            atomicAdd(d_bias + c, tot);
        }
    }
}

static SinkConst sink_consts(int n, int m) {
    SinkConst k;
    k.norm = (float)(-log((double)(n + m)));
    k.log_a_m = k.norm + (float)log((double)(n - m));
    return k;
}
static size_t sink_fwd_smem(int n, int m) { return ((size_t)n * (m + 1) + n + (m + 1) + 256 + 8) * sizeof(float); }
static size_t sink_bwd_smem(int n, int m, int iters) {
    return ((size_t)n * (m + 1) + (size_t)(iters + 1) * n + (size_t)iters * (m + 1) + (size_t)iters * n +
            (size_t)iters * (m + 1) + 256) * sizeof(float);
}

// Threads per image of the register-tile Sinkhorn kernels (n <= 256): 512 = two per token row (default), 256 = one per
// row (VPR_SINK_RT=2), 0 = stay on the shared-memory-tile kernels of sinkhorn_scaled.cuh (VPR_SINK_RT=0; A/B timing, and
// the parity test that keeps that path covered at n <= 256).  Read at every call.
static int sink_rt_threads() {
    const char* e = getenv("VPR_SINK_RT");
    if (e != nullptr && e[0] == '0') return 0;
    if (e != nullptr && e[0] == '2') return 256;
    return 512;
}

static int sinkhorn_fwd(const float* S, const float* dust_bin, void* P, int p_dtype, float* U, float* V, int B, int n,
                        int m, int iters, cudaStream_t st, bool pdl = false) {
    VPR_REQUIRE(m == 16 || m == 32 || m == 64 || m == 128, "num_clusters must be 16/32/64/128 (got %d)", m);
    VPR_REQUIRE(n > m, "Sinkhorn needs n > m (n=%d, m=%d)", n, m);
    VPR_REQUIRE(iters >= 1 && iters <= 8, "iters out of range");
    if (m == 64) {
        SinkConst2 k2;
        k2.norm2 = (float)(-log2((double)(n + m)));
        k2.log_a_m2 = k2.norm2 + (float)log2((double)(n - m));
        const size_t guard = sinkfast_fwd_floats<64>(n) * sizeof(float);
        if (n <= RT_ROWS && sink_rt_threads() != 0) {   // 224 x 224 images: register-resident tile, two barriers per iteration
            size_t rs = sinkrt_fwd_floats() * sizeof(float);
            if (rs < guard) rs = guard;
#define VPR_RT_FWD(TP, NT)                                                                                              \
    do {                                                                                                                \
        VPR_SMEM_OK((sinkhorn_fwd_rt<TP, NT>), rs);                                                                     \
        VPR_CUDA_OK(launch_ex(sinkhorn_fwd_rt<TP, NT>, dim3((unsigned)B), dim3(NT), rs, st, pdl, 1, S, dust_bin, (TP*)P, U, V, n, iters, \
                              k2));                                                                                     \
    } while (0)
            const bool fat = sink_rt_threads() == 256;
            if (p_dtype == VPR_BF16) { if (fat) VPR_RT_FWD(bf16, 256); else VPR_RT_FWD(bf16, 512); }
            else                     { if (fat) VPR_RT_FWD(float, 256); else VPR_RT_FWD(float, 512); }
#undef VPR_RT_FWD
            return launch_ok("sinkhorn_fwd_rt");
        }
        size_t fs = sinkscaled_fwd_floats(n) * sizeof(float);
        if (fs < guard) fs = guard;
        if (fs <= 227 * 1024) {          // scaled tile (one exp2 per element), log-domain body as its range guard
            if (p_dtype == VPR_BF16) {
                VPR_SMEM_OK(sinkhorn_fwd_scaled<bf16>, fs);
                VPR_CUDA_OK(launch_ex(sinkhorn_fwd_scaled<bf16>, dim3((unsigned)B), dim3(SS_NT), fs, st, pdl, 1, S, dust_bin, (bf16*)P, U, V, n,
                                      iters, k2));
            } else {
                VPR_SMEM_OK(sinkhorn_fwd_scaled<float>, fs);
                VPR_CUDA_OK(launch_ex(sinkhorn_fwd_scaled<float>, dim3((unsigned)B), dim3(SS_NT), fs, st, pdl, 1, S, dust_bin, (float*)P, U, V,
                                      n, iters, k2));
            }
            return launch_ok("sinkhorn_fwd_scaled");
        }
        VPR_REQUIRE(guard <= 227 * 1024, "score tile [%d,%d] does not fit in shared memory", n, m);
        if (p_dtype == VPR_BF16) {
            VPR_SMEM_OK((sinkhorn_fwd_fast<bf16, 64>), guard);
            sinkhorn_fwd_fast<bf16, 64><<<B, SK_NT, guard, st>>>(S, dust_bin, (bf16*)P, U, V, n, iters, k2);
        } else {
            VPR_SMEM_OK((sinkhorn_fwd_fast<float, 64>), guard);
            sinkhorn_fwd_fast<float, 64><<<B, SK_NT, guard, st>>>(S, dust_bin, (float*)P, U, V, n, iters, k2);
        }
        return launch_ok("sinkhorn_fwd_fast");
    }
    const size_t smem = sink_fwd_smem(n, m);
    VPR_REQUIRE(smem <= 227 * 1024, "score tile [%d,%d] does not fit in shared memory", n, m);
    const SinkConst k = sink_consts(n, m);
    if (p_dtype == VPR_BF16) {
        VPR_SMEM_OK(sinkhorn_fwd_kernel<bf16>, smem);
        sinkhorn_fwd_kernel<bf16><<<B, 256, smem, st>>>(S, dust_bin, (bf16*)P, U, V, n, m, iters, k);
    } else {
        VPR_SMEM_OK(sinkhorn_fwd_kernel<float>, smem);
        sinkhorn_fwd_kernel<float><<<B, 256, smem, st>>>(S, dust_bin, (float*)P, U, V, n, m, iters, k);
    }
    return launch_ok("sinkhorn_fwd_kernel");
}

static int sinkhorn_bwd(const float* S, const float* dust_bin, const float* U, const float* V, const float* dP, void* dS,
                        int ds_dtype, float* d_bias, int B, int n, int m, int iters, cudaStream_t st, bool pdl = false) {
    VPR_REQUIRE(m == 16 || m == 32 || m == 64 || m == 128, "num_clusters must be 16/32/64/128 (got %d)", m);
    VPR_REQUIRE(n > m, "Sinkhorn needs n > m");
    if (m == 64) {
        SinkConst2 k2;
        k2.norm2 = (float)(-log2((double)(n + m)));
        k2.log_a_m2 = k2.norm2 + (float)log2((double)(n - m));
        const size_t guard = sinkfast_bwd_floats<64>(n, iters) * sizeof(float);
        if (n <= RT_ROWS && iters == 3 && sink_rt_threads() != 0) {
            size_t rs = sinkrt_bwd_floats(3) * sizeof(float);
            if (rs < guard) rs = guard;
#define VPR_RT_BWD(TD, NT)                                                                                              \
    do {                                                                                                                \
        VPR_SMEM_OK((sinkhorn_bwd_rt<TD, 3, NT>), rs);                                                                  \
        VPR_CUDA_OK(launch_ex(sinkhorn_bwd_rt<TD, 3, NT>, dim3((unsigned)B), dim3(NT), rs, st, pdl, 1, S, dust_bin, U, V, dP, (TD*)dS,   \
                              d_bias, n, k2));                                                                          \
    } while (0)
            const bool fat = sink_rt_threads() == 256;
            if (ds_dtype == VPR_BF16) { if (fat) VPR_RT_BWD(bf16, 256); else VPR_RT_BWD(bf16, 512); }
            else                      { if (fat) VPR_RT_BWD(float, 256); else VPR_RT_BWD(float, 512); }
#undef VPR_RT_BWD
            return launch_ok("sinkhorn_bwd_rt");
        }
        size_t fs = sinkscaled_bwd_floats(n, iters) * sizeof(float);
        if (fs < guard) fs = guard;
        if (fs <= 227 * 1024) {
#define VPR_SS_BWD(TD, KIT)                                                                                            \
    do {                                                                                                               \
        VPR_SMEM_OK((sinkhorn_bwd_scaled<TD, KIT>), \
                                         (int)fs);                                                                    \
        VPR_CUDA_OK(launch_ex(sinkhorn_bwd_scaled<TD, KIT>, dim3((unsigned)B), dim3(SS_NT), fs, st, pdl, 1, S, dust_bin, U, V, dP, (TD*)dS,  \
                              d_bias, n, iters, k2));                                                                                    \
    } while (0)
            if (ds_dtype == VPR_BF16) { if (iters == 3) VPR_SS_BWD(bf16, 3); else VPR_SS_BWD(bf16, 0); }
            else                      { if (iters == 3) VPR_SS_BWD(float, 3); else VPR_SS_BWD(float, 0); }
#undef VPR_SS_BWD
            return launch_ok("sinkhorn_bwd_scaled");
        }
        VPR_REQUIRE(guard <= 227 * 1024, "score tile [%d,%d] does not fit in shared memory", n, m);
        if (ds_dtype == VPR_BF16) {
            VPR_SMEM_OK((sinkhorn_bwd_fast<bf16, 64>), guard);
            sinkhorn_bwd_fast<bf16, 64><<<B, SK_NT, guard, st>>>(S, dust_bin, U, V, dP, (bf16*)dS, d_bias, n, iters, k2);
        } else {
            VPR_SMEM_OK((sinkhorn_bwd_fast<float, 64>), guard);
            sinkhorn_bwd_fast<float, 64><<<B, SK_NT, guard, st>>>(S, dust_bin, U, V, dP, (float*)dS, d_bias, n, iters, k2);
        }
        return launch_ok("sinkhorn_bwd_fast");
    }
    const size_t smem = sink_bwd_smem(n, m, iters);
    VPR_REQUIRE(smem <= 227 * 1024, "score tile [%d,%d] does not fit in shared memory", n, m);
    const SinkConst k = sink_consts(n, m);
    if (ds_dtype == VPR_BF16) {
        VPR_SMEM_OK(sinkhorn_bwd_kernel<bf16>, smem);
        sinkhorn_bwd_kernel<bf16><<<B, 256, smem, st>>>(S, dust_bin, U, V, dP, (bf16*)dS, d_bias, n, m, iters, k);
    } else {
        VPR_SMEM_OK(sinkhorn_bwd_kernel<float>, smem);
        sinkhorn_bwd_kernel<float><<<B, 256, smem, st>>>(S, dust_bin, U, V, dP, (float*)dS, d_bias, n, m, iters, k);
    }
    return launch_ok("sinkhorn_bwd_kernel");
}

// ---------------------------------------------------------------------------
// descriptor assembly: intra-cluster L2 norm (over l), token L2 norm, concat, global L2 norm
// ---------------------------------------------------------------------------
#define NORM_EPS 1e-12f

template <typename T>
__global__ void __launch_bounds__(256) finalize_fwd_kernel(const float* __restrict__ AGG, const float* __restrict__ TF,
                                                           T* __restrict__ out, int l, int m, int g) {
    __shared__ float red[256];
    __shared__ float inv_c[128];
    __shared__ float red2[8];
    const int b = blockIdx.x, tid = threadIdx.x;
    const float* A = AGG + (size_t)b * l * m;
    const float* t = TF + (size_t)b * g;
    const int parts = 256 / m, c = tid % m, part = tid / m;
    float acc = 0.f;
    for (int li = part; li < l; li += parts) { const float v = A[li * m + c]; acc += v * v; }
    red[tid] = acc;
    __syncthreads();
    float zz = 0.f;   // ||z||^2 contribution of this thread's cluster
    if (part == 0) {
        float tot = red[c];
        for (int q = 1; q < parts; ++q) tot += red[q * m + c];
        const float nc = fmaxf(sqrtf(tot), NORM_EPS);
        inv_c[c] = 1.f / nc;
        zz = tot / (nc * nc);
    }
    float ts = 0.f;
    for (int i = tid; i < g; i += 256) { const float v = t[i]; ts += v * v; }
    ts = block_sum256(ts, red2);
    const float nt = fmaxf(sqrtf(ts), NORM_EPS);
    zz = block_sum256(zz, red2) + ts / (nt * nt);
    const float invN = 1.f / fmaxf(sqrtf(zz), NORM_EPS);
    T* o = out + (size_t)b * (g + l * m);
    const float invt = invN / nt;
    for (int i = tid; i < g; i += 256) stf(o + i, t[i] * invt);
    for (int e = tid; e < l * m; e += 256) stf(o + g + e, A[e] * inv_c[e % m] * invN);
}

template <typename T>
__global__ void __launch_bounds__(256) finalize_bwd_kernel(const float* __restrict__ AGG, const float* __restrict__ TF,
                                                           const T* __restrict__ d_out, T* __restrict__ dAGG,
                                                           T* __restrict__ dTF, float* __restrict__ g_bias_tok, int l,
                                                           int m, int g) {
    __shared__ float red[256];
    __shared__ float inv_c[128];
    __shared__ float dot_c[128];
    __shared__ float red2[8];
    const int b = blockIdx.x, tid = threadIdx.x;
    const float* A = AGG + (size_t)b * l * m;
    const float* t = TF + (size_t)b * g;
    const T* dy = d_out + (size_t)b * (g + l * m);
    const int parts = 256 / m, c = tid % m, part = tid / m;
    // norms (recomputed)
    float acc = 0.f;
    for (int li = part; li < l; li += parts) { const float v = A[li * m + c]; acc += v * v; }
    red[tid] = acc;
    __syncthreads();
    float zz = 0.f;
    if (part == 0) {
        float tot = red[c];
        for (int q = 1; q < parts; ++q) tot += red[q * m + c];
        const float nc = fmaxf(sqrtf(tot), NORM_EPS);
        inv_c[c] = 1.f / nc;
        zz = tot / (nc * nc);
    }
    float ts = 0.f;
    for (int i = tid; i < g; i += 256) { const float v = t[i]; ts += v * v; }
    ts = block_sum256(ts, red2);
    const float nt = fmaxf(sqrtf(ts), NORM_EPS);
    const float invt = 1.f / nt;
    zz = block_sum256(zz, red2) + ts * invt * invt;
    const float invN = 1.f / fmaxf(sqrtf(zz), NORM_EPS);
    // s = y . dy   with y = z * invN
    float s = 0.f;
    for (int i = tid; i < g; i += 256) s += t[i] * invt * invN * ldf(dy + i);
    for (int e = tid; e < l * m; e += 256) s += A[e] * inv_c[e % m] * invN * ldf(dy + g + e);
    s = block_sum256(s, red2);
    // dz = (dy - y s) * invN ; token part: dt = (dz - that (that . dz)) / nt
    float dt_dot = 0.f;
    for (int i = tid; i < g; i += 256) {
        const float th = t[i] * invt;
        const float dz = (ldf(dy + i) - th * invN * s) * invN;
        dt_dot += th * dz;
    }
    dt_dot = block_sum256(dt_dot, red2);
    T* dT = dTF + (size_t)b * g;
    for (int i = tid; i < g; i += 256) {
        const float th = t[i] * invt;
        const float dz = (ldf(dy + i) - th * invN * s) * invN;
        const float dti = (dz - th * dt_dot) * invt;
        stf(dT + i, dti);
        if (g_bias_tok != nullptr) atomicAdd(g_bias_tok + i, dti);
    }
    // cluster part: per column c, dot_c = sum_l ahat_lc dz_lc
    float dc = 0.f;
    for (int li = part; li < l; li += parts) {
        const float ah = A[li * m + c] * inv_c[c];
        const float dz = (ldf(dy + g + li * m + c) - ah * invN * s) * invN;
        dc += ah * dz;
    }
    __syncthreads();
    red[tid] = dc;
    __syncthreads();
    if (part == 0) {
        float tot = red[c];
        for (int q = 1; q < parts; ++q) tot += red[q * m + c];
        dot_c[c] = tot;
    }
    __syncthreads();
    T* dA = dAGG + (size_t)b * l * m;
    for (int e = tid; e < l * m; e += 256) {
        const int cc = e % m;
        const float ah = A[e] * inv_c[cc];
        const float dz = (ldf(dy + g + e) - ah * invN * s) * invN;
        stf(dA + e, (dz - ah * dot_c[cc]) * inv_c[cc]);
    }
}

// out[c] += sum over rows of x[r][c]  (token_features.2.bias gradient = column sums of dTF); 32 rows per CTA
template <typename T>
__global__ void __launch_bounds__(256) colsum_rows_kernel(const T* __restrict__ x, float* __restrict__ out, int rows, int cols) {
    const int r0 = blockIdx.x * 32, r1 = min(rows, r0 + 32);
    for (int c = threadIdx.x; c < cols; c += 256) {
        float acc = 0.f;
        for (int r = r0; r < r1; ++r) acc += ldf(x + (size_t)r * cols + c);
        atomicAdd(out + c, acc);
    }
}

__global__ void counter_advance_kernel(unsigned long long* c) {
    pdl_wait();                     // launched programmatically: resident early, increments only once the previous launch is done
    pdl_launch_dependents();
    *c += 1ull;
}

__global__ void dropout_mask_kernel(unsigned long long seed, uint32_t thr, long long total, uint8_t* out) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x)
        out[i] = dropout_keep(seed, (unsigned long long)i, thr) ? 1 : 0;
}

// ---------------------------------------------------------------------------
// buffer plans
// ---------------------------------------------------------------------------
struct SavePlan {
    size_t w0cat, w1c, w1s, wt0, wt1, b0cat, H, F, S, P, U, V, AGG, HT, TF, total;
};
struct WorkPlan {
    size_t dAGG, dAGGT, dTF, dF, dP, dS, dH, dHT, total;
};

static SavePlan save_plan(const VprSaladShape& s, int dtype) {
    const size_t e = dtype == VPR_BF16 ? 2 : 4;
    const size_t T = (size_t)s.B * s.n, h = s.hidden;
    SavePlan p;
    size_t off = 0;
    p.w0cat = carve_size(off, 2 * h * s.d * e);
    p.w1c = carve_size(off, (size_t)s.l * h * e);
    p.w1s = carve_size(off, (size_t)s.m * h * e);
    p.wt0 = carve_size(off, h * s.d * e);
    p.wt1 = carve_size(off, (size_t)s.g * h * e);
    p.b0cat = carve_size(off, 2 * h * 4);
    p.H = carve_size(off, T * 2 * h * e);
    p.F = carve_size(off, T * s.l * e);
    p.S = carve_size(off, T * s.m * 4);
    p.P = carve_size(off, T * s.m * e);
    p.U = carve_size(off, (size_t)s.B * s.iters * (s.m + 1) * 4);
    p.V = carve_size(off, (size_t)s.B * s.iters * s.n * 4);
    p.AGG = carve_size(off, (size_t)s.B * s.l * s.m * 4);
    p.HT = carve_size(off, (size_t)s.B * h * e);
    p.TF = carve_size(off, (size_t)s.B * s.g * 4);
    p.total = (off + 255) & ~size_t(255);
    return p;
}

static WorkPlan work_plan(const VprSaladShape& s, int dtype) {
    const size_t e = dtype == VPR_BF16 ? 2 : 4;
    const size_t T = (size_t)s.B * s.n, h = s.hidden;
    WorkPlan p;
    size_t off = 0;
    p.dAGG = carve_size(off, (size_t)s.B * s.l * s.m * e);
    p.dAGGT = carve_size(off, (size_t)s.B * s.l * s.m * e);
    p.dTF = carve_size(off, (size_t)s.B * s.g * e);
    p.dF = carve_size(off, T * s.l * e);
    p.dP = carve_size(off, T * s.m * 4);
    p.dS = carve_size(off, T * s.m * e);
    p.dH = carve_size(off, T * 2 * h * e);
    p.dHT = carve_size(off, (size_t)s.B * h * e);
    p.total = (off + 255) & ~size_t(255);
    return p;
}

static int check_shape(const VprSaladShape& s, int dtype) {
    VPR_REQUIRE(dtype == VPR_F32 || dtype == VPR_BF16, "dtype must be VPR_F32 or VPR_BF16");
    VPR_REQUIRE(s.B > 0 && s.n > 0, "empty batch (B=%d, n=%d)", s.B, s.n);
    VPR_REQUIRE(s.n > s.m, "SALAD needs more patch tokens than clusters (n=%d, m=%d)", s.n, s.m);
    VPR_REQUIRE(s.m == 16 || s.m == 32 || s.m == 64 || s.m == 128, "num_clusters must be 16/32/64/128");
    VPR_REQUIRE(s.d % 8 == 0 && s.hidden % 8 == 0 && s.l % 8 == 0 && s.g % 8 == 0,
                "channel counts must be multiples of 8 (d=%d hidden=%d l=%d g=%d)", s.d, s.hidden, s.l, s.g);
    VPR_REQUIRE(s.iters >= 1 && s.iters <= 8, "iters out of range");
    return 0;
}

template <typename T>
static int salad_fwd_t(const VprSaladParams& p, const T* x, const T* t, T* out, char* save, const VprSaladShape& s,
                       int dtype, float drop_p, uint64_t seed, const uint64_t* seed_dev, cudaStream_t st) {
    const SavePlan sp = save_plan(s, dtype);
    const int h = s.hidden;
    const long long Tk = (long long)s.B * s.n;
    T* w0cat = (T*)(save + sp.w0cat); T* w1c = (T*)(save + sp.w1c); T* w1s = (T*)(save + sp.w1s);
    T* wt0 = (T*)(save + sp.wt0); T* wt1 = (T*)(save + sp.wt1); float* b0cat = (float*)(save + sp.b0cat);
    T* H = (T*)(save + sp.H); T* F = (T*)(save + sp.F); float* S = (float*)(save + sp.S); T* P = (T*)(save + sp.P);
    float* U = (float*)(save + sp.U); float* V = (float*)(save + sp.V); float* AGG = (float*)(save + sp.AGG);
    T* HT = (T*)(save + sp.HT); float* TF = (float*)(save + sp.TF);
    const int okT = dtype == VPR_BF16 ? OUT_BF16 : OUT_F32;

    {
        const long long n0 = (long long)h * s.d;
        PackPlan plan;
        plan.seg[0] = {p.cluster_w0, w0cat, n0 / 4};
        plan.seg[1] = {p.score_w0, w0cat + n0, n0 / 4};
        plan.seg[2] = {p.cluster_w1, w1c, (long long)s.l * h / 4};
        plan.seg[3] = {p.score_w1, w1s, (long long)s.m * h / 4};
        plan.seg[4] = {p.token_w0, wt0, n0 / 4};
        plan.seg[5] = {p.token_w1, wt1, (long long)s.g * h / 4};
        plan.seg[6] = {p.cluster_b0, b0cat, h / 4};
        plan.seg[7] = {p.score_b0, b0cat + h, h / 4};
        plan.total4 = 0;
        for (int k = 0; k < 8; ++k) {
            VPR_REQUIRE((reinterpret_cast<uintptr_t>(plan.seg[k].src) & 15) == 0, "parameter tensors must be 16-byte aligned");
            plan.total4 += plan.seg[k].n4;
        }
        int blocks = (int)((plan.total4 + 255) / 256);
        if (blocks > num_sms() * 8) blocks = num_sms() * 8;
        pack_params_kernel<T><<<blocks, 256, 0, st>>>(plan);
        VPR_TRY(launch_ok("pack_params_kernel"));
    }

    VPR_REQUIRE(Tk < (1ll << 31), "too many tokens");
    GemmDesc g[3];
    // launch 1: H = dropout(relu(X W0cat^T + b0cat)) [T, 2h]   +   HT = relu(t Wt0^T + bt0) [B, h]
    g[0] = GemmDesc(); g[0].dtype = dtype;
    g[0].A = x; g[0].lda = s.d; g[0].B = w0cat; g[0].ldb = s.d; g[0].M = (int)Tk; g[0].N = 2 * h; g[0].K = s.d;
    g[0].out = H; g[0].out_kind = okT; g[0].ldo = 2 * h; g[0].bias = b0cat; g[0].relu = 1;
    g[0].dropout_p = drop_p; g[0].dropout_seed = seed; g[0].dropout_seed_dev = seed_dev;
    g[1] = GemmDesc(); g[1].dtype = dtype;
    g[1].A = t; g[1].lda = s.d; g[1].B = wt0; g[1].ldb = s.d; g[1].M = s.B; g[1].N = h; g[1].K = s.d;
    g[1].out = HT; g[1].out_kind = okT; g[1].ldo = h; g[1].bias = p.token_b0; g[1].relu = 1;
    g[0].pdl = 1;        // a chain of our own launches (pack -> H -> F/S/TF -> Sinkhorn -> aggregation): each prologue
                         // overlaps the previous launch's tail
    VPR_TRY(gemm_run_group(g, 2, st));
    // launch 2: F = H[:, :h] W1c^T + b1c [T, l]   S = H[:, h:] W1s^T + b1s [T, m] fp32   TF = HT Wt1^T + bt1 [B, g] fp32
    const bool fin_fast = finalize_fast_ok(s.l, s.m, s.g);
    const bool fs_fused = s.g == 256 && fs_fwd_supported(dtype, s.l, s.m, h);
    if (fs_fused) {
        if constexpr (sizeof(T) == 2) {
            FsFwdDesc fd;
            fd.H = H; fd.W1c = w1c; fd.W1s = w1s; fd.bias_f = p.cluster_b1; fd.bias_s = p.score_b1;
            fd.F = F; fd.S = S; fd.T = Tk; fd.hidden = h;
            fd.HT = HT; fd.Wt1 = wt1; fd.bias_t = p.token_b1; fd.TF = TF; fd.B = s.B; fd.g = s.g;
            fd.pdl = 1;
            VPR_TRY(fs_fwd_run(fd, st));
        }
    } else {
        g[0] = GemmDesc(); g[0].dtype = dtype;
        g[0].A = H; g[0].lda = 2 * h; g[0].B = w1c; g[0].ldb = h; g[0].M = (int)Tk; g[0].N = s.l; g[0].K = h;
        g[0].out = F; g[0].out_kind = okT; g[0].ldo = s.l; g[0].bias = p.cluster_b1;
        g[1] = GemmDesc(); g[1].dtype = dtype;
        g[1].A = H + h; g[1].lda = 2 * h; g[1].B = w1s; g[1].ldb = h; g[1].M = (int)Tk; g[1].N = s.m; g[1].K = h;
        g[1].out = S; g[1].out_kind = OUT_F32; g[1].ldo = s.m; g[1].bias = p.score_b1;
        g[2] = GemmDesc(); g[2].dtype = dtype;
        g[2].A = HT; g[2].lda = h; g[2].B = wt1; g[2].ldb = h; g[2].M = s.B; g[2].N = s.g; g[2].K = h;
        g[2].out = TF; g[2].out_kind = OUT_F32; g[2].ldo = s.g; g[2].bias = p.token_b1; g[2].bn = 128;
        g[0].pdl = 1;
        VPR_TRY(gemm_run_group(g, 3, st));
    }
    if (sink_agg_fwd_supported(dtype, s.l, s.m, s.g, s.n, s.iters) && agg_fwd_supported(dtype, s.l, s.m, s.g, s.n)) {
        if constexpr (sizeof(T) == 2) {          // Sinkhorn + aggregation + descriptor assembly in one launch, one CTA per image
            SinkAggFwdDesc sd;
            sd.S = S; sd.dust_bin = p.dust_bin; sd.F = F; sd.P = P; sd.U = U; sd.V = V; sd.TF = TF; sd.AGG = AGG; sd.out = out;
            sd.B = s.B; sd.n = s.n; sd.g = s.g; sd.iters = s.iters;
            sd.pdl = 1;
            return sink_agg_fwd_run(sd, st);
        }
    }
    // P = sinkhorn(S)
    VPR_TRY(sinkhorn_fwd(S, p.dust_bin, P, dtype, U, V, s.B, s.n, s.m, s.iters, st, true));
    if (agg_fwd_supported(dtype, s.l, s.m, s.g, s.n)) {
        if constexpr (sizeof(T) == 2) {          // aggregation + descriptor assembly in one launch, one CTA per image
            AggFwdDesc ad;
            ad.F = F; ad.P = P; ad.TF = TF; ad.AGG = AGG; ad.out = out; ad.B = s.B; ad.n = s.n; ad.g = s.g;
            ad.pdl = 1;
            return agg_fwd_run(ad, st);
        }
    }
    // AGG[b][l][c] = sum_j F[b][j][l] P[b][j][c]      both operands MN-major, K = tokens of image b
    g[0] = GemmDesc(); g[0].dtype = dtype;
    g[0].A = F; g[0].a_mn = 1; g[0].lda = s.l; g[0].a_bs = (int64_t)s.n * s.l;
    g[0].B = P; g[0].b_mn = 1; g[0].ldb = s.m; g[0].b_bs = (int64_t)s.n * s.m;
    g[0].M = s.l; g[0].N = s.m; g[0].K = s.n; g[0].batch = s.B;
    g[0].out = AGG; g[0].out_kind = OUT_F32; g[0].ldo = s.m; g[0].out_bs = (int64_t)s.l * s.m;
    VPR_TRY(gemm_run(g[0], st));
    if (fin_fast) {
        finalize_fwd_fast<T><<<s.B, 256, 0, st>>>(AGG, TF, out, s.l, s.m, s.g, HT, (const T*)nullptr, p.token_b1, h);
        return launch_ok("finalize_fwd_fast");
    }
    finalize_fwd_kernel<T><<<s.B, 256, 0, st>>>(AGG, TF, out, s.l, s.m, s.g);
    return launch_ok("finalize_fwd_kernel");
}

template <typename T>
static int salad_bwd_t(const VprSaladParams& p, const T* x, const T* t, const T* d_out, const char* save, char* ws,
                       VprSaladGrads& gr, T* dx, T* dt, const VprSaladShape& s, int dtype, float drop_p,
                       int phase, int reserve_sms, cudaStream_t st) {
    const SavePlan sp = save_plan(s, dtype);
    const WorkPlan wp = work_plan(s, dtype);
    const int h = s.hidden;
    const long long Tk = (long long)s.B * s.n;
    const T* w0cat = (const T*)(save + sp.w0cat); const T* w1c = (const T*)(save + sp.w1c);
    const T* w1s = (const T*)(save + sp.w1s); const T* wt0 = (const T*)(save + sp.wt0);
    const T* wt1 = (const T*)(save + sp.wt1);
    const T* H = (const T*)(save + sp.H); const T* F = (const T*)(save + sp.F); const float* S = (const float*)(save + sp.S);
    const T* P = (const T*)(save + sp.P); const float* U = (const float*)(save + sp.U);
    const float* V = (const float*)(save + sp.V); const float* AGG = (const float*)(save + sp.AGG);
    const T* HT = (const T*)(save + sp.HT); const float* TF = (const float*)(save + sp.TF);
    T* dAGG = (T*)(ws + wp.dAGG); T* dTF = (T*)(ws + wp.dTF); T* dF = (T*)(ws + wp.dF); float* dP = (float*)(ws + wp.dP);
    T* dS = (T*)(ws + wp.dS); T* dH = (T*)(ws + wp.dH); T* dHT = (T*)(ws + wp.dHT);
    const int okT = dtype == VPR_BF16 ? OUT_BF16 : OUT_F32;
    const float keep_scale = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
    const int nsm = num_sms();

    if (phase != VPR_BWD_INPUTS) {
    const bool fused = dh_fused_supported(dtype, s.l, s.m, h);
    const bool per_image = fused && h == 512 && s.g <= 1024 && sizeof(T) == 2 && agg_bwd_supported(dtype, s.l, s.m, s.g);
    static int use_side = -1;
    if (use_side < 0) { const char* e = getenv("VPR_TOKEN_SIDE"); use_side = (e != nullptr && e[0] == '0') ? 0 : 1; }
    SideBranch* side = nullptr;                            // memset + token-MLP gradients running beside the main chain
    cudaStream_t zst = st;
    if (per_image && use_side) {
        // the gradient accumulators are first touched by the Sinkhorn backward: zero them on the side stream while the
        // per-image aggregation backward (which then accumulates nothing) runs on the main one
        VPR_TRY(side_branch(&side));
        VPR_CUDA_OK(cudaEventRecord(side->fork, st));
        VPR_CUDA_OK(cudaStreamWaitEvent(side->stream, side->fork, 0));
        zst = side->stream;
    }
    // zero every gradient buffer that is accumulated with atomics (adjacent buffers -> one memset)
    {
        struct Rng { float* p; size_t n; };
        const Rng r[13] = {{gr.cluster_w0, (size_t)h * s.d}, {gr.cluster_b0, (size_t)h}, {gr.cluster_w1, (size_t)s.l * h},
                           {gr.cluster_b1, (size_t)s.l}, {gr.score_w0, (size_t)h * s.d}, {gr.score_b0, (size_t)h},
                           {gr.score_w1, (size_t)s.m * h}, {gr.score_b1, (size_t)s.m}, {gr.token_w0, (size_t)h * s.d},
                           {gr.token_b0, (size_t)h}, {gr.token_w1, (size_t)s.g * h}, {gr.token_b1, (size_t)s.g},
                           {gr.dust_bin, 1}};    // d/d dust_bin == 0 identically (SURVEY A-3b)
        int i = 0;
        while (i < 13) {
            float* p0 = r[i].p;
            size_t cnt = r[i].n;
            int j = i + 1;
            while (j < 13 && r[j].p == p0 + cnt) { cnt += r[j].n; ++j; }
            VPR_CUDA_OK(cudaMemsetAsync(p0, 0, cnt * sizeof(float), zst));
            i = j;
        }
    }
    if (side != nullptr) VPR_CUDA_OK(cudaEventRecord(side->zeroed, side->stream));

    GemmDesc g[3];
    const int kb_tok = (int)((Tk + 63) / 64);
    const bool fin_fast = finalize_fast_ok(s.l, s.m, s.g);
    bool fin_dht = false;                                  // dHT already produced (finalize_bwd_fast / side-branch GEMM)
    bool tok_side = false;
    if (per_image) {
        // ---- one CTA per image: cluster part of the norm backward + dF + dP (dAGG stays in shared memory) ----
        AggBwdDesc ad;
        ad.AGG = AGG; ad.TF = TF; ad.d_out = d_out; ad.P = P; ad.F = F; ad.dF = dF; ad.dP = dP;
        ad.dTF = dTF; ad.g_bias_tok = side != nullptr ? nullptr : gr.token_b1;    // side branch: column sums taken there
        ad.B = s.B; ad.n = s.n; ad.g = s.g;
        ad.pdl = side != nullptr ? 1 : 0;                // no memset node between the previous launch and this one
        VPR_TRY(agg_bwd_run(ad, st));
        tok_side = true;
    } else {
    // norms backward: d_out -> dAGG [B,l,m] (and its transpose [B,m,l]), dTF [B,g] (+ token_features.2.bias gradient)
    T* dAGGT = (T*)(ws + wp.dAGGT);
    if (fin_fast) {
        const size_t fsm = finalize_bwd_fast_smem<T>(s.l, s.m);
        VPR_SMEM_OK(finalize_bwd_fast<T>, fsm);
        fin_dht = h == 512;                                // the in-kernel token-MLP gradient is laid out for 512 hidden units
        finalize_bwd_fast<T><<<s.B, 256, fsm, st>>>(AGG, TF, d_out, dAGG, dAGGT, dTF, gr.token_b1, s.l, s.m, s.g,
                                                    wt1, HT, fin_dht ? dHT : (T*)nullptr, gr.token_b0, h);
        VPR_TRY(launch_ok("finalize_bwd_fast"));
    } else {
        finalize_bwd_kernel<T><<<s.B, 256, 0, st>>>(AGG, TF, d_out, dAGG, dTF, gr.token_b1, s.l, s.m, s.g);
        VPR_TRY(launch_ok("finalize_bwd_kernel"));
    }

    // ---- aggregation backward (one launch when the transposed dAGG is available) ----
    g[0] = GemmDesc(); g[0].dtype = dtype;    // dF[b][j][l] = sum_c P[b][j][c] dAGG[b][l][c]   (+ cluster bias gradient)
    g[0].A = P; g[0].lda = s.m; g[0].a_bs = (int64_t)s.n * s.m; g[0].B = dAGG; g[0].ldb = s.m; g[0].b_bs = (int64_t)s.l * s.m;
    g[0].M = s.n; g[0].N = s.l; g[0].K = s.m; g[0].batch = s.B;
    g[0].out = dF; g[0].out_kind = okT; g[0].ldo = s.l; g[0].out_bs = (int64_t)s.n * s.l;
    g[0].colsum = fused ? nullptr : gr.cluster_b1;       // the fused dH launch sums dF's columns on the tensor core
    g[1] = GemmDesc(); g[1].dtype = dtype;    // dP[b][j][c] = sum_l F[b][j][l] dAGG[b][l][c]
    g[1].A = F; g[1].lda = s.l; g[1].a_bs = (int64_t)s.n * s.l;
    g[1].M = s.n; g[1].N = s.m; g[1].K = s.l; g[1].batch = s.B;
    g[1].out = dP; g[1].out_kind = OUT_F32; g[1].ldo = s.m; g[1].out_bs = (int64_t)s.n * s.m;
    if (fin_fast) {
        g[1].B = dAGGT; g[1].ldb = s.l; g[1].b_bs = (int64_t)s.l * s.m;
        VPR_TRY(gemm_run_group(g, 2, st));
    } else {
        g[1].B = dAGG; g[1].b_mn = 1; g[1].ldb = s.m; g[1].b_bs = (int64_t)s.l * s.m;
        VPR_TRY(gemm_run(g[0], st));
        VPR_TRY(gemm_run(g[1], st));
    }
    }
    // dWt1[g,h] = dTF^T HT (token MLP, K = B): rides along with whichever split-K launch comes next
    GemmDesc gtok1 = GemmDesc(); gtok1.dtype = dtype;
    gtok1.A = dTF; gtok1.a_mn = 1; gtok1.lda = s.g; gtok1.B = HT; gtok1.b_mn = 1; gtok1.ldb = h; gtok1.M = s.g; gtok1.N = h;
    gtok1.K = s.B; gtok1.out = gr.token_w1; gtok1.out_kind = OUT_ATOMIC_F32; gtok1.ldo = h;
    // dHT = (dTF Wt1) * (HT > 0)  (+ token_features.0.bias gradient)
    GemmDesc gtokh = GemmDesc(); gtokh.dtype = dtype;
    gtokh.A = dTF; gtokh.lda = s.g; gtokh.B = wt1; gtokh.b_mn = 1; gtokh.ldb = h; gtokh.M = s.B; gtokh.N = h; gtokh.K = s.g;
    gtokh.out = dHT; gtokh.out_kind = okT; gtokh.ldo = h; gtokh.mask = HT; gtokh.ld_mask = h; gtokh.colsum = gr.token_b0;
    if (tok_side) {
        // dHT depends only on dTF: a few tiles on a side branch next to the Sinkhorn backward / fused dH launch
        if (side != nullptr) {
            VPR_CUDA_OK(cudaEventRecord(side->fork2, st));                     // dTF is ready
            VPR_CUDA_OK(cudaStreamWaitEvent(side->stream, side->fork2, 0));
            colsum_rows_kernel<T><<<(s.B + 31) / 32, 256, 0, side->stream>>>(dTF, gr.token_b1, s.B, s.g);
            VPR_TRY(launch_ok("colsum_rows_kernel"));
            VPR_TRY(gemm_run(gtokh, side->stream));
            VPR_CUDA_OK(cudaEventRecord(side->join, side->stream));
            VPR_CUDA_OK(cudaStreamWaitEvent(st, side->zeroed, 0));             // the Sinkhorn backward accumulates a bias gradient
        } else {
            VPR_TRY(gemm_run(gtokh, st));
        }
        fin_dht = true;
    }
    // ---- Sinkhorn backward: dP -> dS (+ score bias gradient) ----
    VPR_TRY(sinkhorn_bwd(S, p.dust_bin, U, V, dP, dS, dtype, gr.score_b1, s.B, s.n, s.m, s.iters, st, true));

    if (fused) {
        // ---- one pass over H: dH, the second-layer weight gradients and the first-layer bias gradients ----
        DhFusedDesc fd;
        fd.dF = dF; fd.dS = dS; fd.W1c = w1c; fd.W1s = w1s; fd.H = H; fd.dH = dH;
        fd.dW1c = gr.cluster_w1; fd.dW1s = gr.score_w1; fd.db0c = gr.cluster_b0; fd.db0s = gr.score_b0;
        fd.db1c = gr.cluster_b1;
        fd.T = Tk; fd.hidden = h; fd.mask_scale = keep_scale;
        fd.pdl = 1;
        fd.reserve_sms = reserve_sms;
        VPR_TRY(dh_fused_run(fd, st));
        if (!fin_dht) VPR_TRY(gemm_run(gtokh, st));       // otherwise finalize_bwd_fast already produced dHT
    } else {
        // ---- second-layer weight gradients (split-K + fp32 reductions), one launch ----
        g[0] = GemmDesc(); g[0].dtype = dtype;    // dW1c[l,h] = dF^T H[:, :h]
        g[0].A = dF; g[0].a_mn = 1; g[0].lda = s.l; g[0].B = H; g[0].b_mn = 1; g[0].ldb = 2 * h; g[0].M = s.l; g[0].N = h;
        g[0].K = (int)Tk; g[0].out = gr.cluster_w1; g[0].out_kind = OUT_ATOMIC_F32; g[0].ldo = h;
        g[1] = GemmDesc(); g[1].dtype = dtype;    // dW1s[m,h] = dS^T H[:, h:]
        g[1].A = dS; g[1].a_mn = 1; g[1].lda = s.m; g[1].B = H + h; g[1].b_mn = 1; g[1].ldb = 2 * h; g[1].M = s.m; g[1].N = h;
        g[1].K = (int)Tk; g[1].out = gr.score_w1; g[1].out_kind = OUT_ATOMIC_F32; g[1].ldo = h;
        g[2] = gtok1;
        // split K so that the two token-long problems fill the machine; the token-MLP problem's few short
        // work items come last and are absorbed by the CTAs that finish first
        const GemmTilePlan tp = gemm_tile_plan(g, 3);
        const int tiles = ((s.l + tp.tile_m - 1) / tp.tile_m + (s.m + tp.tile_m - 1) / tp.tile_m) * ((h + tp.tile_n - 1) / tp.tile_n);
        int ks = tp.slots / tiles; if (ks < 1) ks = 1; if (ks > kb_tok) ks = kb_tok;
        g[0].k_splits = ks; g[1].k_splits = ks;
        VPR_TRY(gemm_run_group(g, 3, st));
        // ---- dH = [dF W1c | dS W1s] * (H > 0) * keep_scale,  dHT; first-layer bias gradients ----
        g[0] = GemmDesc(); g[0].dtype = dtype;
        g[0].A = dF; g[0].lda = s.l; g[0].B = w1c; g[0].b_mn = 1; g[0].ldb = h; g[0].M = (int)Tk; g[0].N = h; g[0].K = s.l;
        g[0].out = dH; g[0].out_kind = okT; g[0].ldo = 2 * h; g[0].mask = H; g[0].ld_mask = 2 * h; g[0].mask_scale = keep_scale;
        g[0].colsum = gr.cluster_b0;
        g[1] = GemmDesc(); g[1].dtype = dtype;
        g[1].A = dS; g[1].lda = s.m; g[1].B = w1s; g[1].b_mn = 1; g[1].ldb = h; g[1].M = (int)Tk; g[1].N = h; g[1].K = s.m;
        g[1].out = dH + h; g[1].out_kind = okT; g[1].ldo = 2 * h; g[1].mask = H + h; g[1].ld_mask = 2 * h; g[1].mask_scale = keep_scale;
        g[1].colsum = gr.score_b0;
        g[2] = gtokh;
        VPR_TRY(gemm_run_group(g, fin_dht ? 2 : 3, st));      // finalize_bwd_fast already produced dHT
    }
    if (side != nullptr) VPR_CUDA_OK(cudaStreamWaitEvent(st, side->join, 0));      // dTF / dHT / token bias gradients are in
    // ---- first-layer weight gradients: dW0[h,d] = dH^T X (two parameter tensors) and dWt0 = dHT^T t ----
    {
        GemmDesc g4[4];
        for (int half = 0; half < 2; ++half) {
            g4[half] = GemmDesc(); g4[half].dtype = dtype;
            g4[half].A = dH + half * h; g4[half].a_mn = 1; g4[half].lda = 2 * h; g4[half].B = x; g4[half].b_mn = 1; g4[half].ldb = s.d;
            g4[half].M = h; g4[half].N = s.d; g4[half].K = (int)Tk;
            g4[half].out = half == 0 ? gr.cluster_w0 : gr.score_w0; g4[half].out_kind = OUT_ATOMIC_F32; g4[half].ldo = s.d;
        }
        g4[2] = GemmDesc(); g4[2].dtype = dtype;
        g4[2].A = dHT; g4[2].a_mn = 1; g4[2].lda = h; g4[2].B = t; g4[2].b_mn = 1; g4[2].ldb = s.d; g4[2].M = h; g4[2].N = s.d; g4[2].K = s.B;
        g4[2].out = gr.token_w0; g4[2].out_kind = OUT_ATOMIC_F32; g4[2].ldo = s.d;
        int cnt = 3;
        if (fused) g4[cnt++] = gtok1;
        const GemmTilePlan tp = gemm_tile_plan(g4, cnt);
        const int tiles = 2 * ((h + tp.tile_m - 1) / tp.tile_m) * ((s.d + tp.tile_n - 1) / tp.tile_n);
        int ks = tp.slots / tiles; if (ks < 1) ks = 1; if (ks > kb_tok) ks = kb_tok;
        g4[0].k_splits = ks; g4[1].k_splits = ks;
        g4[0].pdl = 1;
        // no reserve_sms here: this launch is ONE wave of split-K work items sized to the machine (12 pair tiles x 6
        // K-splits on 74 CTA pairs); with fewer CTAs the leftover items form a second wave and the launch takes twice
        // as long (measured: 79 -> 132 us with 8 SMs held back).  A collective that runs beside the backward has to be
        // done by the time this launch starts.
        VPR_TRY(gemm_run_group(g4, cnt, st));
    }
    }   // phase != VPR_BWD_INPUTS
    if (phase == VPR_BWD_PARAMS) return 0;
    // ---- input gradients: dX = dH W0cat, dt = dHT Wt0 ----
    {
        GemmDesc g[2];
        int cnt = 0;
        if (dx != nullptr) {
            g[cnt] = GemmDesc(); g[cnt].dtype = dtype;
            g[cnt].A = dH; g[cnt].lda = 2 * h; g[cnt].B = w0cat; g[cnt].b_mn = 1; g[cnt].ldb = s.d; g[cnt].M = (int)Tk; g[cnt].N = s.d;
            g[cnt].K = 2 * h; g[cnt].out = dx; g[cnt].out_kind = okT; g[cnt].ldo = s.d;
            g[cnt].reserve_sms = reserve_sms;
            g[cnt].pdl = phase == VPR_BWD_ALL ? 2 : 0;     // one-call form: follows the dW0 launch, reads nothing it writes
            ++cnt;
        }
        if (dt != nullptr) {
            g[cnt] = GemmDesc(); g[cnt].dtype = dtype;
            g[cnt].A = dHT; g[cnt].lda = h; g[cnt].B = wt0; g[cnt].b_mn = 1; g[cnt].ldb = s.d; g[cnt].M = s.B; g[cnt].N = s.d; g[cnt].K = h;
            g[cnt].out = dt; g[cnt].out_kind = okT; g[cnt].ldo = s.d;
            g[cnt].pdl = phase == VPR_BWD_ALL ? 2 : 0;
            ++cnt;
        }
        if (cnt > 0) VPR_TRY(gemm_run_group(g, cnt, st));
    }
    return 0;
}

}  // namespace vpr

using namespace vpr;

extern "C" {

size_t vpr_salad_save_bytes(const VprSaladShape* s, int dtype) { return save_plan(*s, dtype).total; }
size_t vpr_salad_workspace_bytes(const VprSaladShape* s, int dtype) { return work_plan(*s, dtype).total; }

int vpr_salad_fwd(const VprSaladParams* p, const void* x, const void* t, void* out, void* save, void* workspace,
                  size_t workspace_bytes, const VprSaladShape* s, int dtype, float dropout_p, uint64_t seed,
                  const uint64_t* seed_offset_dev, cudaStream_t stream) {
    VPR_TRY(::vpr::check_device());
    (void)workspace; (void)workspace_bytes;
    VPR_REQUIRE(p && x && t && out && save && s, "null argument");
    VPR_TRY(check_shape(*s, dtype));
    VPR_REQUIRE(dropout_p >= 0.f && dropout_p < 1.f, "dropout_p out of range");
    if (dtype == VPR_BF16)
        return salad_fwd_t<bf16>(*p, (const bf16*)x, (const bf16*)t, (bf16*)out, (char*)save, *s, dtype, dropout_p, seed, seed_offset_dev, stream);
    return salad_fwd_t<float>(*p, (const float*)x, (const float*)t, (float*)out, (char*)save, *s, dtype, dropout_p, seed, seed_offset_dev, stream);
}

int vpr_salad_bwd(const VprSaladParams* p, const void* x, const void* t, const void* d_out, const void* save,
                  void* workspace, size_t workspace_bytes, VprSaladGrads* grads, void* dx, void* dt,
                  const VprSaladShape* s, int dtype, float dropout_p, int phase, int reserve_sms, cudaStream_t stream) {
    VPR_TRY(::vpr::check_device());
    VPR_REQUIRE(p && x && t && d_out && save && workspace && grads && s, "null argument");
    VPR_REQUIRE(phase >= VPR_BWD_ALL && phase <= VPR_BWD_INPUTS && reserve_sms >= 0, "bad phase / reserve_sms");
    VPR_TRY(check_shape(*s, dtype));
    VPR_REQUIRE(workspace_bytes >= work_plan(*s, dtype).total, "workspace too small: %zu < %zu", workspace_bytes,
                work_plan(*s, dtype).total);
    if (dtype == VPR_BF16)
        return salad_bwd_t<bf16>(*p, (const bf16*)x, (const bf16*)t, (const bf16*)d_out, (const char*)save,
                                 (char*)workspace, *grads, (bf16*)dx, (bf16*)dt, *s, dtype, dropout_p, phase, reserve_sms, stream);
    return salad_bwd_t<float>(*p, (const float*)x, (const float*)t, (const float*)d_out, (const char*)save,
                              (char*)workspace, *grads, (float*)dx, (float*)dt, *s, dtype, dropout_p, phase, reserve_sms, stream);
}

int vpr_counter_advance(uint64_t* counter, cudaStream_t stream) {
    VPR_TRY(::vpr::check_device());
    VPR_REQUIRE(counter != nullptr, "null counter");
    VPR_CUDA_OK(launch_ex(counter_advance_kernel, dim3(1), dim3(1), (size_t)0, stream, true, 1,
                          reinterpret_cast<unsigned long long*>(counter)));
    return launch_ok("counter_advance_kernel");
}

int vpr_dropout_mask(uint64_t seed, float p, int64_t rows, int64_t cols, uint8_t* out, cudaStream_t stream) {
    VPR_TRY(::vpr::check_device());
    VPR_REQUIRE(out && rows > 0 && cols > 0, "bad argument");
    dropout_mask_kernel<<<num_sms() * 4, 256, 0, stream>>>(seed, p > 0.f ? dropout_threshold(p) : 0u, rows * cols, out);
    return launch_ok("dropout_mask_kernel");
}

int vpr_sinkhorn_fwd(const float* S, const float* dust_bin, void* P, int p_dtype, float* U, float* V, int B, int n,
                     int m, int iters, cudaStream_t stream) {
    VPR_TRY(::vpr::check_device());
    VPR_REQUIRE(S && dust_bin && P && U && V && B > 0, "bad argument");
    return sinkhorn_fwd(S, dust_bin, P, p_dtype, U, V, B, n, m, iters, stream);
}

int vpr_sinkhorn_bwd(const float* S, const float* dust_bin, const float* U, const float* V, const float* dP, void* dS,
                     int ds_dtype, float* d_bias, int B, int n, int m, int iters, cudaStream_t stream) {
    VPR_TRY(::vpr::check_device());
    VPR_REQUIRE(S && dust_bin && U && V && dP && dS && B > 0, "bad argument");
    return sinkhorn_bwd(S, dust_bin, U, V, dP, dS, ds_dtype, d_bias, B, n, m, iters, stream);
}

}  // extern "C"
