// Descriptor assembly (intra-cluster L2 norm over l, token L2 norm, concat, global L2 norm) and its
// backward, one CTA per image, for aggregates of up to 8192 elements (upstream: l*m = 128*64).
// Everything an image needs is read once into registers with 16-byte loads; every norm, and in the
// backward every projection term (y.dy, ahat_c . dz_c, that . dz_t), is a closed form of ONE round of
// per-cluster sums (sum a^2, sum a*dy) and two token scalars, so there is a single block-wide
// reduction instead of one per term.  The backward also writes dAGG transposed ([m][l]) through a
// shared-memory tile so that the dP GEMM can consume it K-major in the same launch as the dF GEMM.
#pragma once
#include "common.cuh"

namespace vpr {

#define VPR_NORM_EPS 1e-12f

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 ld4(const bf16* p) {
    const uint2 u = *reinterpret_cast<const uint2*>(p);
    return make_float4(__uint_as_float(u.x << 16), __uint_as_float(u.x & 0xFFFF0000u), __uint_as_float(u.y << 16),
                       __uint_as_float(u.y & 0xFFFF0000u));
}
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ void st4(bf16* p, float4 v) {
    __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
    uint2 u;
    u.x = *reinterpret_cast<uint32_t*>(&lo);
    u.y = *reinterpret_cast<uint32_t*>(&hi);
    *reinterpret_cast<uint2*>(p) = u;
}
__device__ __forceinline__ void ld8(const float* p, float (&w)[8]) {
    const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w; w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
}
__device__ __forceinline__ void ld8(const bf16* p, float (&w)[8]) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
    w[0] = __uint_as_float(u.x << 16); w[1] = __uint_as_float(u.x & 0xFFFF0000u);
    w[2] = __uint_as_float(u.y << 16); w[3] = __uint_as_float(u.y & 0xFFFF0000u);
    w[4] = __uint_as_float(u.z << 16); w[5] = __uint_as_float(u.z & 0xFFFF0000u);
    w[6] = __uint_as_float(u.w << 16); w[7] = __uint_as_float(u.w & 0xFFFF0000u);
}
__device__ __forceinline__ float ldf1(const float* p) { return *p; }
__device__ __forceinline__ float ldf1(const bf16* p) { return __bfloat162float(*p); }
__device__ __forceinline__ void stf1(float* p, float v) { *p = v; }
__device__ __forceinline__ void stf1(bf16* p, float v) { *p = __float2bfloat16_rn(v); }

static constexpr int FIN_MAXV = 8;      // 16-byte vectors of the aggregate per thread (256 threads)

static inline bool finalize_fast_ok(int l, int m, int g) {
    const long long e = (long long)l * m;
    return (m == 16 || m == 32 || m == 64 || m == 128) && e % 1024 == 0 && e / 1024 <= FIN_MAXV && g % 4 == 0 && l % 4 == 0 &&
           g <= 1024;
}
template <typename T>
static inline size_t finalize_bwd_fast_smem(int l, int m) { return (size_t)m * (l + 4) * sizeof(T); }

// lanes that hold the same 4-cluster chunk (same tid % (m/4)) add up their partial sums
__device__ __forceinline__ void fin_share(float4& x, int cpr) {
    for (int o = cpr; o < 32; o <<= 1) {
        x.x += __shfl_xor_sync(0xffffffffu, x.x, o); x.y += __shfl_xor_sync(0xffffffffu, x.y, o);
        x.z += __shfl_xor_sync(0xffffffffu, x.z, o); x.w += __shfl_xor_sync(0xffffffffu, x.w, o);
    }
}

// wt1 != nullptr: the token MLP's second layer TF[b, :] = HT[b, :] Wt1^T + bt1 is computed here (one image per
// CTA, g dot products of length hid: a warp per output, lanes along k) and written to TF for the backward.
template <typename T>
__global__ void __launch_bounds__(256) finalize_fwd_fast(const float* __restrict__ AGG, float* __restrict__ TF,
                                                         T* __restrict__ out, int l, int m, int g,
                                                         const T* __restrict__ HT, const T* __restrict__ wt1,
                                                         const float* __restrict__ bt1, int hid) {
    __shared__ __align__(16) float red[8][128];
    __shared__ __align__(16) float inv_c[128];
    __shared__ float wred[8];
    __shared__ float sc[4];
    __shared__ __align__(16) float sht[1024];                  // HT row (hid <= 1024)
    __shared__ float stf[1024];                                // TF row (g <= 1024)
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float* A = AGG + (size_t)b * l * m;
    const float* t = TF + (size_t)b * g;
    if (wt1 != nullptr) {
        for (int i = tid; i < hid; i += 256) sht[i] = ldf1(HT + (size_t)b * hid + i);
        __syncthreads();
        for (int o = warp; o < g; o += 8) {
            float acc = 0.f;
            for (int k8 = lane * 8; k8 < hid; k8 += 256) {
                float w[8];
                ld8(wt1 + (size_t)o * hid + k8, w);
#pragma unroll
                for (int i = 0; i < 8; ++i) acc = fmaf(w[i], sht[k8 + i], acc);
            }
            acc = warp_sum(acc);
            if (lane == 0) {
                const float v = acc + __ldg(bt1 + o);
                stf[o] = v;
                TF[(size_t)b * g + o] = v;
            }
        }
        __syncthreads();
        t = stf;
    }
    T* o = out + (size_t)b * (g + l * m);
    const int nv = (l * m) >> 10, cpr = m >> 2, c0 = (tid % cpr) * 4;
    float4 a[FIN_MAXV];
    float4 ss = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int i = 0; i < FIN_MAXV; ++i)
        if (i < nv) {
            a[i] = ld4(A + (size_t)(i * 256 + tid) * 4);
            ss.x = fmaf(a[i].x, a[i].x, ss.x); ss.y = fmaf(a[i].y, a[i].y, ss.y);
            ss.z = fmaf(a[i].z, a[i].z, ss.z); ss.w = fmaf(a[i].w, a[i].w, ss.w);
        }
    fin_share(ss, cpr);
    if (lane < cpr) *reinterpret_cast<float4*>(&red[warp][c0]) = ss;
    float ts = 0.f;
    for (int i = tid; i < g; i += 256) { const float v = t[i]; ts = fmaf(v, v, ts); }
    ts = warp_sum(ts);
    if (lane == 0) wred[warp] = ts;
    __syncthreads();
    if (warp == 0) {
        float zz = 0.f;
        for (int c = lane; c < m; c += 32) {
            float tot = 0.f;
#pragma unroll
            for (int w = 0; w < 8; ++w) tot += red[w][c];
            const float nc = fmaxf(sqrtf(tot), VPR_NORM_EPS);
            inv_c[c] = 1.f / nc;
            zz += tot / (nc * nc);
        }
        zz = warp_sum(zz);
        float tsum = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) tsum += wred[w];
        const float nt = fmaxf(sqrtf(tsum), VPR_NORM_EPS);
        zz += tsum / (nt * nt);
        const float invN = 1.f / fmaxf(sqrtf(zz), VPR_NORM_EPS);
        if (lane == 0) { sc[0] = invN; sc[1] = invN / nt; }
    }
    __syncthreads();
    const float invN = sc[0], invt = sc[1];
    float4 ic = *reinterpret_cast<const float4*>(&inv_c[c0]);
    ic.x *= invN; ic.y *= invN; ic.z *= invN; ic.w *= invN;
#pragma unroll
    for (int i = 0; i < FIN_MAXV; ++i)
        if (i < nv) st4(o + g + (size_t)(i * 256 + tid) * 4, make_float4(a[i].x * ic.x, a[i].y * ic.y, a[i].z * ic.z, a[i].w * ic.w));
    for (int i = tid; i < g; i += 256) stf1(o + i, t[i] * invt);
}

template <typename T>
__global__ void __launch_bounds__(256) finalize_bwd_fast(const float* __restrict__ AGG, const float* __restrict__ TF,
                                                         const T* __restrict__ d_out, T* __restrict__ dAGG,
                                                         T* __restrict__ dAGGT, T* __restrict__ dTF,
                                                         float* __restrict__ g_bias_tok, int l, int m, int g,
                                                         const T* __restrict__ wt1, const T* __restrict__ HT,
                                                         T* __restrict__ dHT, float* __restrict__ g_bias_tok0, int hid) {
    extern __shared__ __align__(16) uint8_t fin_smem[];
    T* tileT = reinterpret_cast<T*>(fin_smem);                 // [m][l + 4]
    __shared__ __align__(16) float red_s[8][128];
    __shared__ __align__(16) float red_q[8][128];
    __shared__ __align__(16) float inv_c[128];
    __shared__ __align__(16) float dot_c[128];
    __shared__ float wred[2][8];
    __shared__ float sc[4];
    __shared__ float sdt[1024];                                // this image's dTF row (g <= 1024)
    __shared__ float spart[4][512];                            // partial dHT rows of the four k-quarters
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float* A = AGG + (size_t)b * l * m;
    const float* t = TF + (size_t)b * g;
    const T* dy = d_out + (size_t)b * (g + l * m);
    const int nv = (l * m) >> 10, cpr = m >> 2, c0 = (tid % cpr) * 4, lp = l + 4;
    float4 a[FIN_MAXV], d[FIN_MAXV];
    float4 ss = make_float4(0.f, 0.f, 0.f, 0.f), qq = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int i = 0; i < FIN_MAXV; ++i)
        if (i < nv) {
            a[i] = ld4(A + (size_t)(i * 256 + tid) * 4);
            d[i] = ld4(dy + g + (size_t)(i * 256 + tid) * 4);
            ss.x = fmaf(a[i].x, a[i].x, ss.x); ss.y = fmaf(a[i].y, a[i].y, ss.y);
            ss.z = fmaf(a[i].z, a[i].z, ss.z); ss.w = fmaf(a[i].w, a[i].w, ss.w);
            qq.x = fmaf(a[i].x, d[i].x, qq.x); qq.y = fmaf(a[i].y, d[i].y, qq.y);
            qq.z = fmaf(a[i].z, d[i].z, qq.z); qq.w = fmaf(a[i].w, d[i].w, qq.w);
        }
    fin_share(ss, cpr);
    fin_share(qq, cpr);
    if (lane < cpr) {
        *reinterpret_cast<float4*>(&red_s[warp][c0]) = ss;
        *reinterpret_cast<float4*>(&red_q[warp][c0]) = qq;
    }
    float ts = 0.f, tp = 0.f;
    for (int i = tid; i < g; i += 256) { const float v = t[i]; ts = fmaf(v, v, ts); tp = fmaf(v, ldf1(dy + i), tp); }
    ts = warp_sum(ts);
    tp = warp_sum(tp);
    if (lane == 0) { wred[0][warp] = ts; wred[1][warp] = tp; }
    __syncthreads();
    if (warp == 0) {
        float zz = 0.f, sq = 0.f;
        float ssc[4], qc[4], icv[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int c = lane + 32 * k;
            ssc[k] = 0.f; qc[k] = 0.f; icv[k] = 0.f;
            if (c < m) {
#pragma unroll
                for (int w = 0; w < 8; ++w) { ssc[k] += red_s[w][c]; qc[k] += red_q[w][c]; }
                const float nc = fmaxf(sqrtf(ssc[k]), VPR_NORM_EPS);
                icv[k] = 1.f / nc;
                inv_c[c] = icv[k];
                zz += ssc[k] * icv[k] * icv[k];
                sq += icv[k] * qc[k];
            }
        }
        zz = warp_sum(zz);
        sq = warp_sum(sq);
        float tsum = 0.f, psum = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) { tsum += wred[0][w]; psum += wred[1][w]; }
        const float invt = 1.f / fmaxf(sqrtf(tsum), VPR_NORM_EPS);
        zz += tsum * invt * invt;
        const float invN = 1.f / fmaxf(sqrtf(zz), VPR_NORM_EPS);
        const float s = invN * (psum * invt + sq);                          // y . dy
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int c = lane + 32 * k;                                    // ahat_c . dz_c
            if (c < m) dot_c[c] = invN * (icv[k] * qc[k] - ssc[k] * icv[k] * icv[k] * invN * s);
        }
        if (lane == 0) {
            sc[0] = invN; sc[1] = invt; sc[2] = s;
            sc[3] = invN * (psum * invt - tsum * invt * invt * invN * s);   // that . dz_t
        }
    }
    __syncthreads();
    const float invN = sc[0], invt = sc[1], s = sc[2], dt_dot = sc[3];
    const float4 ic = *reinterpret_cast<const float4*>(&inv_c[c0]);
    const float4 dc = *reinterpret_cast<const float4*>(&dot_c[c0]);
    const float ns = invN * s;
    T* dA = dAGG + (size_t)b * l * m;
#pragma unroll
    for (int i = 0; i < FIN_MAXV; ++i)
        if (i < nv) {
            const int v = i * 256 + tid;
            const int row = (v * 4) / m;
            float4 r;
            { const float ah = a[i].x * ic.x; r.x = ((d[i].x - ah * ns) * invN - ah * dc.x) * ic.x; }
            { const float ah = a[i].y * ic.y; r.y = ((d[i].y - ah * ns) * invN - ah * dc.y) * ic.y; }
            { const float ah = a[i].z * ic.z; r.z = ((d[i].z - ah * ns) * invN - ah * dc.z) * ic.z; }
            { const float ah = a[i].w * ic.w; r.w = ((d[i].w - ah * ns) * invN - ah * dc.w) * ic.w; }
            st4(dA + (size_t)v * 4, r);
            stf1(tileT + (size_t)(c0 + 0) * lp + row, r.x);
            stf1(tileT + (size_t)(c0 + 1) * lp + row, r.y);
            stf1(tileT + (size_t)(c0 + 2) * lp + row, r.z);
            stf1(tileT + (size_t)(c0 + 3) * lp + row, r.w);
        }
    T* dT = dTF + (size_t)b * g;
    for (int i = tid; i < g; i += 256) {
        const float th = t[i] * invt;
        const float dz = (ldf1(dy + i) - th * ns) * invN;
        const float dti = (dz - th * dt_dot) * invt;
        stf1(dT + i, dti);
        sdt[i] = dti;
        if (g_bias_tok != nullptr) atomicAdd(g_bias_tok + i, dti);
    }
    __syncthreads();
    T* dAT = dAGGT + (size_t)b * l * m;
    const int lq = l >> 2;                                      // 4-element vectors per transposed row
    for (int v = tid; v < m * lq; v += 256) {
        const int c = v / lq, l4 = (v - c * lq) * 4;
        const T* src = tileT + (size_t)c * lp + l4;
        if constexpr (sizeof(T) == 2) *reinterpret_cast<uint2*>(dAT + (size_t)c * l + l4) = *reinterpret_cast<const uint2*>(src);
        else *reinterpret_cast<float4*>(dAT + (size_t)c * l + l4) = *reinterpret_cast<const float4*>(src);
    }
    // token MLP, hidden-layer gradient of this image: dHT[b, :] = (dTF[b, :] Wt1) * (HT[b, :] > 0), and its
    // contribution to token_features.0.bias -- 2 * g * hid MACs, not worth a launch of its own.  Thread =
    // (8 adjacent hidden units, one quarter of the g rows): 16-byte weight loads, 16 in flight per thread.
    if (dHT != nullptr) {                                      // requires hid == 512 (host checks)
        const int c8 = (tid & 63) * 8, ksl = tid >> 6;
        float acc[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = 0.f;
#pragma unroll 16
        for (int k = ksl; k < g; k += 4) {
            const float dk = sdt[k];
            float w[8];
            ld8(wt1 + (size_t)k * hid + c8, w);
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] = fmaf(dk, w[i], acc[i]);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) spart[ksl][c8 + i] = acc[i];
        __syncthreads();
        const int c2 = tid * 2;
        float a0 = spart[0][c2] + spart[1][c2] + spart[2][c2] + spart[3][c2];
        float a1 = spart[0][c2 + 1] + spart[1][c2 + 1] + spart[2][c2 + 1] + spart[3][c2 + 1];
        const T* hp = HT + (size_t)b * hid + c2;
        a0 = ldf1(hp) > 0.f ? a0 : 0.f;
        a1 = ldf1(hp + 1) > 0.f ? a1 : 0.f;
        T* op = dHT + (size_t)b * hid + c2;
        stf1(op, a0);
        stf1(op + 1, a1);
        if (g_bias_tok0 != nullptr) {                          // column sums of the stored (rounded) values
            atomicAdd(g_bias_tok0 + c2, ldf1(op));
            atomicAdd(g_bias_tok0 + c2 + 1, ldf1(op + 1));
        }
    }
}

}  // namespace vpr
