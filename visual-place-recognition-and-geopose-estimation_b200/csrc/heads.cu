// Geopose heads + losses, forward and backward in one call (SURVEY.md section 8 rows a8, a9,
// Appendix B).
//
//   W_aug [hid_pad, D] = [ regressor.0.weight ; angle head rows ; 0 ]   (run dtype, packed once per weight update)
//   zacc  [B, hid_pad] = desc W_aug^T            one split-K tcgen05 GEMM: regressor pre-activation AND the angle
//                                                head's raw outputs (columns hid, hid+1)
//   tail                                         ReLU, Dropout, Linear(hid,2), every loss variant and its gradient,
//                                                8 rows per CTA; writes dZ_aug [B, hid_pad] = [dZ | d sin,cos | 0]
//   d_desc = dZ_aug W_aug                        one GEMM (K = hid_pad): regressor and angle head together
//   dW0    = dZ^T desc                           one GEMM
//   dWa    = d(sin,cos)^T desc                   small kernel (2 x B x D)
// With Dropout in front of the angle head (dino_v2_gemini.py:104,111) the head's mask differs from the regressor's
// input, so its three products leave the shared GEMMs and run in two small masked kernels instead.
#include "common.cuh"
#include "gemm.h"
#include "../../include/vpr_b200.h"

#include <math.h>
#include <stdlib.h>

namespace vpr {

__device__ __forceinline__ float hldf(const float* p) { return *p; }
__device__ __forceinline__ float hldf(const bf16* p) { return __bfloat162float(*p); }
__device__ __forceinline__ void hstf(float* p, float v) { *p = v; }
__device__ __forceinline__ void hstf(bf16* p, float v) { *p = __float2bfloat16_rn(v); }

__device__ __forceinline__ void load8(const float* p, float (&v)[8]) {
    const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void load8(const bf16* p, float (&v)[8]) {
    const uint4 u = *reinterpret_cast<const uint4*>(p);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        v[2 * i] = __uint_as_float(w[i] << 16);
        v[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u);
    }
}
__device__ __forceinline__ void store8(float* p, const float (&v)[8]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void store8(bf16* p, const float (&v)[8]) {
    uint4 u;
    uint32_t* w = reinterpret_cast<uint32_t*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
        w[i] = *reinterpret_cast<uint32_t*>(&h);
    }
    *reinterpret_cast<uint4*>(p) = u;
}

// keep flags of 8 consecutive elements starting at idx0 (idx0 % 8 == 0: all in one 32-element run)
__device__ __forceinline__ void keep8(unsigned long long seed, unsigned long long idx0, uint32_t thr, float scale,
                                      float (&k)[8]) {
    uint32_t w = dropout_mix(seed, idx0 >> 5);
    const int s = (int)(idx0 & 31);
    for (int q = 0; q < s; ++q) w = w * DROP_LCG_A + DROP_LCG_C;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        k[q] = w >= thr ? scale : 0.f;
        w = w * DROP_LCG_A + DROP_LCG_C;
    }
}

#define RAD2DEG 57.29577951308232f
#define DEG2RAD 0.017453292519943295f
// the angle head's input dropout draws from its own stream of the counter-based generator
#define HEAD_DROP_SEED_OFFSET 0x9E3779B97F4A7C15ull

// python-style remainder, result in [0, 360)
__device__ __forceinline__ float mod360(float x) {
    float r = fmodf(x, 360.f);
    if (r < 0.f) r += 360.f;
    return r;
}

// W_aug [hid_pad, D]: rows [0,hid) = regressor.0.weight, then the angle head's rows, rest 0; four elements per thread
struct HeadsPackArgs { const float* w0; const float* wa; void* waug; long long n0_4, na_4, total_4; };

template <typename T>
__global__ void __launch_bounds__(256) pack_heads_kernel(const HeadsPackArgs a) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < a.total_4;
         i += (long long)gridDim.x * blockDim.x) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i < a.n0_4) v = __ldg(reinterpret_cast<const float4*>(a.w0) + i);
        else if (i < a.n0_4 + a.na_4) v = __ldg(reinterpret_cast<const float4*>(a.wa) + (i - a.n0_4));
        if (sizeof(T) == 2) {
            __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
            uint2 pk;
            pk.x = *reinterpret_cast<uint32_t*>(&lo);
            pk.y = *reinterpret_cast<uint32_t*>(&hi);
            reinterpret_cast<uint2*>(a.waug)[i] = pk;
        } else {
            reinterpret_cast<float4*>(a.waug)[i] = v;
        }
    }
}

// Dropout in front of the angle head: sc_raw[b][k] = sum_d keep[b,d] * desc[b,d] * W_aug[hid+k][d]
template <typename T>
__global__ void __launch_bounds__(256) angle_fwd_masked_kernel(const T* __restrict__ desc, const T* __restrict__ wa,
                                                               float* __restrict__ sc_raw, int D, uint32_t thr,
                                                               float scale, unsigned long long seed,
                                                               const unsigned long long* __restrict__ seed_dev) {
    __shared__ float red[2][8];
    pdl_wait();
    pdl_launch_dependents();
    if (seed_dev != nullptr) seed += *seed_dev;
    seed += HEAD_DROP_SEED_OFFSET;
    const int b = blockIdx.x;
    const T* x = desc + (size_t)b * D;
    float a0 = 0.f, a1 = 0.f;
    for (int d0 = threadIdx.x * 8; d0 < D; d0 += 256 * 8) {
        float v[8], w0[8], w1[8], k[8];
        load8(x + d0, v);
        load8(wa + d0, w0);
        load8(wa + D + d0, w1);
        keep8(seed, (unsigned long long)b * D + d0, thr, scale, k);
#pragma unroll
        for (int q = 0; q < 8; ++q) { a0 += v[q] * k[q] * w0[q]; a1 += v[q] * k[q] * w1[q]; }
    }
    a0 = warp_sum(a0); a1 = warp_sum(a1);
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = a0; red[1][threadIdx.x >> 5] = a1; }
    __syncthreads();
    if (threadIdx.x < 2) {
        float t = 0.f;
        for (int i = 0; i < 8; ++i) t += red[threadIdx.x][i];
        sc_raw[b * 2 + threadIdx.x] = t;
    }
}

struct TailArgs {
    const float* zacc;      // [B, hid_pad] pre-activation (no bias); columns hid, hid+1 = raw angle-head outputs
    const float* sc_raw;    // [B, 2] raw angle-head outputs when the head has its own dropout, else null
    const float* b0; const float* w1; const float* b1; const float* ba;
    const float* latlon_target; const float* angle_target_deg;
    float* latlon_pred; float* sincos_pred; float* losses;
    float* g_b0; float* g_w1; float* g_b1; float* g_ba;
    void* dzaug;            // [B, hid_pad] run dtype
    int B, hid, hid_pad, latlon_loss, angle_loss, global_B, forward_only;
    float huber_delta, loss_scale;
    uint32_t drop_thr; float drop_scale; unsigned long long seed; const unsigned long long* seed_dev;
};

static constexpr int TAIL_ROWS = 8, TAIL_THREADS = 256, TAIL_WARPS = TAIL_THREADS / 32;

// 8 descriptor rows per CTA: the per-column gradient sums (regressor.2.weight, regressor.0.bias) are accumulated over
// the CTA's rows in registers, so each column costs three atomics per CTA instead of three per row.  Every pass first
// issues all of its loads (8 rows per column) and only then computes / stores, so the rows' latencies overlap.
template <typename T>
__global__ void __launch_bounds__(TAIL_THREADS) heads_tail_kernel(const TailArgs a) {
    __shared__ float part[TAIL_WARPS][2 * TAIL_ROWS];
    __shared__ float dpred_s[TAIL_ROWS][2];
    __shared__ float dsc_s[TAIL_ROWS][2];
    pdl_wait();
    pdl_launch_dependents();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int row0 = blockIdx.x * TAIL_ROWS;
    const int rows = min(TAIL_ROWS, a.B - row0);
    const unsigned long long seed = a.seed + (a.seed_dev != nullptr ? *a.seed_dev : 0ull);
    const float* __restrict__ zacc = a.zacc;
    const float* __restrict__ b0 = a.b0;
    const float* __restrict__ w1p = a.w1;
    // ---- z = dropout(relu(zacc + b0)); pred = W1 z + b1 ----
    float p[2 * TAIL_ROWS];
#pragma unroll
    for (int r = 0; r < 2 * TAIL_ROWS; ++r) p[r] = 0.f;
    for (int i = tid; i < a.hid; i += TAIL_THREADS) {
        const float bi = __ldg(b0 + i), w0 = __ldg(w1p + i), w1 = __ldg(w1p + a.hid + i);
        float zr[TAIL_ROWS];
#pragma unroll
        for (int r = 0; r < TAIL_ROWS; ++r) zr[r] = r < rows ? __ldcg(zacc + (size_t)(row0 + r) * a.hid_pad + i) : 0.f;
#pragma unroll
        for (int r = 0; r < TAIL_ROWS; ++r) {
            float z = fmaxf(zr[r] + bi, 0.f);
            if (a.drop_thr != 0u && r < rows)
                z = dropout_keep(seed, (unsigned long long)(row0 + r) * a.hid + i, a.drop_thr) ? z * a.drop_scale : 0.f;
            if (r >= rows) z = 0.f;
            p[2 * r] += z * w0;
            p[2 * r + 1] += z * w1;
        }
    }
#pragma unroll
    for (int r = 0; r < 2 * TAIL_ROWS; ++r) {
        const float s = warp_sum(p[r]);
        if (lane == 0) part[warp][r] = s;
    }
    __syncthreads();
    if (warp == 0) {
        // lane r < rows owns row row0 + r
        const int r = lane, b = row0 + lane;
        float loss = 0.f, aloss = 0.f, gb1_0 = 0.f, gb1_1 = 0.f, gba_0 = 0.f, gba_1 = 0.f;
        if (r < rows) {
            float pred[2] = {a.b1[0], a.b1[1]};
#pragma unroll
            for (int w = 0; w < TAIL_WARPS; ++w) { pred[0] += part[w][2 * r]; pred[1] += part[w][2 * r + 1]; }
            a.latlon_pred[b * 2 + 0] = pred[0];
            a.latlon_pred[b * 2 + 1] = pred[1];
            const float inv2B = 1.f / (2.f * (float)a.global_B);
            float gk[2];
            for (int k = 0; k < 2; ++k) {
                const float tv = a.latlon_target[b * 2 + k];
                float g;
                if (a.latlon_loss == VPR_LATLON_EXTERNAL) {          // tv = dL/dpred handed in by the caller
                    g = tv * a.loss_scale;
                } else {
                    const float e = pred[k] - tv;
                    if (a.latlon_loss == VPR_LATLON_MSE) { loss += e * e; g = 2.f * e; }
                    else if (fabsf(e) <= a.huber_delta) { loss += 0.5f * e * e; g = e; }
                    else { loss += a.huber_delta * (fabsf(e) - 0.5f * a.huber_delta); g = a.huber_delta * (e > 0.f ? 1.f : -1.f); }
                    g *= inv2B * a.loss_scale;
                }
                gk[k] = g;
                dpred_s[r][k] = g;
            }
            loss *= inv2B;
            gb1_0 = gk[0]; gb1_1 = gk[1];
            // ---- angle head ----
            float ds = 0.f, dc = 0.f;
            if (a.angle_loss != VPR_ANGLE_NONE) {
                const float raw0 = a.sc_raw != nullptr ? a.sc_raw[b * 2 + 0] : __ldcg(zacc + (size_t)b * a.hid_pad + a.hid);
                const float raw1 = a.sc_raw != nullptr ? a.sc_raw[b * 2 + 1] : __ldcg(zacc + (size_t)b * a.hid_pad + a.hid + 1);
                const float invB = 1.f / (float)a.global_B;
                if (a.angle_loss == VPR_ANGLE_SCALAR_MAAE) {
                    // dino_angle_finetuning.py:41-52: pred = head(x) % 360; d = |pred - target| % 360; min(d, 360 - d)
                    const float rawp = raw0 + a.ba[0];
                    const float pdeg = mod360(rawp);
                    a.sincos_pred[b * 2 + 0] = pdeg;
                    a.sincos_pred[b * 2 + 1] = rawp;
                    const float diff = pdeg - a.angle_target_deg[b];
                    const float d = fmodf(fabsf(diff), 360.f);
                    float gd = diff > 0.f ? 1.f : (diff < 0.f ? -1.f : 0.f);
                    if (d <= 360.f - d) aloss = d; else { aloss = 360.f - d; gd = -gd; }
                    ds = gd;
                } else {
                    const float s = raw0 + a.ba[0], c = raw1 + a.ba[1];
                    a.sincos_pred[b * 2 + 0] = s;
                    a.sincos_pred[b * 2 + 1] = c;
                    const float tdeg = mod360(a.angle_target_deg[b]);
                    if (a.angle_loss == VPR_ANGLE_ATAN2_MAAE) {
                        const float deg = mod360(atan2f(s, c) * RAD2DEG);
                        const float diff = deg - tdeg;
                        const float d = fabsf(diff);
                        float gd = diff > 0.f ? 1.f : (diff < 0.f ? -1.f : 0.f);
                        if (d <= 360.f - d) aloss = d; else { aloss = 360.f - d; gd = -gd; }
                        const float r2 = s * s + c * c;
                        ds = gd * RAD2DEG * c / r2;
                        dc = -gd * RAD2DEG * s / r2;
                    } else if (a.angle_loss == VPR_ANGLE_ACOS) {
                        const float ts = sinf(tdeg * DEG2RAD), tc = cosf(tdeg * DEG2RAD);
                        const float nrm = sqrtf(s * s + c * c);
                        const float den = fmaxf(nrm, 1e-6f);
                        const float ps = s / den, pc = c / den;
                        const float dot = ps * ts + pc * tc;
                        const float cl = fminf(fmaxf(dot, -0.999999f), 0.999999f);
                        aloss = acosf(cl) * RAD2DEG;
                        float gdot = 0.f;
                        if (dot >= -0.999999f && dot <= 0.999999f) gdot = -RAD2DEG / sqrtf(1.f - cl * cl);
                        const float gps = gdot * ts, gpc = gdot * tc;
                        if (nrm > 1e-6f) {
                            const float pg = ps * gps + pc * gpc;
                            ds = (gps - ps * pg) / den;
                            dc = (gpc - pc * pg) / den;
                        } else { ds = gps / den; dc = gpc / den; }
                    } else {   // MSE on [sin, cos], mean over B*2
                        const float ts = sinf(tdeg * DEG2RAD), tc = cosf(tdeg * DEG2RAD);
                        aloss = 0.5f * ((s - ts) * (s - ts) + (c - tc) * (c - tc));
                        ds = (s - ts);
                        dc = (c - tc);
                    }
                }
                aloss *= invB;
                ds *= invB * a.loss_scale;
                dc *= invB * a.loss_scale;
                gba_0 = ds; gba_1 = dc;
            }
            dsc_s[r][0] = ds;
            dsc_s[r][1] = dc;
        }
        loss = warp_sum(loss); aloss = warp_sum(aloss);
        gb1_0 = warp_sum(gb1_0); gb1_1 = warp_sum(gb1_1); gba_0 = warp_sum(gba_0); gba_1 = warp_sum(gba_1);
        if (lane == 0) {
            atomicAdd(a.losses + 0, loss);
            if (a.angle_loss != VPR_ANGLE_NONE) atomicAdd(a.losses + 1, aloss);
            atomicAdd(a.losses + 2, loss + aloss);
            if (!a.forward_only) {
                atomicAdd(a.g_b1 + 0, gb1_0);
                atomicAdd(a.g_b1 + 1, gb1_1);
                if (a.angle_loss != VPR_ANGLE_NONE) {
                    atomicAdd(a.g_ba + 0, gba_0);
                    if (a.angle_loss != VPR_ANGLE_SCALAR_MAAE) atomicAdd(a.g_ba + 1, gba_1);
                }
            }
        }
    }
    if (a.forward_only) return;
    __syncthreads();
    // ---- dZ_aug rows, regressor.2.weight / regressor.0.bias gradients ----
    T* __restrict__ dz = reinterpret_cast<T*>(a.dzaug);
    float g0s[TAIL_ROWS], g1s[TAIL_ROWS];
#pragma unroll
    for (int r = 0; r < TAIL_ROWS; ++r) { g0s[r] = r < rows ? dpred_s[r][0] : 0.f; g1s[r] = r < rows ? dpred_s[r][1] : 0.f; }
    for (int i = tid; i < a.hid_pad; i += TAIL_THREADS) {
        if (i < a.hid) {
            const float bi = __ldg(b0 + i), w0 = __ldg(w1p + i), w1 = __ldg(w1p + a.hid + i);
            float zr[TAIL_ROWS];
#pragma unroll
            for (int r = 0; r < TAIL_ROWS; ++r) zr[r] = r < rows ? __ldcg(zacc + (size_t)(row0 + r) * a.hid_pad + i) : 0.f;
            float gw0 = 0.f, gw1 = 0.f, gb = 0.f;
#pragma unroll
            for (int r = 0; r < TAIL_ROWS; ++r) {
                if (r < rows) {
                    const float z = fmaxf(zr[r] + bi, 0.f);
                    float keep = 1.f;
                    if (a.drop_thr != 0u)
                        keep = dropout_keep(seed, (unsigned long long)(row0 + r) * a.hid + i, a.drop_thr) ? a.drop_scale : 0.f;
                    const float zd = z * keep;
                    gw0 += g0s[r] * zd;
                    gw1 += g1s[r] * zd;
                    const float d = z > 0.f ? (g0s[r] * w0 + g1s[r] * w1) * keep : 0.f;
                    hstf(dz + (size_t)(row0 + r) * a.hid_pad + i, d);
                    gb += d;
                }
            }
            atomicAdd(a.g_w1 + i, gw0);
            atomicAdd(a.g_w1 + a.hid + i, gw1);
            atomicAdd(a.g_b0 + i, gb);
        } else {
            for (int r = 0; r < rows; ++r)
                hstf(dz + (size_t)(row0 + r) * a.hid_pad + i, i - a.hid < 2 ? dsc_s[r][i - a.hid] : 0.f);
        }
    }
}

// Angle-head weight gradient  dWa[k][d] += sum_b dsc[b][k] * keep[b,d] * desc[b][d]  and -- only when the head has its
// own input dropout, i.e. its rows were left out of the d_desc GEMM -- d_desc[b][d] += keep[b,d] * sum_k dsc[b][k] Wa[k][d].
// block (32, 8): x = group of 8 columns, y = row phase; grid (ceil(D / 256), row splits).
template <typename T>
__global__ void __launch_bounds__(256) angle_bwd_kernel(const T* __restrict__ desc, const T* __restrict__ dzaug,
                                                        const T* __restrict__ wa, float* __restrict__ g_wa,
                                                        T* __restrict__ d_desc, int B, int D, int hid, int hid_pad,
                                                        int n_ang, int rows_per_block, uint32_t thr, float scale,
                                                        unsigned long long seed,
                                                        const unsigned long long* __restrict__ seed_dev) {
    __shared__ float red[8][32][17];
    pdl_wait();
    pdl_launch_dependents();
    if (seed_dev != nullptr) seed += *seed_dev;
    seed += HEAD_DROP_SEED_OFFSET;
    const int d0 = (blockIdx.x * 32 + threadIdx.x) * 8;
    const bool active = d0 < D;
    const int b_begin = blockIdx.y * rows_per_block, b_end = min(B, b_begin + rows_per_block);
    float acc[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) acc[q] = 0.f;
    if (active) {
        float w0[8], w1[8];
        const bool fix = thr != 0u && d_desc != nullptr;
        if (fix) { load8(wa + d0, w0); load8(wa + D + d0, w1); }
        for (int b = b_begin + threadIdx.y; b < b_end; b += 8) {
            const float ds = hldf(dzaug + (size_t)b * hid_pad + hid), dc = hldf(dzaug + (size_t)b * hid_pad + hid + 1);
            float v[8], k[8];
            load8(desc + (size_t)b * D + d0, v);
            if (thr != 0u) keep8(seed, (unsigned long long)b * D + d0, thr, scale, k);
            else {
#pragma unroll
                for (int q = 0; q < 8; ++q) k[q] = 1.f;
            }
#pragma unroll
            for (int q = 0; q < 8; ++q) { acc[q] += ds * k[q] * v[q]; acc[8 + q] += dc * k[q] * v[q]; }
            if (fix) {
                float dd[8];
                load8(d_desc + (size_t)b * D + d0, dd);
#pragma unroll
                for (int q = 0; q < 8; ++q) dd[q] += k[q] * (ds * w0[q] + dc * w1[q]);
                store8(d_desc + (size_t)b * D + d0, dd);
            }
        }
    }
#pragma unroll
    for (int q = 0; q < 16; ++q) red[threadIdx.y][threadIdx.x][q] = acc[q];
    __syncthreads();
    // 256 threads = 32 column groups x 8 columns: thread (x, y) finishes column d0 + y for both head rows
    if (active) {
        for (int k = 0; k < n_ang; ++k) {
            float t = 0.f;
#pragma unroll
            for (int y = 0; y < 8; ++y) t += red[y][threadIdx.x][8 * k + threadIdx.y];
            atomicAdd(g_wa + (size_t)k * D + d0 + threadIdx.y, t);
        }
    }
}

static int angle_rows(int angle_loss) { return angle_loss == VPR_ANGLE_NONE ? 0 : angle_loss == VPR_ANGLE_SCALAR_MAAE ? 1 : 2; }

struct HeadsPlan { size_t waug, zacc, sc_raw, dzaug, total; int hid_pad; };

static HeadsPlan heads_plan(const VprHeadsShape& s, int dtype) {
    const size_t e = dtype == VPR_BF16 ? 2 : 4;
    HeadsPlan p;
    p.hid_pad = (s.hidden + 2 + 7) & ~7;
    size_t off = 0;
    p.waug = carve_size(off, (size_t)p.hid_pad * s.D * e);       // first: vpr_heads_pack fills it ahead of the call
    p.zacc = carve_size(off, (size_t)s.B * p.hid_pad * 4);
    p.sc_raw = carve_size(off, (size_t)s.B * 2 * 4);
    p.dzaug = carve_size(off, (size_t)s.B * p.hid_pad * e);
    p.total = (off + 255) & ~size_t(255);
    return p;
}

template <typename T>
static int heads_pack_t(const VprHeadsParams& p, char* ws, const VprHeadsShape& s, int dtype, cudaStream_t st) {
    const HeadsPlan hp = heads_plan(s, dtype);
    HeadsPackArgs a;
    a.w0 = p.reg_w0;
    a.wa = angle_rows(s.angle_loss) > 0 ? p.ang_w : nullptr;
    a.waug = ws + hp.waug;
    a.n0_4 = (long long)s.hidden * s.D / 4;
    a.na_4 = a.wa != nullptr ? (long long)angle_rows(s.angle_loss) * s.D / 4 : 0;
    a.total_4 = (long long)hp.hid_pad * s.D / 4;
    VPR_REQUIRE((reinterpret_cast<uintptr_t>(a.w0) & 15) == 0 && (reinterpret_cast<uintptr_t>(a.wa) & 15) == 0,
                "head weights must be 16-byte aligned");
    long long blocks = (a.total_4 + 255) / 256;
    if (blocks > (long long)num_sms() * 8) blocks = (long long)num_sms() * 8;
    pack_heads_kernel<T><<<(unsigned)blocks, 256, 0, st>>>(a);
    return launch_ok("pack_heads_kernel");
}

// zero every accumulator of one call: zacc (workspace), the three losses and the gradient buffers that are reduced
// into with atomics; ranges that touch are merged into one memset
static int heads_zero(float* losses, const VprHeadsGrads& gr, float* zacc, const VprHeadsShape& s, const HeadsPlan& hp,
                      bool fwd_only, cudaStream_t st) {
    const int n_ang = angle_rows(s.angle_loss);
    const bool angle = n_ang > 0;
    struct Rng { void* p; size_t n; };
    Rng r[8];
    int cnt = 0;
    r[cnt++] = {zacc, (size_t)s.B * hp.hid_pad * 4};
    r[cnt++] = {losses, 3 * sizeof(float)};
    if (!fwd_only) {
        r[cnt++] = {gr.reg_b0, sizeof(float) * s.hidden};
        r[cnt++] = {gr.reg_w1, sizeof(float) * 2 * s.hidden};
        r[cnt++] = {gr.reg_b1, sizeof(float) * 2};
        if (gr.ang_w) r[cnt++] = {gr.ang_w, sizeof(float) * (angle ? n_ang : 2) * s.D};
        if (gr.ang_b) r[cnt++] = {gr.ang_b, sizeof(float) * (angle ? n_ang : 2)};
    }
    for (int i = 1; i < cnt; ++i)
        for (int j = i; j > 0 && (char*)r[j].p < (char*)r[j - 1].p; --j) { Rng t = r[j]; r[j] = r[j - 1]; r[j - 1] = t; }
    int i = 0;
    while (i < cnt) {
        char* p0 = (char*)r[i].p;
        size_t nb = r[i].n;
        int j = i + 1;
        while (j < cnt && (char*)r[j].p == p0 + nb) { nb += r[j].n; ++j; }
        VPR_CUDA_OK(cudaMemsetAsync(p0, 0, nb, st));
        i = j;
    }
    return 0;
}

template <typename T>
static int heads_t(const VprHeadsParams& p, const T* desc, const float* latlon_target, const float* angle_target_deg,
                   float* latlon_pred, float* sincos_pred, float* losses, VprHeadsGrads* grp, T* d_desc, char* ws,
                   const VprHeadsShape& s, int dtype, cudaStream_t st) {
    const HeadsPlan hp = heads_plan(s, dtype);
    T* waug = (T*)(ws + hp.waug); float* zacc = (float*)(ws + hp.zacc); float* sc_raw = (float*)(ws + hp.sc_raw);
    T* dzaug = (T*)(ws + hp.dzaug);
    const int okT = dtype == VPR_BF16 ? OUT_BF16 : OUT_F32;
    const int n_ang = angle_rows(s.angle_loss);
    const bool angle = n_ang > 0;
    const bool fwd_only = (s.flags & VPR_HEADS_FORWARD_ONLY) != 0;
    const bool head_drop = angle && s.head_dropout_p > 0.f;
    const bool fold = angle && !head_drop;             // angle rows ride in the regressor's GEMMs
    const int nsm = num_sms();
    VprHeadsGrads none = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    VprHeadsGrads& gr = grp != nullptr ? *grp : none;
    const unsigned long long* seed_dev = reinterpret_cast<const unsigned long long*>(s.seed_offset_dev);
    const bool skip_wgrad = (s.flags & VPR_HEADS_SKIP_WGRAD) != 0, wgrad_only = (s.flags & VPR_HEADS_WGRAD_ONLY) != 0;
    const uint32_t head_thr = head_drop ? dropout_threshold(s.head_dropout_p) : 0u;
    const float head_scale = head_drop ? 1.f / (1.f - s.head_dropout_p) : 1.f;
    GemmDesc g;
    auto angle_bwd = [&](bool pdl) -> int {
        const int col_blocks = (s.D + 255) / 256;
        int splits = (2 * nsm) / col_blocks; if (splits < 1) splits = 1;
        int rows_per = (s.B + splits - 1) / splits; rows_per = (rows_per + 7) & ~7;
        splits = (s.B + rows_per - 1) / rows_per;
        VPR_CUDA_OK(launch_ex(angle_bwd_kernel<T>, dim3((unsigned)col_blocks, (unsigned)splits), dim3(32, 8), (size_t)0, st, pdl, 1,
                              desc, (const T*)dzaug, (const T*)(waug + (size_t)s.hidden * s.D), gr.ang_w, d_desc, s.B, s.D,
                              s.hidden, hp.hid_pad, n_ang, rows_per, head_thr, head_scale, (unsigned long long)s.seed, seed_dev));
        return launch_ok("angle_bwd_kernel");
    };
    // the two launches nothing downstream of d_desc waits for: dW0[hid, D] = dZ^T desc and (unless it also patches
    // d_desc) the angle head's weight gradient
    auto wgrad = [&](bool first_pdl) -> int {
        g = GemmDesc(); g.dtype = dtype;
        g.A = dzaug; g.a_mn = 1; g.lda = hp.hid_pad; g.B = desc; g.b_mn = 1; g.ldb = s.D; g.M = s.hidden; g.N = s.D; g.K = s.B;
        g.out = gr.reg_w0; g.out_kind = OUT_F32; g.ldo = s.D;
        g.pdl = first_pdl ? 1 : 0;
        VPR_TRY(gemm_run(g, st));
        if (angle && !head_drop) VPR_TRY(angle_bwd(true));
        return 0;
    };
    if (wgrad_only) return wgrad(false);

    // accumulators -> 0 first, so that the kernels below form an uninterrupted chain of programmatic dependent
    // launches (or, with VPR_HEADS_PREZEROED, already done by the caller off the critical path: vpr_heads_zero)
    if (!(s.flags & VPR_HEADS_PREZEROED)) VPR_TRY(heads_zero(losses, gr, zacc, s, hp, fwd_only, st));
    if (!(s.flags & VPR_HEADS_PREPACKED)) VPR_TRY(heads_pack_t<T>(p, ws, s, dtype, st));

    // zacc = desc W_aug^T   (split-K: only ceil(B/128) * ceil(N/256) output tiles exist)
    g = GemmDesc(); g.dtype = dtype;
    g.A = desc; g.lda = s.D; g.B = waug; g.ldb = s.D; g.M = s.B; g.N = fold ? hp.hid_pad : s.hidden; g.K = s.D;
    g.out = zacc; g.out_kind = OUT_ATOMIC_F32; g.ldo = hp.hid_pad;
    {
        static int bn_env = -2, ks_env = -2;              // A/B knobs: VPR_HEADS_FWD_BN, VPR_HEADS_FWD_KS
        if (bn_env == -2) { const char* e = getenv("VPR_HEADS_FWD_BN"); bn_env = e ? atoi(e) : 0; }
        if (ks_env == -2) { const char* e = getenv("VPR_HEADS_FWD_KS"); ks_env = e ? atoi(e) : 0; }
        // measured (tools/heads_bench.py, B=256): 128-wide single-CTA tiles x 12-16 K-splits 10.5 us, 256-wide pair tiles
        // x 24 splits 12.4 us -- the launch is all prologue, split-K atomics and drain
        g.bn = bn_env > 0 ? bn_env : (s.B <= 512 ? 128 : 0);
        const GemmTilePlan tp = gemm_tile_plan(&g, 1);
        const int tiles = ((s.B + tp.tile_m - 1) / tp.tile_m) * ((g.N + tp.tile_n - 1) / tp.tile_n);
        g.k_splits = ks_env > 0 ? ks_env : tp.slots / tiles;
        if (g.k_splits < 1) g.k_splits = 1;
    }
    g.pdl = 1;
    VPR_TRY(gemm_run(g, st));
    if (head_drop) {
        VPR_CUDA_OK(launch_ex(angle_fwd_masked_kernel<T>, dim3((unsigned)s.B), dim3(256), (size_t)0, st, true, 1, desc,
                              (const T*)(waug + (size_t)s.hidden * s.D), sc_raw, s.D, head_thr, head_scale,
                              (unsigned long long)s.seed, seed_dev));
        VPR_TRY(launch_ok("angle_fwd_masked_kernel"));
    }
    TailArgs a;
    a.zacc = zacc; a.sc_raw = head_drop ? sc_raw : nullptr; a.b0 = p.reg_b0; a.w1 = p.reg_w1; a.b1 = p.reg_b1; a.ba = p.ang_b;
    a.latlon_target = latlon_target; a.angle_target_deg = angle_target_deg;
    a.latlon_pred = latlon_pred; a.sincos_pred = sincos_pred; a.losses = losses;
    a.g_b0 = gr.reg_b0; a.g_w1 = gr.reg_w1; a.g_b1 = gr.reg_b1; a.g_ba = gr.ang_b;
    a.dzaug = dzaug; a.B = s.B; a.hid = s.hidden; a.hid_pad = hp.hid_pad;
    a.latlon_loss = s.latlon_loss; a.angle_loss = s.angle_loss; a.global_B = s.global_B > 0 ? s.global_B : s.B;
    a.forward_only = fwd_only ? 1 : 0;
    a.huber_delta = s.huber_delta; a.loss_scale = s.loss_scale;
    a.drop_thr = s.reg_dropout_p > 0.f ? dropout_threshold(s.reg_dropout_p) : 0u;
    a.drop_scale = s.reg_dropout_p > 0.f ? 1.f / (1.f - s.reg_dropout_p) : 1.f;
    a.seed = s.seed; a.seed_dev = seed_dev;
    VPR_CUDA_OK(launch_ex(heads_tail_kernel<T>, dim3((unsigned)((s.B + TAIL_ROWS - 1) / TAIL_ROWS)), dim3(TAIL_THREADS), (size_t)0, st,
                          true, 1, a));
    VPR_TRY(launch_ok("heads_tail_kernel"));
    if (fwd_only) return 0;
    if (d_desc != nullptr) {   // d_desc = [dZ | dsc] [W0 ; Wa]   -- first: the SALAD backward waits for it
        g = GemmDesc(); g.dtype = dtype;
        g.A = dzaug; g.lda = hp.hid_pad; g.B = waug; g.b_mn = 1; g.ldb = s.D; g.M = s.B; g.N = s.D;
        g.K = fold ? hp.hid_pad : s.hidden;
        g.out = d_desc; g.out_kind = okT; g.ldo = s.D;
        // few rows: 128-wide single-CTA tiles fill the machine (2 x 66 tiles at B = 256), 256-wide pair tiles do not
        static int bn_env = -2;                           // A/B knob: VPR_HEADS_DDESC_BN
        if (bn_env == -2) { const char* e = getenv("VPR_HEADS_DDESC_BN"); bn_env = e ? atoi(e) : 0; }
        g.bn = bn_env > 0 ? bn_env : (((s.B + 255) / 256) * ((s.D + 255) / 256) * 2 <= nsm ? 128 : 0);
        g.pdl = 1;
        VPR_TRY(gemm_run(g, st));
    }
    if (angle && head_drop) VPR_TRY(angle_bwd(true));     // patches d_desc: stays on the main chain
    if (skip_wgrad) return 0;
    return wgrad(true);
}

static int check_heads_shape(const VprHeadsShape* s, int dtype) {
    VPR_REQUIRE(dtype == VPR_F32 || dtype == VPR_BF16, "dtype must be VPR_F32 or VPR_BF16");
    VPR_REQUIRE(s->B > 0 && s->D > 0 && s->D % 8 == 0 && s->hidden % 8 == 0, "bad heads shape");
    VPR_REQUIRE(s->latlon_loss >= VPR_LATLON_MSE && s->latlon_loss <= VPR_LATLON_EXTERNAL, "bad latlon_loss");
    VPR_REQUIRE(s->angle_loss >= VPR_ANGLE_NONE && s->angle_loss <= VPR_ANGLE_SCALAR_MAAE, "bad angle_loss");
    VPR_REQUIRE(!((s->flags & VPR_HEADS_SKIP_WGRAD) && (s->flags & VPR_HEADS_WGRAD_ONLY)) &&
                !((s->flags & VPR_HEADS_FORWARD_ONLY) && (s->flags & (VPR_HEADS_SKIP_WGRAD | VPR_HEADS_WGRAD_ONLY))),
                "contradictory VPR_HEADS_* flags");
    VPR_REQUIRE(s->reg_dropout_p >= 0.f && s->reg_dropout_p < 1.f && s->head_dropout_p >= 0.f && s->head_dropout_p < 1.f,
                "dropout probability out of range");
    return 0;
}

}  // namespace vpr

using namespace vpr;

extern "C" {

size_t vpr_heads_workspace_bytes(const VprHeadsShape* s, int dtype) { return heads_plan(*s, dtype).total; }

int vpr_heads_pack(const VprHeadsParams* p, void* workspace, size_t workspace_bytes, const VprHeadsShape* s, int dtype,
                   cudaStream_t stream) {
    VPR_REQUIRE(p && workspace && s && p->reg_w0, "null argument");
    VPR_TRY(check_device());
    VPR_TRY(check_heads_shape(s, dtype));
    VPR_REQUIRE(angle_rows(s->angle_loss) == 0 || p->ang_w, "angle head requested but its weight is null");
    VPR_REQUIRE(workspace_bytes >= heads_plan(*s, dtype).total, "workspace too small");
    if (dtype == VPR_BF16) return heads_pack_t<bf16>(*p, (char*)workspace, *s, dtype, stream);
    return heads_pack_t<float>(*p, (char*)workspace, *s, dtype, stream);
}

int vpr_heads_zero(float* losses, const VprHeadsGrads* grads, void* workspace, size_t workspace_bytes, const VprHeadsShape* s,
                   int dtype, cudaStream_t stream) {
    VPR_REQUIRE(losses && workspace && s, "null argument");
    VPR_TRY(check_device());
    VPR_TRY(check_heads_shape(s, dtype));
    const bool fwd_only = (s->flags & VPR_HEADS_FORWARD_ONLY) != 0;
    VPR_REQUIRE(fwd_only || (grads && grads->reg_b0 && grads->reg_w1 && grads->reg_b1), "gradient buffers are null");
    const HeadsPlan hp = heads_plan(*s, dtype);
    VPR_REQUIRE(workspace_bytes >= hp.total, "workspace too small");
    VprHeadsGrads none = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    return heads_zero(losses, grads ? *grads : none, (float*)((char*)workspace + hp.zacc), *s, hp, fwd_only, stream);
}

int vpr_heads_loss_fwd_bwd(const VprHeadsParams* p, const void* desc, const float* latlon_target,
                           const float* angle_target_deg, float* latlon_pred, float* sincos_pred, float* losses,
                           VprHeadsGrads* grads, void* d_desc, void* workspace, size_t workspace_bytes,
                           const VprHeadsShape* s, int dtype, cudaStream_t stream) {
    VPR_REQUIRE(p && desc && latlon_target && latlon_pred && losses && workspace && s, "null argument");
    VPR_TRY(check_device());
    VPR_TRY(check_heads_shape(s, dtype));
    const bool fwd_only = (s->flags & VPR_HEADS_FORWARD_ONLY) != 0;
    VPR_REQUIRE(fwd_only || (grads && grads->reg_w0 && grads->reg_b0 && grads->reg_w1 && grads->reg_b1),
                "gradient buffers are null (set VPR_HEADS_FORWARD_ONLY for inference)");
    if (s->angle_loss != VPR_ANGLE_NONE)
        VPR_REQUIRE(angle_target_deg && sincos_pred && p->ang_w && p->ang_b && (fwd_only || (grads->ang_w && grads->ang_b)),
                    "angle head requested but its buffers are null");
    VPR_REQUIRE(workspace_bytes >= heads_plan(*s, dtype).total, "workspace too small");
    if (dtype == VPR_BF16)
        return heads_t<bf16>(*p, (const bf16*)desc, latlon_target, angle_target_deg, latlon_pred, sincos_pred, losses,
                             grads, (bf16*)d_desc, (char*)workspace, *s, dtype, stream);
    return heads_t<float>(*p, (const float*)desc, latlon_target, angle_target_deg, latlon_pred, sincos_pred, losses,
                          grads, (float*)d_desc, (char*)workspace, *s, dtype, stream);
}

}  // extern "C"
