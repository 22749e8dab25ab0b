// Backbone -> head seam: the backbone's final LayerNorm, the CLS / patch-token
// split and the cast to the head's run dtype in ONE pass over the last hidden state.
//
//   hidden [B, 1 + skip + n, d]   (CLS, `skip` register tokens, n patch tokens; pre-LayerNorm)
//   t      [B, d]      = LN(hidden[:, 0])                 ([UPSTREAM] x_norm_clstoken)
//   x_tok  [B, n, d]   = LN(hidden[:, 1 + skip:])         ([UPSTREAM] x_norm_patchtokens, token-major: the layout
//                                                          SaladHead consumes without the reshape/permute)
//
// HBM-bound: every row is read once and written once.  One warp owns a row; a lane holds its 8-element
// slices of the row in registers (16-byte loads and stores), so mean, variance and the affine output need
// a single read.  The backward (for a trainable last block feeding the head's dX back) recomputes x-hat from
// the saved (mean, rstd) pair and accumulates the LayerNorm weight/bias gradients per CTA before its atomics.
#include "common.cuh"
#include "../../include/vpr_b200.h"

namespace vpr {

static constexpr int SEAM_MAXCH = 8;           // 8 chunks x 256 elements: d <= 2048
static constexpr int SEAM_WARPS = 8;

__device__ __forceinline__ void seam_ld8(const float* p, float (&v)[8]) {
    const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void seam_ld8(const bf16* p, float (&v)[8]) {
    const uint4 u = *reinterpret_cast<const uint4*>(p);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) { v[2 * k] = __uint_as_float(w[k] << 16); v[2 * k + 1] = __uint_as_float(w[k] & 0xFFFF0000u); }
}
__device__ __forceinline__ void seam_ldg8(const float* p, float (&v)[8]) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p + 4));
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void seam_st8(float* p, const float (&v)[8]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void seam_st8(bf16* p, const float (&v)[8]) {
    uint32_t w[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        __nv_bfloat162 h2 = __floats2bfloat162_rn(v[2 * k], v[2 * k + 1]);
        w[k] = *reinterpret_cast<uint32_t*>(&h2);
    }
    *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
}

// raw 8-element slices (kept packed while a prefetched row waits in registers)
struct Raw8f { float4 a, b; };
__device__ __forceinline__ void seam_ldraw(const float* p, Raw8f& r) {
    r.a = *reinterpret_cast<const float4*>(p); r.b = *reinterpret_cast<const float4*>(p + 4);
}
__device__ __forceinline__ void seam_ldraw(const bf16* p, uint4& r) { r = *reinterpret_cast<const uint4*>(p); }
__device__ __forceinline__ void seam_unpack(const Raw8f& r, float (&v)[8]) {
    v[0] = r.a.x; v[1] = r.a.y; v[2] = r.a.z; v[3] = r.a.w; v[4] = r.b.x; v[5] = r.b.y; v[6] = r.b.z; v[7] = r.b.w;
}
__device__ __forceinline__ void seam_unpack(const uint4& u, float (&v)[8]) {
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) { v[2 * k] = __uint_as_float(w[k] << 16); v[2 * k + 1] = __uint_as_float(w[k] & 0xFFFF0000u); }
}
template <typename T> struct RawOf;
template <> struct RawOf<float> { typedef Raw8f type; };
template <> struct RawOf<bf16> { typedef uint4 type; };

// output row r in [0, B*(1+n)): image b = r / (1+n), j = r % (1+n); j == 0 is the CLS token
struct SeamRow {
    long long in_row;       // row of `hidden`
    long long tok_row;      // row of x_tok, or -1 for the CLS token
    int b;
};
// rows = B * (1 + n) < 2^31 is checked by the host entry points, so the split is one 32-bit division
__device__ __forceinline__ SeamRow seam_row(long long r, int n, int skip) {
    SeamRow o;
    const unsigned ru = (unsigned)r, per = (unsigned)(1 + n);
    const unsigned b = ru / per, j = ru - b * per;
    o.b = (int)b;
    o.in_row = (long long)b * (1 + skip + n) + (j == 0 ? 0 : j + skip);
    o.tok_row = j == 0 ? -1 : (long long)b * n + (j - 1);
    return o;
}

// Persistent over rows: each warp walks rows r, r + W, r + 2W ... and issues the loads of its next row before it
// reduces the current one, so two rows per warp (and ~190 KB per SM) are in flight.
template <typename TI, typename TO, int NCH, bool FULL>
__global__ void __launch_bounds__(SEAM_WARPS * 32, NCH <= 4 ? 3 : 1)
seam_fwd_kernel(const TI* __restrict__ hidden, const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                TO* __restrict__ x_tok, TO* __restrict__ t, float* __restrict__ stats, long long rows, int n, int skip, int d) {
    const int lane = threadIdx.x & 31;
    const long long stride = (long long)gridDim.x * SEAM_WARPS;
    long long r = (long long)blockIdx.x * SEAM_WARPS + (threadIdx.x >> 5);
    if (r >= rows) return;
    float v[NCH][8];
    typename RawOf<TI>::type nx[NCH];
    {
        const TI* x = hidden + seam_row(r, n, skip).in_row * d;
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
            const int e = c * 256 + lane * 8;
            if (FULL || e < d) seam_ldraw(x + e, nx[c]);
        }
    }
    for (; r < rows; r += stride) {
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
            const int e = c * 256 + lane * 8;
            if (FULL || e < d) seam_unpack(nx[c], v[c]);
        }
        if (r + stride < rows) {
            const TI* xn = hidden + seam_row(r + stride, n, skip).in_row * d;
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                const int e = c * 256 + lane * 8;
                if (FULL || e < d) seam_ldraw(xn + e, nx[c]);
            }
        }
        const SeamRow rw = seam_row(r, n, skip);
        float s = 0.f;
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
            const int e = c * 256 + lane * 8;
            if (FULL || e < d) {
#pragma unroll
                for (int i = 0; i < 8; ++i) s += v[c][i];
            }
        }
        const float mean = warp_sum(s) / (float)d;
        float q = 0.f;
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
            const int e = c * 256 + lane * 8;
            if (FULL || e < d) {
#pragma unroll
                for (int i = 0; i < 8; ++i) { const float u = v[c][i] - mean; q = fmaf(u, u, q); }
            }
        }
        const float rstd = rsqrtf(warp_sum(q) / (float)d + eps);
        if (stats != nullptr && lane == 0) { stats[2 * r] = mean; stats[2 * r + 1] = rstd; }
        TO* y = rw.tok_row < 0 ? t + (long long)rw.b * d : x_tok + rw.tok_row * d;
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
            const int e = c * 256 + lane * 8;
            if (FULL || e < d) {
                float g[8], bt[8], o[8];
                seam_ldg8(gamma + e, g);
                seam_ldg8(beta + e, bt);
#pragma unroll
                for (int i = 0; i < 8; ++i) o[i] = fmaf((v[c][i] - mean) * rstd, g[i], bt[i]);
                seam_st8(y + e, o);
            }
        }
    }
}

// d_hidden = LN backward of (d_x_tok, d_t); register-token rows get zero.  dgamma / dbeta accumulate in registers over
// the rows a warp walks, are combined per CTA in shared memory and leave with one atomic per element per CTA
// (the caller zeroes them).  A row's x and dy slices stay packed in registers and are unpacked twice (once for the
// two row sums, once for the output) so the next row's loads can already be in flight.
template <typename TI, typename TO, int NCH, bool FULL>
__global__ void __launch_bounds__(SEAM_WARPS * 32, NCH <= 4 ? 2 : 1)
seam_bwd_kernel(const TI* __restrict__ hidden, const float* __restrict__ gamma, const float* __restrict__ stats,
                const TO* __restrict__ d_x_tok, const TO* __restrict__ d_t, TI* __restrict__ d_hidden,
                float* __restrict__ dgamma, float* __restrict__ dbeta, long long rows, int n, int skip, int d) {
    __shared__ float sg[SEAM_MAXCH * 256], sb[SEAM_MAXCH * 256];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < d; i += SEAM_WARPS * 32) { sg[i] = 0.f; sb[i] = 0.f; }
    __syncthreads();
    float ag[NCH][8], ab[NCH][8];
#pragma unroll
    for (int c = 0; c < NCH; ++c)
#pragma unroll
        for (int i = 0; i < 8; ++i) { ag[c][i] = 0.f; ab[c][i] = 0.f; }
    typedef typename RawOf<TI>::type RawI;
    typedef typename RawOf<TO>::type RawO;
    RawI nxx[NCH];
    RawO nxd[NCH];
    bool nx_has_dy = false;
    float nx_mean = 0.f, nx_rstd = 0.f;
    const long long stride = (long long)gridDim.x * SEAM_WARPS;
    long long r = (long long)blockIdx.x * SEAM_WARPS + warp;
    auto issue = [&](long long rr) {
        const SeamRow rw = seam_row(rr, n, skip);
        const TI* x = hidden + rw.in_row * d;
        const TO* dy = rw.tok_row < 0 ? (d_t != nullptr ? d_t + (long long)rw.b * d : nullptr)
                                      : (d_x_tok != nullptr ? d_x_tok + rw.tok_row * d : nullptr);
        nx_has_dy = dy != nullptr;
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
            const int e = c * 256 + lane * 8;
            if (FULL || e < d) {
                seam_ldraw(x + e, nxx[c]);
                if (nx_has_dy) seam_ldraw(dy + e, nxd[c]);
            }
        }
        nx_mean = __ldg(stats + 2 * rr);
        nx_rstd = __ldg(stats + 2 * rr + 1);
    };
    if (r < rows) issue(r);
    for (; r < rows; r += stride) {
        RawI cx[NCH];
        RawO cd[NCH];
#pragma unroll
        for (int c = 0; c < NCH; ++c) { cx[c] = nxx[c]; cd[c] = nxd[c]; }
        const bool has_dy = nx_has_dy;
        const float mean = nx_mean, rstd = nx_rstd;
        if (r + stride < rows) issue(r + stride);
        const SeamRow rw = seam_row(r, n, skip);
        TI* dx = d_hidden + rw.in_row * d;
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
            const int e = c * 256 + lane * 8;
            if (FULL || e < d) {
                float xv[8], dv[8], gm[8];
                seam_unpack(cx[c], xv);
                if (has_dy) seam_unpack(cd[c], dv);
                else {
#pragma unroll
                    for (int i = 0; i < 8; ++i) dv[i] = 0.f;
                }
                seam_ldg8(gamma + e, gm);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float xh = (xv[i] - mean) * rstd, g = dv[i] * gm[i];
                    s1 += g;
                    s2 = fmaf(g, xh, s2);
                    ag[c][i] = fmaf(dv[i], xh, ag[c][i]);
                    ab[c][i] += dv[i];
                }
            }
        }
        const float m1 = warp_sum(s1) / (float)d, m2 = warp_sum(s2) / (float)d;
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
            const int e = c * 256 + lane * 8;
            if (FULL || e < d) {
                float xv[8], dv[8], gm[8], o[8];
                seam_unpack(cx[c], xv);
                if (has_dy) seam_unpack(cd[c], dv);
                else {
#pragma unroll
                    for (int i = 0; i < 8; ++i) dv[i] = 0.f;
                }
                seam_ldg8(gamma + e, gm);
#pragma unroll
                for (int i = 0; i < 8; ++i) o[i] = rstd * (dv[i] * gm[i] - m1 - (xv[i] - mean) * rstd * m2);
                seam_st8(dx + e, o);
            }
        }
        // register tokens of this image: zero gradient (written by the warp that owns the image's CLS row)
        if (rw.tok_row < 0 && skip > 0) {
            const float z[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            for (long long e = lane * 8; e < (long long)skip * d; e += 256) seam_st8(dx + d + e, z);
        }
    }
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
        const int e = c * 256 + lane * 8;
        if (FULL || e < d) {
#pragma unroll
            for (int i = 0; i < 8; ++i) { atomicAdd(&sg[e + i], ag[c][i]); atomicAdd(&sb[e + i], ab[c][i]); }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < d; i += SEAM_WARPS * 32) {
        if (dgamma != nullptr) atomicAdd(dgamma + i, sg[i]);
        if (dbeta != nullptr) atomicAdd(dbeta + i, sb[i]);
    }
}

template <typename TI, typename TO, int NCH>
static int seam_fwd_launch(const void* hidden, const float* g, const float* b, float eps, void* x_tok, void* t, float* stats,
                           long long rows, int n, int skip, int d, cudaStream_t st) {
    const long long want = (rows + SEAM_WARPS - 1) / SEAM_WARPS, cap = (long long)num_sms() * (NCH <= 4 ? 3 : 2);
    const unsigned grid = (unsigned)(want < cap ? want : cap);          // all CTAs resident: rows spread evenly
    if (d == NCH * 256)      // every lane owns a full slice of every chunk: no bounds predicates
        seam_fwd_kernel<TI, TO, NCH, true><<<grid, SEAM_WARPS * 32, 0, st>>>((const TI*)hidden, g, b, eps, (TO*)x_tok, (TO*)t,
                                                                             stats, rows, n, skip, d);
    else
        seam_fwd_kernel<TI, TO, NCH, false><<<grid, SEAM_WARPS * 32, 0, st>>>((const TI*)hidden, g, b, eps, (TO*)x_tok, (TO*)t,
                                                                              stats, rows, n, skip, d);
    return launch_ok("seam_fwd_kernel");
}
template <typename TI, typename TO, int NCH>
static int seam_bwd_launch(const void* hidden, const float* g, const float* stats, const void* dx_tok, const void* dt,
                           void* d_hidden, float* dg, float* db, long long rows, int n, int skip, int d, cudaStream_t st) {
    const long long want = (rows + SEAM_WARPS - 1) / SEAM_WARPS, cap = (long long)num_sms() * (NCH <= 4 ? 2 : 1);
    const unsigned grid = (unsigned)(want < cap ? want : cap);
    if (d == NCH * 256)
        seam_bwd_kernel<TI, TO, NCH, true><<<grid, SEAM_WARPS * 32, 0, st>>>((const TI*)hidden, g, stats, (const TO*)dx_tok,
                                                                             (const TO*)dt, (TI*)d_hidden, dg, db, rows, n, skip, d);
    else
        seam_bwd_kernel<TI, TO, NCH, false><<<grid, SEAM_WARPS * 32, 0, st>>>((const TI*)hidden, g, stats, (const TO*)dx_tok,
                                                                              (const TO*)dt, (TI*)d_hidden, dg, db, rows, n, skip, d);
    return launch_ok("seam_bwd_kernel");
}

#define SEAM_DISPATCH_NCH(FN, TI, TO, ...)                                   \
    do {                                                                     \
        const int nch = (d + 255) / 256;                                     \
        if (nch <= 3) return FN<TI, TO, 3>(__VA_ARGS__);                     \
        if (nch <= 4) return FN<TI, TO, 4>(__VA_ARGS__);                     \
        if (nch <= 6) return FN<TI, TO, 6>(__VA_ARGS__);                     \
        return FN<TI, TO, 8>(__VA_ARGS__);                                   \
    } while (0)

static int seam_fwd_dispatch(const void* hidden, const float* g, const float* b, float eps, void* x_tok, void* t, float* stats,
                             long long rows, int n, int skip, int d, int dt_in, int dt_out, cudaStream_t st) {
    if (dt_in == VPR_F32 && dt_out == VPR_F32) SEAM_DISPATCH_NCH(seam_fwd_launch, float, float, hidden, g, b, eps, x_tok, t, stats, rows, n, skip, d, st);
    if (dt_in == VPR_F32 && dt_out == VPR_BF16) SEAM_DISPATCH_NCH(seam_fwd_launch, float, bf16, hidden, g, b, eps, x_tok, t, stats, rows, n, skip, d, st);
    if (dt_in == VPR_BF16 && dt_out == VPR_F32) SEAM_DISPATCH_NCH(seam_fwd_launch, bf16, float, hidden, g, b, eps, x_tok, t, stats, rows, n, skip, d, st);
    SEAM_DISPATCH_NCH(seam_fwd_launch, bf16, bf16, hidden, g, b, eps, x_tok, t, stats, rows, n, skip, d, st);
}
static int seam_bwd_dispatch(const void* hidden, const float* g, const float* stats, const void* dx_tok, const void* dt,
                             void* d_hidden, float* dg, float* db, long long rows, int n, int skip, int d, int dt_in, int dt_out,
                             cudaStream_t st) {
    if (dt_in == VPR_F32 && dt_out == VPR_F32) SEAM_DISPATCH_NCH(seam_bwd_launch, float, float, hidden, g, stats, dx_tok, dt, d_hidden, dg, db, rows, n, skip, d, st);
    if (dt_in == VPR_F32 && dt_out == VPR_BF16) SEAM_DISPATCH_NCH(seam_bwd_launch, float, bf16, hidden, g, stats, dx_tok, dt, d_hidden, dg, db, rows, n, skip, d, st);
    if (dt_in == VPR_BF16 && dt_out == VPR_F32) SEAM_DISPATCH_NCH(seam_bwd_launch, bf16, float, hidden, g, stats, dx_tok, dt, d_hidden, dg, db, rows, n, skip, d, st);
    SEAM_DISPATCH_NCH(seam_bwd_launch, bf16, bf16, hidden, g, stats, dx_tok, dt, d_hidden, dg, db, rows, n, skip, d, st);
}

}  // namespace vpr

using namespace vpr;

extern "C" {

int vpr_seam_ln_split_fwd(const void* hidden, const float* ln_weight, const float* ln_bias, float eps, void* x_tok, void* t,
                          float* stats, int B, int n, int skip, int d, int dtype_in, int dtype_out, cudaStream_t stream) {
    VPR_TRY(::vpr::check_device());
    VPR_REQUIRE(hidden && ln_weight && ln_bias && x_tok && t, "null argument");
    VPR_REQUIRE(B > 0 && n > 0 && skip >= 0 && (long long)B * (1 + n) < (1ll << 31), "bad seam shape");
    VPR_REQUIRE(d > 0 && d % 8 == 0 && d <= SEAM_MAXCH * 256, "d must be a multiple of 8 and at most 2048");
    VPR_REQUIRE((dtype_in == VPR_F32 || dtype_in == VPR_BF16) && (dtype_out == VPR_F32 || dtype_out == VPR_BF16),
                "dtype must be VPR_F32 or VPR_BF16");
    return seam_fwd_dispatch(hidden, ln_weight, ln_bias, eps, x_tok, t, stats, (long long)B * (1 + n), n, skip, d, dtype_in,
                             dtype_out, stream);
}

int vpr_seam_ln_split_bwd(const void* hidden, const float* ln_weight, const float* stats, const void* d_x_tok, const void* d_t,
                          void* d_hidden, float* d_ln_weight, float* d_ln_bias, int B, int n, int skip, int d, int dtype_in,
                          int dtype_out, cudaStream_t stream) {
    VPR_TRY(::vpr::check_device());
    VPR_REQUIRE(hidden && ln_weight && stats && d_hidden, "null argument");
    VPR_REQUIRE(B > 0 && n > 0 && skip >= 0 && (long long)B * (1 + n) < (1ll << 31), "bad seam shape");
    VPR_REQUIRE(d > 0 && d % 8 == 0 && d <= SEAM_MAXCH * 256, "d must be a multiple of 8 and at most 2048");
    VPR_REQUIRE((dtype_in == VPR_F32 || dtype_in == VPR_BF16) && (dtype_out == VPR_F32 || dtype_out == VPR_BF16),
                "dtype must be VPR_F32 or VPR_BF16");
    if (d_ln_weight) VPR_CUDA_OK(cudaMemsetAsync(d_ln_weight, 0, sizeof(float) * d, stream));
    if (d_ln_bias) VPR_CUDA_OK(cudaMemsetAsync(d_ln_bias, 0, sizeof(float) * d, stream));
    return seam_bwd_dispatch(hidden, ln_weight, stats, d_x_tok, d_t, d_hidden, d_ln_weight, d_ln_bias, (long long)B * (1 + n), n,
                             skip, d, dtype_in, dtype_out, stream);
}

}  // extern "C"
