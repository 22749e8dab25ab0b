// Image preprocessing in front of the backbone: the reference's
//   transforms.Resize((224, 224)) -> ToTensor() -> Normalize(mean, std)       (dinov2salad/dinov2salad_finetuning.py:45-50)
// applied to a PIL image on the data-loader's main thread.  torchvision hands the resize to Pillow, whose 8-bit
// resampler is integer arithmetic: a separable triangle filter whose support grows with the down-scaling factor
// (antialiasing), coefficients rounded to 22 fractional bits, a horizontal pass whose result is rounded and clamped
// to uint8 before the vertical pass.  The kernels below restate that arithmetic ([THIRD-PARTY] Pillow
// src/libImaging/Resample.c: precompute_coeffs, normalize_coeffs_8bpc, ImagingResampleHorizontal_8bpc /
// Vertical_8bpc) so the resized image is bit-identical; ToTensor and Normalize are the float32 operations
// u/255, (x - mean)/std with IEEE division, fused into the vertical pass, written channel-major in the run dtype.
//
// Coefficients are computed on the device in double with explicit round-to-nearest operations (no FMA contraction),
// in the same operation order as the C source, so the tables equal Pillow's.
#include "common.cuh"
#include "../../include/vpr_b200.h"

namespace vpr {

static constexpr int PP_PRECISION_BITS = 32 - 8 - 2;

__host__ __device__ inline int pp_ksize(int in_size, int out_size) {
    double fs = (double)in_size / (double)out_size;
    if (fs < 1.0) fs = 1.0;
    const double support = 1.0 * fs;                     // bilinear filter support 1.0
    int c = (int)support;
    if ((double)c < support) ++c;                        // ceil
    return c * 2 + 1;
}

// one thread per output coordinate: bounds[xx] = (xmin, count), coef[xx][ksize] (22-bit fixed point)
__global__ void pp_coeffs_kernel(int in_size, int out_size, int ksize, int* __restrict__ bounds, int* __restrict__ coef) {
    const int xx = blockIdx.x * blockDim.x + threadIdx.x;
    if (xx >= out_size) return;
    const double scale = __ddiv_rn((double)in_size, (double)out_size);
    double filterscale = scale;
    if (filterscale < 1.0) filterscale = 1.0;
    const double support = __dmul_rn(1.0, filterscale);
    const double center = __dadd_rn(0.0, __dmul_rn(__dadd_rn((double)xx, 0.5), scale));
    const double ss = __ddiv_rn(1.0, filterscale);
    int xmin = (int)__dadd_rn(__dadd_rn(center, -support), 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = (int)__dadd_rn(__dadd_rn(center, support), 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    int* k = coef + (size_t)xx * ksize;
    double ww = 0.0;
    for (int x = 0; x < xmax; ++x) {
        double a = __dmul_rn(__dadd_rn(__dadd_rn((double)(x + xmin), -center), 0.5), ss);
        if (a < 0.0) a = -a;
        const double w = a < 1.0 ? __dadd_rn(1.0, -a) : 0.0;
        ww = __dadd_rn(ww, w);
    }
    for (int x = 0; x < ksize; ++x) {
        int q = 0;
        if (x < xmax) {
            double a = __dmul_rn(__dadd_rn(__dadd_rn((double)(x + xmin), -center), 0.5), ss);
            if (a < 0.0) a = -a;
            double w = a < 1.0 ? __dadd_rn(1.0, -a) : 0.0;
            if (ww != 0.0) w = __ddiv_rn(w, ww);
            const double sc = __dmul_rn(w, (double)(1 << PP_PRECISION_BITS));
            q = w < 0.0 ? (int)__dadd_rn(-0.5, sc) : (int)__dadd_rn(0.5, sc);
        }
        k[x] = q;
    }
    bounds[2 * xx] = xmin;
    bounds[2 * xx + 1] = xmax;
}

__device__ __forceinline__ int pp_clip8(int v) {
    v >>= PP_PRECISION_BITS;
    return v < 0 ? 0 : (v > 255 ? 255 : v);
}

// horizontal pass: in [H, W, 3] uint8 -> tmp [H, out_w, 3] uint8   (a batch of equal-sized images is simply B * H rows)
__global__ void __launch_bounds__(256) pp_horizontal_kernel(const uint8_t* __restrict__ in, int H, int W, int out_w, int ksize,
                                                            const int* __restrict__ bounds, const int* __restrict__ coef,
                                                            uint8_t* __restrict__ tmp) {
    const long long total = (long long)H * out_w;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int y = (int)(i / out_w), xx = (int)(i - (long long)y * out_w);
        const int xmin = __ldg(bounds + 2 * xx), cnt = __ldg(bounds + 2 * xx + 1);
        const int* k = coef + (size_t)xx * ksize;
        const uint8_t* row = in + ((size_t)y * W + xmin) * 3;
        int s0 = 1 << (PP_PRECISION_BITS - 1), s1 = s0, s2 = s0;
        for (int x = 0; x < cnt; ++x) {
            const int kv = __ldg(k + x);
            s0 += (int)row[3 * x] * kv; s1 += (int)row[3 * x + 1] * kv; s2 += (int)row[3 * x + 2] * kv;
        }
        uint8_t* o = tmp + (size_t)i * 3;
        o[0] = (uint8_t)pp_clip8(s0); o[1] = (uint8_t)pp_clip8(s1); o[2] = (uint8_t)pp_clip8(s2);
    }
}

__device__ __forceinline__ void pp_store(float* p, float v) { *p = v; }
__device__ __forceinline__ void pp_store(bf16* p, float v) { *p = __float2bfloat16_rn(v); }

struct PpNorm { float mean[3], stdv[3]; };

// vertical pass + ToTensor + Normalize: tmp [H, out_w, 3] uint8 -> out [3, out_h, out_w] (and optionally the resized
// uint8 image [out_h, out_w, 3] for inspection / parity)
template <typename TO>
__global__ void __launch_bounds__(256) pp_vertical_kernel(const uint8_t* __restrict__ tmp_all, int H, int B, int out_h, int out_w,
                                                          int ksize, const int* __restrict__ bounds, const int* __restrict__ coef,
                                                          const PpNorm nrm, TO* __restrict__ out_all, uint8_t* __restrict__ out_u8_all) {
    const long long total = (long long)out_h * out_w;
    for (long long gi = blockIdx.x * (long long)blockDim.x + threadIdx.x; gi < total * B; gi += (long long)gridDim.x * blockDim.x) {
        const int img = (int)(gi / total);
        const long long i = gi - (long long)img * total;
        const uint8_t* tmp = tmp_all + (size_t)img * H * out_w * 3;
        TO* out = out_all + (size_t)img * 3 * total;
        uint8_t* out_u8 = out_u8_all != nullptr ? out_u8_all + (size_t)img * total * 3 : nullptr;
        const int yy = (int)(i / out_w), xx = (int)(i - (long long)yy * out_w);
        const int ymin = __ldg(bounds + 2 * yy), cnt = __ldg(bounds + 2 * yy + 1);
        const int* k = coef + (size_t)yy * ksize;
        int s0 = 1 << (PP_PRECISION_BITS - 1), s1 = s0, s2 = s0;
        for (int y = 0; y < cnt; ++y) {
            const uint8_t* p = tmp + ((size_t)(y + ymin) * out_w + xx) * 3;
            const int kv = __ldg(k + y);
            s0 += (int)p[0] * kv; s1 += (int)p[1] * kv; s2 += (int)p[2] * kv;
        }
        const int u[3] = {pp_clip8(s0), pp_clip8(s1), pp_clip8(s2)};
        if (out_u8 != nullptr) {
            out_u8[i * 3] = (uint8_t)u[0]; out_u8[i * 3 + 1] = (uint8_t)u[1]; out_u8[i * 3 + 2] = (uint8_t)u[2];
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float v = __fdiv_rn(__fsub_rn(__fdiv_rn((float)u[c], 255.f), nrm.mean[c]), nrm.stdv[c]);
            pp_store(out + (size_t)c * total + i, v);
        }
    }
}

struct PpPlan { size_t tmp, bh, ch, bv, cv, total; int kh, kv; };
static PpPlan pp_plan(int H, int W, int out_h, int out_w, int B = 1) {
    PpPlan p;
    p.kh = pp_ksize(W, out_w);
    p.kv = pp_ksize(H, out_h);
    size_t off = 0;
    auto carve = [&](size_t bytes) { const size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
    p.tmp = carve((size_t)B * H * out_w * 3);
    p.bh = carve((size_t)out_w * 2 * sizeof(int));
    p.ch = carve((size_t)out_w * p.kh * sizeof(int));
    p.bv = carve((size_t)out_h * 2 * sizeof(int));
    p.cv = carve((size_t)out_h * p.kv * sizeof(int));
    p.total = off;
    return p;
}

}  // namespace vpr

using namespace vpr;

extern "C" {

size_t vpr_preprocess_workspace_bytes(int H, int W, int out_h, int out_w) {
    if (H <= 0 || W <= 0 || out_h <= 0 || out_w <= 0) return 0;
    return pp_plan(H, W, out_h, out_w).total;
}

size_t vpr_preprocess_batch_workspace_bytes(int B, int H, int W, int out_h, int out_w) {
    if (B <= 0 || H <= 0 || W <= 0 || out_h <= 0 || out_w <= 0) return 0;
    return pp_plan(H, W, out_h, out_w, B).total;
}

int vpr_preprocess_batch(const uint8_t* rgb_bhwc, int B, int H, int W, int out_h, int out_w, const float* mean3,
                         const float* std3, void* out_bchw, int dtype_out, uint8_t* out_u8_bhwc, void* workspace,
                         size_t workspace_bytes, cudaStream_t stream) {
    VPR_TRY(::vpr::check_device());
    VPR_REQUIRE(rgb_bhwc && mean3 && std3 && out_bchw && workspace, "null argument");
    VPR_REQUIRE(B > 0 && H > 0 && W > 0 && out_h > 0 && out_w > 0 && (long long)B * H < (1ll << 31), "bad image shape");
    VPR_REQUIRE(dtype_out == VPR_F32 || dtype_out == VPR_BF16, "dtype must be VPR_F32 or VPR_BF16");
    const PpPlan p = pp_plan(H, W, out_h, out_w, B);
    VPR_REQUIRE(workspace_bytes >= p.total, "workspace too small");
    uint8_t* ws = (uint8_t*)workspace;
    int* bh = (int*)(ws + p.bh); int* ch = (int*)(ws + p.ch); int* bv = (int*)(ws + p.bv); int* cv = (int*)(ws + p.cv);
    uint8_t* tmp = ws + p.tmp;
    pp_coeffs_kernel<<<(out_w + 127) / 128, 128, 0, stream>>>(W, out_w, p.kh, bh, ch);
    VPR_TRY(launch_ok("pp_coeffs_kernel"));
    pp_coeffs_kernel<<<(out_h + 127) / 128, 128, 0, stream>>>(H, out_h, p.kv, bv, cv);
    VPR_TRY(launch_ok("pp_coeffs_kernel"));
    const long long cap = (long long)num_sms() * 8;
    long long gh = ((long long)B * H * out_w + 255) / 256; if (gh > cap) gh = cap;
    pp_horizontal_kernel<<<(unsigned)gh, 256, 0, stream>>>(rgb_bhwc, B * H, W, out_w, p.kh, bh, ch, tmp);
    VPR_TRY(launch_ok("pp_horizontal_kernel"));
    PpNorm nrm;
    for (int c = 0; c < 3; ++c) { nrm.mean[c] = mean3[c]; nrm.stdv[c] = std3[c]; }
    long long gv = ((long long)B * out_h * out_w + 255) / 256; if (gv > cap) gv = cap;
    if (dtype_out == VPR_F32)
        pp_vertical_kernel<float><<<(unsigned)gv, 256, 0, stream>>>(tmp, H, B, out_h, out_w, p.kv, bv, cv, nrm, (float*)out_bchw,
                                                                    out_u8_bhwc);
    else
        pp_vertical_kernel<bf16><<<(unsigned)gv, 256, 0, stream>>>(tmp, H, B, out_h, out_w, p.kv, bv, cv, nrm, (bf16*)out_bchw,
                                                                   out_u8_bhwc);
    return launch_ok("pp_vertical_kernel");
}

int vpr_preprocess_image(const uint8_t* rgb_hwc, int H, int W, int out_h, int out_w, const float* mean3, const float* std3,
                         void* out_chw, int dtype_out, uint8_t* out_u8_hwc, void* workspace, size_t workspace_bytes,
                         cudaStream_t stream) {
    VPR_TRY(::vpr::check_device());
    return vpr_preprocess_batch(rgb_hwc, 1, H, W, out_h, out_w, mean3, std3, out_chw, dtype_out, out_u8_hwc, workspace,
                                workspace_bytes, stream);
}

}  // extern "C"
