// Validation metrics that directly follow the hot path, computed where the
// predictions already are instead of after a device->host copy of every batch:
//   * lat/lon: de-normalise with the z-score scaler, then the sums behind the reference's
//     final_loss = 0.5 * (sum dlat^2 + sum dlon^2) / N   (dinov2salad/dinov2salad_validation.py:101) and
//     MSE / RMSE / MAE / MAE-lat / MAE-lon                (dinov2salad/dinov2salad_finetuning_2.py:211-215);
//   * angle: atan2 -> degrees in [0,360) -> min(|d|, 360-|d|), summed for the MAAE
//     (angle_prediction/efficient_net/validation_script.py:180-202; [sin,cos] order of dino_v2_gemini.py:135-141);
//   * recall@k over the kNN output (a query is a hit at k if one of its first k neighbours carries its label).
// Sums are accumulated in double so the result does not depend on the reduction order beyond 1e-12.
#include "common.cuh"
#include "../../include/vpr_b200.h"

#include <math.h>

namespace vpr {

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// out[0..3] += sum dlat^2, sum dlon^2, sum |dlat|, sum |dlon|   (original units)
__global__ void __launch_bounds__(256) metrics_latlon_kernel(const float* __restrict__ pred, const float* __restrict__ target,
                                                             const double* __restrict__ mean, const double* __restrict__ scale,
                                                             int targets_scaled, long long N, double* __restrict__ out) {
    __shared__ double red[4][8];
    double m0 = 0.0, m1 = 0.0, s0 = 1.0, s1 = 1.0;
    if (mean != nullptr) { m0 = mean[0]; m1 = mean[1]; }
    if (scale != nullptr) { s0 = scale[0]; s1 = scale[1]; }
    double a[4] = {0.0, 0.0, 0.0, 0.0};
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x) {
        const float2 p = *reinterpret_cast<const float2*>(pred + 2 * i);
        const float2 t = *reinterpret_cast<const float2*>(target + 2 * i);
        const double plat = (double)p.x * s0 + m0, plon = (double)p.y * s1 + m1;
        const double tlat = targets_scaled ? (double)t.x * s0 + m0 : (double)t.x;
        const double tlon = targets_scaled ? (double)t.y * s1 + m1 : (double)t.y;
        const double d0 = plat - tlat, d1 = plon - tlon;
        a[0] += d0 * d0; a[1] += d1 * d1; a[2] += fabs(d0); a[3] += fabs(d1);
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const double v = warp_sum_d(a[k]);
        if (lane == 0) red[k][warp] = v;
    }
    __syncthreads();
    if (threadIdx.x < 4) {
        double tot = 0.0;
        for (int w = 0; w < 8; ++w) tot += red[threadIdx.x][w];
        atomicAdd(out + threadIdx.x, tot);
    }
}

// pred_deg[i] = atan2(sin, cos) in degrees, mapped to [0, 360); out[0] += sum of min(|d|, 360 - |d|).
// cos_first = 0: rows are [sin, cos] (dino_v2_gemini.py); 1: rows are [cos, sin] (efficient_net/validation_script.py)
__global__ void __launch_bounds__(256) metrics_angle_kernel(const float* __restrict__ sincos, const float* __restrict__ true_deg,
                                                            int cos_first, long long N, float* __restrict__ pred_deg,
                                                            double* __restrict__ out) {
    __shared__ double red[8];
    double a = 0.0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x) {
        const float2 p = *reinterpret_cast<const float2*>(sincos + 2 * i);
        const float sn = cos_first ? p.y : p.x, cs = cos_first ? p.x : p.y;
        float deg = atan2f(sn, cs) * 57.29577951308232f;
        float tgt = true_deg[i];
        if (cos_first) {                        // validation_script.py:182: (deg + 360) % 360, target as given
            deg = fmodf(deg + 360.f, 360.f);
        } else {                                // dino_v2_gemini.py:120-122,140: python-style % 360 on both
            deg = fmodf(deg, 360.f); if (deg < 0.f) deg += 360.f;
            tgt = fmodf(tgt, 360.f); if (tgt < 0.f) tgt += 360.f;
        }
        if (pred_deg != nullptr) pred_deg[i] = deg;
        const float diff = fabsf(deg - tgt);
        a += (double)fminf(diff, 360.f - diff);
    }
    a = warp_sum_d(a);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = a;
    __syncthreads();
    if (threadIdx.x == 0) {
        double tot = 0.0;
        for (int w = 0; w < 8; ++w) tot += red[w];
        atomicAdd(out, tot);
    }
}

// hits[j] += number of queries whose label appears among their first ks[j] neighbours; one thread per query
__global__ void __launch_bounds__(256) recall_kernel(const long long* __restrict__ idx, const long long* __restrict__ gallery_label,
                                                     const long long* __restrict__ query_label, long long Q, int k, long long G,
                                                     const int* __restrict__ ks, int nk, unsigned long long* __restrict__ hits) {
    __shared__ unsigned int cnt[16];
    if (threadIdx.x < 16) cnt[threadIdx.x] = 0u;
    __syncthreads();
    for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < Q; q += (long long)gridDim.x * blockDim.x) {
        const long long want = query_label[q];
        int first = 0x7fffffff;                                 // rank of the first correct neighbour (none: never a hit,
                                                                // whatever cut-offs the caller asks for)
        for (int j = 0; j < k; ++j) {
            const long long g = idx[q * k + j];
            if (g >= 0 && g < G && gallery_label[g] == want) { first = j; break; }
        }
        for (int j = 0; j < nk; ++j)
            if (first < ks[j]) atomicAdd(&cnt[j], 1u);
    }
    __syncthreads();
    if (threadIdx.x < nk && cnt[threadIdx.x] != 0u) atomicAdd(hits + threadIdx.x, (unsigned long long)cnt[threadIdx.x]);
}

static unsigned metrics_grid(long long N) {
    long long g = (N + 255) / 256;
    const long long cap = (long long)num_sms() * 4;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (unsigned)g;
}

}  // namespace vpr

using namespace vpr;

extern "C" {

int vpr_metrics_latlon(const float* pred, const float* target, const double* scaler_mean, const double* scaler_scale,
                       int targets_scaled, int64_t N, double* out_sums, cudaStream_t stream) {
    VPR_TRY(::vpr::check_device());
    VPR_REQUIRE(pred && target && out_sums, "null argument");
    VPR_REQUIRE(N >= 0, "bad row count");
    VPR_CUDA_OK(cudaMemsetAsync(out_sums, 0, 4 * sizeof(double), stream));
    if (N == 0) return 0;
    metrics_latlon_kernel<<<metrics_grid(N), 256, 0, stream>>>(pred, target, scaler_mean, scaler_scale, targets_scaled, N, out_sums);
    return launch_ok("metrics_latlon_kernel");
}

int vpr_metrics_angle(const float* sincos_pred, const float* true_deg, int cos_first, int64_t N, float* pred_deg,
                      double* out_sum, cudaStream_t stream) {
    VPR_TRY(::vpr::check_device());
    VPR_REQUIRE(sincos_pred && true_deg && out_sum, "null argument");
    VPR_REQUIRE(N >= 0, "bad row count");
    VPR_CUDA_OK(cudaMemsetAsync(out_sum, 0, sizeof(double), stream));
    if (N == 0) return 0;
    metrics_angle_kernel<<<metrics_grid(N), 256, 0, stream>>>(sincos_pred, true_deg, cos_first, N, pred_deg, out_sum);
    return launch_ok("metrics_angle_kernel");
}

int vpr_recall_at_k(const int64_t* idx, const int64_t* gallery_label, const int64_t* query_label, int64_t Q, int k, int64_t G,
                    const int* ks, int nk, unsigned long long* hits, cudaStream_t stream) {
    VPR_TRY(::vpr::check_device());
    VPR_REQUIRE(idx && gallery_label && query_label && ks && hits, "null argument");
    VPR_REQUIRE(Q >= 0 && k > 0 && G > 0 && nk > 0 && nk <= 16, "bad recall shape (1 <= nk <= 16)");
    VPR_CUDA_OK(cudaMemsetAsync(hits, 0, nk * sizeof(unsigned long long), stream));
    if (Q == 0) return 0;
    recall_kernel<<<metrics_grid(Q), 256, 0, stream>>>((const long long*)idx, (const long long*)gallery_label,
                                                       (const long long*)query_label, Q, k, G, ks, nk, hits);
    return launch_ok("recall_kernel");
}

}  // extern "C"
