// Geopose heads + losses, forward and backward in one call (SURVEY.md section 8 rows a8, a9,
// Appendix B).  The 8448 -> 512 regressor layer and its two backward GEMMs run on the shared
// tcgen05 GEMM; everything after it (ReLU, dropout, the 512 -> 2 layer, the sin/cos head, all
// loss variants and their gradients) is one fused per-row kernel.
#include "common.cuh"
#include "gemm.h"
#include "../../include/vpr_b200.h"

#include <math.h>

namespace vpr {

__device__ __forceinline__ float hldf(const float* p) { return *p; }
__device__ __forceinline__ float hldf(const bf16* p) { return __bfloat162float(*p); }
__device__ __forceinline__ void hstf(float* p, float v) { *p = v; }
__device__ __forceinline__ void hstf(bf16* p, float v) { *p = __float2bfloat16_rn(v); }

#define RAD2DEG 57.29577951308232f
#define DEG2RAD 0.017453292519943295f

// python-style remainder, result in [0, 360)
__device__ __forceinline__ float mod360(float x) {
    float r = fmodf(x, 360.f);
    if (r < 0.f) r += 360.f;
    return r;
}

// W_aug [hid_pad, D]: rows [0,hid) = regressor.0.weight, rows hid, hid+1 = angle head, rest 0
template <typename T>
__global__ void pack_heads_kernel(const float* __restrict__ w0, const float* __restrict__ wa, T* __restrict__ waug,
                                  long long hidD, long long D, long long total) {
    pdl_wait();
    pdl_launch_dependents();
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        float v = 0.f;
        if (i < hidD) v = w0[i];
        else if (i < hidD + 2 * D && wa != nullptr) v = wa[i - hidD];
        hstf(waug + i, v);
    }
}

// sc_raw[b][k] = desc[b] . Wa[k]   (fp32 accumulate against the fp32 master head weights)
template <typename T>
__global__ void __launch_bounds__(256) angle_fwd_kernel(const T* __restrict__ desc, const float* __restrict__ wa,
                                                        float* __restrict__ sc_raw, int D) {
    __shared__ float red[2][8];
    pdl_wait();
    pdl_launch_dependents();
    const int b = blockIdx.x;
    const T* x = desc + (size_t)b * D;
    float a0 = 0.f, a1 = 0.f;
    for (int i = threadIdx.x; i < D; i += 256) {
        const float v = hldf(x + i);
        a0 += v * __ldg(wa + i);
        a1 += v * __ldg(wa + D + i);
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
    const float* zacc;      // [B, hid] pre-activation (no bias)
    const float* sc_raw;    // [B, 2] or null
    const float* b0; const float* w1; const float* b1; const float* ba;
    const float* latlon_target; const float* angle_target_deg;
    float* latlon_pred; float* sincos_pred; float* losses;
    float* g_b0; float* g_w1; float* g_b1; float* g_ba;
    void* dzaug;            // [B, hid_pad] run dtype
    int B, hid, hid_pad, latlon_loss, angle_loss, global_B;
    float huber_delta, loss_scale;
    uint32_t drop_thr; float drop_scale; unsigned long long seed;
};

template <typename T>
__global__ void __launch_bounds__(128) heads_tail_kernel(const TailArgs a) {
    __shared__ float red[2][4];
    __shared__ float dpred_s[2];
    pdl_wait();
    pdl_launch_dependents();
    const int b = blockIdx.x, tid = threadIdx.x;
    const float* zrow = a.zacc + (size_t)b * a.hid;
    // z = dropout(relu(zacc + b0)); pred = W1 z + b1
    float p0 = 0.f, p1 = 0.f;
    for (int i = tid; i < a.hid; i += 128) {
        float z = fmaxf(zrow[i] + a.b0[i], 0.f);
        if (a.drop_thr != 0u)
            z = dropout_keep(a.seed, (unsigned long long)b * a.hid + i, a.drop_thr) ? z * a.drop_scale : 0.f;
        p0 += z * a.w1[i];
        p1 += z * a.w1[a.hid + i];
    }
    p0 = warp_sum(p0); p1 = warp_sum(p1);
    if ((tid & 31) == 0) { red[0][tid >> 5] = p0; red[1][tid >> 5] = p1; }
    __syncthreads();
    if (tid == 0) {
        float pred[2];
        pred[0] = red[0][0] + red[0][1] + red[0][2] + red[0][3] + a.b1[0];
        pred[1] = red[1][0] + red[1][1] + red[1][2] + red[1][3] + a.b1[1];
        a.latlon_pred[b * 2 + 0] = pred[0];
        a.latlon_pred[b * 2 + 1] = pred[1];
        const float inv2B = 1.f / (2.f * (float)a.global_B);
        float loss = 0.f;
        for (int k = 0; k < 2; ++k) {
            const float e = pred[k] - a.latlon_target[b * 2 + k];
            float g;
            if (a.latlon_loss == VPR_LATLON_MSE) { loss += e * e; g = 2.f * e; }
            else if (fabsf(e) <= a.huber_delta) { loss += 0.5f * e * e; g = e; }
            else { loss += a.huber_delta * (fabsf(e) - 0.5f * a.huber_delta); g = a.huber_delta * (e > 0.f ? 1.f : -1.f); }
            g *= inv2B * a.loss_scale;
            dpred_s[k] = g;
            atomicAdd(a.g_b1 + k, g);
        }
        atomicAdd(a.losses + 0, loss * inv2B);
        // ---- sin/cos angle head ----
        float ds = 0.f, dc = 0.f;
        if (a.angle_loss != VPR_ANGLE_NONE) {
            const float s = a.sc_raw[b * 2 + 0] + a.ba[0], c = a.sc_raw[b * 2 + 1] + a.ba[1];
            a.sincos_pred[b * 2 + 0] = s;
            a.sincos_pred[b * 2 + 1] = c;
            const float invB = 1.f / (float)a.global_B;
            const float tdeg = mod360(a.angle_target_deg[b]);
            float aloss = 0.f;
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
            atomicAdd(a.losses + 1, aloss * invB);
            ds *= invB * a.loss_scale;
            dc *= invB * a.loss_scale;
            atomicAdd(a.g_ba + 0, ds);
            atomicAdd(a.g_ba + 1, dc);
        }
        hstf(reinterpret_cast<T*>(a.dzaug) + (size_t)b * a.hid_pad + a.hid + 0, ds);
        hstf(reinterpret_cast<T*>(a.dzaug) + (size_t)b * a.hid_pad + a.hid + 1, dc);
    }
    __syncthreads();
    const float g0 = dpred_s[0], g1 = dpred_s[1];
    T* dz = reinterpret_cast<T*>(a.dzaug) + (size_t)b * a.hid_pad;
    for (int i = tid; i < a.hid_pad; i += 128) {
        if (i < a.hid) {
            float z = fmaxf(zrow[i] + a.b0[i], 0.f);
            float keep = 1.f;
            if (a.drop_thr != 0u)
                keep = dropout_keep(a.seed, (unsigned long long)b * a.hid + i, a.drop_thr) ? a.drop_scale : 0.f;
            const float zd = z * keep;
            atomicAdd(a.g_w1 + i, g0 * zd);
            atomicAdd(a.g_w1 + a.hid + i, g1 * zd);
            const float d = z > 0.f ? (g0 * a.w1[i] + g1 * a.w1[a.hid + i]) * keep : 0.f;
            hstf(dz + i, d);
            if (d != 0.f) atomicAdd(a.g_b0 + i, d);
        } else if (i >= a.hid + 2) {
            hstf(dz + i, 0.f);
        }
    }
}

// dWa[k][d] = sum_b dzaug[b][hid+k] * desc[b][d]
template <typename T>
__global__ void __launch_bounds__(256) angle_wgrad_kernel(const T* __restrict__ desc, const T* __restrict__ dzaug,
                                                          float* __restrict__ g_wa, int B, int D, int hid, int hid_pad) {
    pdl_wait();
    pdl_launch_dependents();
    const int d = blockIdx.x * 256 + threadIdx.x;
    if (d >= D) return;
    float a0 = 0.f, a1 = 0.f;
    for (int b = 0; b < B; ++b) {
        const float v = hldf(desc + (size_t)b * D + d);
        a0 += hldf(dzaug + (size_t)b * hid_pad + hid) * v;
        a1 += hldf(dzaug + (size_t)b * hid_pad + hid + 1) * v;
    }
    g_wa[d] = a0;
    g_wa[D + d] = a1;
}

struct HeadsPlan { size_t waug, zacc, sc_raw, dzaug, total; int hid_pad; };

static HeadsPlan heads_plan(const VprHeadsShape& s, int dtype) {
    const size_t e = dtype == VPR_BF16 ? 2 : 4;
    HeadsPlan p;
    p.hid_pad = (s.hidden + 2 + 7) & ~7;
    size_t off = 0;
    p.waug = carve_size(off, (size_t)p.hid_pad * s.D * e);
    p.zacc = carve_size(off, (size_t)s.B * s.hidden * 4);
    p.sc_raw = carve_size(off, (size_t)s.B * 2 * 4);
    p.dzaug = carve_size(off, (size_t)s.B * p.hid_pad * e);
    p.total = (off + 255) & ~size_t(255);
    return p;
}

template <typename T>
static int heads_t(const VprHeadsParams& p, const T* desc, const float* latlon_target, const float* angle_target_deg,
                   float* latlon_pred, float* sincos_pred, float* losses, VprHeadsGrads& gr, T* d_desc, char* ws,
                   const VprHeadsShape& s, int dtype, cudaStream_t st) {
    const HeadsPlan hp = heads_plan(s, dtype);
    T* waug = (T*)(ws + hp.waug); float* zacc = (float*)(ws + hp.zacc); float* sc_raw = (float*)(ws + hp.sc_raw);
    T* dzaug = (T*)(ws + hp.dzaug);
    const int okT = dtype == VPR_BF16 ? OUT_BF16 : OUT_F32;
    const bool angle = s.angle_loss != VPR_ANGLE_NONE;
    const int nsm = num_sms();

    const long long hidD = (long long)s.hidden * s.D;
    // accumulators -> 0 first (adjacent buffers in one memset), so that the kernels below form an uninterrupted chain
    // of programmatic dependent launches
    {
        struct Rng { void* p; size_t n; };
        Rng r[7];
        int cnt = 0;
        r[cnt++] = {zacc, (size_t)s.B * s.hidden * 4};
        r[cnt++] = {losses, 2 * sizeof(float)};
        r[cnt++] = {gr.reg_b0, sizeof(float) * s.hidden};
        r[cnt++] = {gr.reg_w1, sizeof(float) * 2 * s.hidden};
        r[cnt++] = {gr.reg_b1, sizeof(float) * 2};
        if (gr.ang_w && !angle) r[cnt++] = {gr.ang_w, sizeof(float) * 2 * s.D};
        if (gr.ang_b) r[cnt++] = {gr.ang_b, sizeof(float) * 2};
        int i = 0;
        while (i < cnt) {
            char* p0 = (char*)r[i].p;
            size_t nb = r[i].n;
            int j = i + 1;
            while (j < cnt && (char*)r[j].p == p0 + nb) { nb += r[j].n; ++j; }
            VPR_CUDA_OK(cudaMemsetAsync(p0, 0, nb, st));
            i = j;
        }
    }
    VPR_CUDA_OK(launch_ex(pack_heads_kernel<T>, dim3((unsigned)(nsm * 4)), dim3(256), (size_t)0, st, false, 1, p.reg_w0,
                          angle ? p.ang_w : (const float*)nullptr, waug, hidD, (long long)s.D, (long long)hp.hid_pad * s.D));
    VPR_TRY(launch_ok("pack_heads_kernel"));

    GemmDesc g;
    // zacc = desc W0^T   (split-K: only ceil(B/128) * hidden/256 output tiles exist)
    g = GemmDesc(); g.dtype = dtype;
    g.A = desc; g.lda = s.D; g.B = waug; g.ldb = s.D; g.M = s.B; g.N = s.hidden; g.K = s.D;
    g.out = zacc; g.out_kind = OUT_ATOMIC_F32; g.ldo = s.hidden;
    {
        int tiles = ((s.B + 127) / 128) * ((s.hidden + 255) / 256);
        g.k_splits = nsm / tiles; if (g.k_splits < 1) g.k_splits = 1;
    }
    g.pdl = 1;
    VPR_TRY(gemm_run(g, st));
    if (angle) {
        VPR_CUDA_OK(launch_ex(angle_fwd_kernel<T>, dim3((unsigned)s.B), dim3(256), (size_t)0, st, true, 1, desc, p.ang_w, sc_raw, s.D));
        VPR_TRY(launch_ok("angle_fwd_kernel"));
    }
    TailArgs a;
    a.zacc = zacc; a.sc_raw = sc_raw; a.b0 = p.reg_b0; a.w1 = p.reg_w1; a.b1 = p.reg_b1; a.ba = p.ang_b;
    a.latlon_target = latlon_target; a.angle_target_deg = angle_target_deg;
    a.latlon_pred = latlon_pred; a.sincos_pred = sincos_pred; a.losses = losses;
    a.g_b0 = gr.reg_b0; a.g_w1 = gr.reg_w1; a.g_b1 = gr.reg_b1; a.g_ba = gr.ang_b;
    a.dzaug = dzaug; a.B = s.B; a.hid = s.hidden; a.hid_pad = hp.hid_pad;
    a.latlon_loss = s.latlon_loss; a.angle_loss = s.angle_loss; a.global_B = s.global_B > 0 ? s.global_B : s.B;
    a.huber_delta = s.huber_delta; a.loss_scale = s.loss_scale;
    a.drop_thr = s.reg_dropout_p > 0.f ? dropout_threshold(s.reg_dropout_p) : 0u;
    a.drop_scale = s.reg_dropout_p > 0.f ? 1.f / (1.f - s.reg_dropout_p) : 1.f;
    a.seed = s.seed;
    VPR_CUDA_OK(launch_ex(heads_tail_kernel<T>, dim3((unsigned)s.B), dim3(128), (size_t)0, st, true, 1, a));
    VPR_TRY(launch_ok("heads_tail_kernel"));
    // dW0[hid, D] = dZ^T desc
    g = GemmDesc(); g.dtype = dtype;
    g.A = dzaug; g.a_mn = 1; g.lda = hp.hid_pad; g.B = desc; g.b_mn = 1; g.ldb = s.D; g.M = s.hidden; g.N = s.D; g.K = s.B;
    g.out = gr.reg_w0; g.out_kind = OUT_F32; g.ldo = s.D;
    g.pdl = 1;
    VPR_TRY(gemm_run(g, st));
    if (angle) {
        VPR_CUDA_OK(launch_ex(angle_wgrad_kernel<T>, dim3((unsigned)((s.D + 255) / 256)), dim3(256), (size_t)0, st, true, 1, desc,
                              (const T*)dzaug, gr.ang_w, s.B, s.D, s.hidden, hp.hid_pad));
        VPR_TRY(launch_ok("angle_wgrad_kernel"));
    }
    if (d_desc != nullptr) {   // d_desc = [dZ | dsc] [W0 ; Wa]
        g = GemmDesc(); g.dtype = dtype;
        g.A = dzaug; g.lda = hp.hid_pad; g.B = waug; g.b_mn = 1; g.ldb = s.D; g.M = s.B; g.N = s.D; g.K = hp.hid_pad;
        g.out = d_desc; g.out_kind = okT; g.ldo = s.D;
        g.pdl = 1;
        VPR_TRY(gemm_run(g, st));
    }
    return 0;
}

}  // namespace vpr

using namespace vpr;

extern "C" {

size_t vpr_heads_workspace_bytes(const VprHeadsShape* s, int dtype) { return heads_plan(*s, dtype).total; }

int vpr_heads_loss_fwd_bwd(const VprHeadsParams* p, const void* desc, const float* latlon_target,
                           const float* angle_target_deg, float* latlon_pred, float* sincos_pred, float* losses,
                           VprHeadsGrads* grads, void* d_desc, void* workspace, size_t workspace_bytes,
                           const VprHeadsShape* s, int dtype, cudaStream_t stream) {
    VPR_REQUIRE(p && desc && latlon_target && latlon_pred && losses && grads && workspace && s, "null argument");
    VPR_REQUIRE(dtype == VPR_F32 || dtype == VPR_BF16, "dtype must be VPR_F32 or VPR_BF16");
    VPR_REQUIRE(s->B > 0 && s->D > 0 && s->D % 8 == 0 && s->hidden % 8 == 0, "bad heads shape");
    VPR_REQUIRE(s->latlon_loss == VPR_LATLON_MSE || s->latlon_loss == VPR_LATLON_HUBER, "bad latlon_loss");
    VPR_REQUIRE(s->angle_loss >= VPR_ANGLE_NONE && s->angle_loss <= VPR_ANGLE_MSE, "bad angle_loss");
    if (s->angle_loss != VPR_ANGLE_NONE)
        VPR_REQUIRE(angle_target_deg && sincos_pred && p->ang_w && p->ang_b && grads->ang_w && grads->ang_b,
                    "angle head requested but its buffers are null");
    VPR_REQUIRE(workspace_bytes >= heads_plan(*s, dtype).total, "workspace too small");
    if (dtype == VPR_BF16)
        return heads_t<bf16>(*p, (const bf16*)desc, latlon_target, angle_target_deg, latlon_pred, sincos_pred, losses,
                             *grads, (bf16*)d_desc, (char*)workspace, *s, dtype, stream);
    return heads_t<float>(*p, (const float*)desc, latlon_target, angle_target_deg, latlon_pred, sincos_pred, losses,
                          *grads, (float*)d_desc, (char*)workspace, *s, dtype, stream);
}

}  // extern "C"
