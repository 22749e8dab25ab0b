// SALAD backward, first step, one CTA per image: backward of the descriptor assembly (global L2 norm,
// per-cluster L2 norm) fused with the backward of the soft-assignment aggregation.
//
//   dAGG = d(out)/d(AGG)            closed form of one round of per-cluster sums (see finalize_fast.cuh)
//   dF[j, l] = sum_c P[j, c] dAGG[l, c]        dP[j, c] = sum_l F[j, l] dAGG[l, c]
//
// The 128 compute threads each own one cluster-feature row l: they read their AGG row and their row of the
// upstream gradient, reduce the per-cluster sums with register butterflies, and write their bf16 dAGG row
// into ONE swizzled shared-memory tile.  That tile is the B operand of both products -- K-major for dF
// (N = l, K = c) and, through an MN-major descriptor over the same bytes, for dP (N = c, K = l) -- so
// neither dAGG nor a transposed copy of it ever goes to HBM.  P[b] / F[b] arrive by TMA as 128-token
// chunks (A operands, K-major); the same threads then drain the accumulators: dF as bf16, dP as fp32,
// staged per warp and written with TMA tile stores.
// The token part of the norm backward (dTF and its column sums) is done by the same threads while the first
// chunk's MMAs run; the caller follows with the small masked dHT GEMM on a side branch.
//
//   warps 0-3: norm backward, then epilogue (one TMEM lane quarter each)   warp 4: TMA producer   warp 5: MMA issuer
#include "common.cuh"
#include "gemm.h"

#include <stdlib.h>

namespace vpr {

int make_map(CUtensorMap* map, const void* ptr, int esize, int64_t inner, int64_t rows, int64_t batch, int64_t ld,
             int64_t bs, int box_inner, int box_rows, const char* what);

static constexpr int GB_THREADS = 192;
static constexpr int GB_W_PROD = 4, GB_W_MMA = 5;
static constexpr int GB_BOX = 128 * 128;                 // 128 rows x 128 B
static constexpr int GB_STAGE = 3 * GB_BOX;              // P chunk (1 box) + F chunk (2 boxes: l 0-63, 64-127)
static constexpr int GB_STAGES = 1;                     // 81 KB per CTA -> two images per SM; the next chunk loads during the epilogue
static constexpr int GB_WBOX = 32 * 128;
static constexpr int GB_SMEM = 1024 + GB_STAGES * GB_STAGE + GB_BOX /* dAGG tile */ + 4 * GB_WBOX + 256;
static_assert(2 * GB_SMEM <= 220 * 1024, "two CTAs per SM");
static constexpr uint32_t GB_TM_DF = 0, GB_TM_DP = 128, GB_TM_COLS = 256;
#define GB_EPS 1e-12f

struct GbArgs {
    int n, g;
    const float* AGG;        // [B, 128, 64] fp32
    const float* TF;         // [B, g] fp32
    const bf16* d_out;       // [B, g + 8192]
    bf16* dTF;               // [B, g]: gradient of the token-MLP output
    float* g_bias_tok;       // [g] += column sums of dTF (token_features.2.bias gradient), or null
};

__device__ __forceinline__ float gb_colsum32(float (&v)[32], int lane) {
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
        const bool up = (lane & o) != 0;
#pragma unroll
        for (int i = 0; i < o; ++i) {
            const float send = up ? v[i] : v[i + o];
            const float keep = up ? v[i + o] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
        }
    }
    return v[0];
}

__global__ void __launch_bounds__(GB_THREADS, 2)
agg_bwd_kernel(const __grid_constant__ CUtensorMap tmP, const __grid_constant__ CUtensorMap tmF,
               const __grid_constant__ CUtensorMap tmDF, const __grid_constant__ CUtensorMap tmDP, const GbArgs args) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ float part[2][4][64];         // per-warp partial sums (sum a^2, sum a*dy) per cluster
    __shared__ float s_ic[64], s_dot[64];
    __shared__ float wred[2][4];
    __shared__ float sc[5];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t sDA = base + GB_STAGES * GB_STAGE;        // dAGG tile: [128 l rows][64 c] bf16, SWIZZLE_128B
    const uint32_t sStage = sDA + GB_BOX;
    const uint32_t bar = sStage + 4 * GB_WBOX;
    auto full = [&](int s) { return bar + 8u * s; };
    auto empty = [&](int s) { return bar + 8u * (GB_STAGES + s); };
    const uint32_t daready = bar + 8u * (2 * GB_STAGES);
    const uint32_t tfull = bar + 8u * (2 * GB_STAGES + 1);
    const uint32_t tempty = bar + 8u * (2 * GB_STAGES + 2);
    const uint32_t tmem_slot = bar + 8u * (2 * GB_STAGES + 3);
    uint8_t* smem_al = smem_raw + (base - smem_u32(smem_raw));
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_al + (tmem_slot - base));

    // warp index through a shuffle: provably warp-uniform (see gemm.cu: single-thread instructions then take uniform-register
    // operands directly instead of an election loop each)
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    const int b = blockIdx.x;
    const int chunks = (args.n + 127) / 128;

    if (warp == GB_W_PROD && lane == 0) {
        tma_prefetch_desc(&tmP); tma_prefetch_desc(&tmF); tma_prefetch_desc(&tmDF); tma_prefetch_desc(&tmDP);
    }
    if (warp == GB_W_MMA) {
        if (lane == 0) {
            for (int s = 0; s < GB_STAGES; ++s) { mbar_init(full(s), 1); mbar_init(empty(s), 1); }
            mbar_init(daready, 4); mbar_init(tfull, 1); mbar_init(tempty, 4);
            mbar_fence_init();
        }
        __syncwarp();
        tmem_alloc(tmem_slot, GB_TM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    pdl_wait();                  // everything above overlapped the previous launch's tail
    pdl_launch_dependents();

    if (warp == GB_W_PROD) {
        if (lane == 0) {
            for (int ch = 0; ch < chunks; ++ch) {
                const int s = ch % GB_STAGES;
                mbar_wait(empty(s), (((uint32_t)(ch / GB_STAGES)) & 1u) ^ 1u);
                const uint32_t st = base + (uint32_t)s * GB_STAGE;
                mbar_expect_tx(full(s), GB_STAGE);
                tma_load_3d(st, &tmP, full(s), 0, ch * 128, b);                      // rows beyond n are zero-filled
                tma_load_3d(st + GB_BOX, &tmF, full(s), 0, ch * 128, b);
                tma_load_3d(st + 2 * GB_BOX, &tmF, full(s), 64, ch * 128, b);
            }
        }
    } else if (warp == GB_W_MMA) {
        if (lane == 0) {
            constexpr uint32_t idesc_df = make_idesc_bf16(128, 128, false, false);   // A = P chunk, B = dAGG [l][c] K-major
            constexpr uint32_t idesc_dp = make_idesc_bf16(128, 64, false, true);     // A = F chunk, B = dAGG viewed [c contiguous][l rows]
            mbar_wait(daready, 0);
            tc_fence_after();
            for (int ch = 0; ch < chunks; ++ch) {
                const int s = ch % GB_STAGES;
                mbar_wait(full(s), ((uint32_t)(ch / GB_STAGES)) & 1u);
                mbar_wait(tempty, ((uint32_t)ch & 1u) ^ 1u);                         // previous chunk's accumulators drained
                tc_fence_after();
                const uint32_t st = base + (uint32_t)s * GB_STAGE;
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)                                       // dF chunk: K = c = 64
                    tc_mma_bf16(tmem_base + GB_TM_DF, make_smem_desc(st + ks * 32, 16u, 1024u),
                                make_smem_desc(sDA + ks * 32, 16u, 1024u), idesc_df, ks > 0 ? 1u : 0u);
#pragma unroll
                for (int ks = 0; ks < 8; ++ks)                                       // dP chunk: K = l = 128
                    tc_mma_bf16(tmem_base + GB_TM_DP, make_smem_desc(st + GB_BOX + (ks >> 2) * GB_BOX + (ks & 3) * 32, 16u, 1024u),
                                make_smem_desc(sDA + ks * 2048, GB_BOX, 1024u), idesc_dp, ks > 0 ? 1u : 0u);
                tc_commit(empty(s));
                tc_commit(tfull);
            }
        }
    } else {
        // ------------------------------ norm backward: thread = cluster-feature row l ------------------------------
        const int l = warp * 32 + lane;
        const int tid = threadIdx.x;             // 0..127
        const float* t = args.TF + (size_t)b * args.g;
        const bf16* dy = args.d_out + (size_t)b * (args.g + 8192);
        float ts = 0.f, tp = 0.f;
        for (int i = tid; i < args.g; i += 128) {
            const float v = __ldg(t + i);
            ts = fmaf(v, v, ts);
            tp = fmaf(v, __bfloat162float(dy[i]), tp);
        }
        ts = warp_sum(ts);
        tp = warp_sum(tp);
        if (lane == 0) { wred[0][warp] = ts; wred[1][warp] = tp; }
        float a[64], d[64];
        {
            const float* ag = args.AGG + ((size_t)b * 128 + l) * 64;
            const bf16* dr = dy + args.g + (size_t)l * 64;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const float4 v = __ldg(reinterpret_cast<const float4*>(ag) + i);
                a[4 * i] = v.x; a[4 * i + 1] = v.y; a[4 * i + 2] = v.z; a[4 * i + 3] = v.w;
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const uint4 u = __ldg(reinterpret_cast<const uint4*>(dr) + i);
                const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    d[8 * i + 2 * k] = __uint_as_float(w[k] << 16);
                    d[8 * i + 2 * k + 1] = __uint_as_float(w[k] & 0xFFFF0000u);
                }
            }
        }
        {
            float q0[32], q1[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) { q0[i] = a[i] * a[i]; q1[i] = a[32 + i] * a[32 + i]; }
            part[0][warp][lane] = gb_colsum32(q0, lane);
            part[0][warp][32 + lane] = gb_colsum32(q1, lane);
#pragma unroll
            for (int i = 0; i < 32; ++i) { q0[i] = a[i] * d[i]; q1[i] = a[32 + i] * d[32 + i]; }
            part[1][warp][lane] = gb_colsum32(q0, lane);
            part[1][warp][32 + lane] = gb_colsum32(q1, lane);
        }
        named_bar_sync(1, 128);
        if (warp == 0) {
            float zz = 0.f, sq = 0.f, ssc[2], qc[2], icv[2];
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const int c = lane + 32 * k;
                ssc[k] = part[0][0][c] + part[0][1][c] + part[0][2][c] + part[0][3][c];
                qc[k] = part[1][0][c] + part[1][1][c] + part[1][2][c] + part[1][3][c];
                const float nc = fmaxf(sqrtf(ssc[k]), GB_EPS);
                icv[k] = 1.f / nc;
                s_ic[c] = icv[k];
                zz += ssc[k] * icv[k] * icv[k];
                sq += icv[k] * qc[k];
            }
            zz = warp_sum(zz);
            sq = warp_sum(sq);
            const float tsum = wred[0][0] + wred[0][1] + wred[0][2] + wred[0][3];
            const float psum = wred[1][0] + wred[1][1] + wred[1][2] + wred[1][3];
            const float invt = 1.f / fmaxf(sqrtf(tsum), GB_EPS);
            zz += tsum * invt * invt;
            const float invN = 1.f / fmaxf(sqrtf(zz), GB_EPS);
            const float s = invN * (psum * invt + sq);                      // y . dy
#pragma unroll
            for (int k = 0; k < 2; ++k)
                s_dot[lane + 32 * k] = invN * (icv[k] * qc[k] - ssc[k] * icv[k] * icv[k] * invN * s);
            if (lane == 0) {
                sc[0] = invN; sc[1] = s; sc[2] = invt;
                sc[3] = invN * (psum * invt - tsum * invt * invt * invN * s);      // that . dz_t
            }
        }
        named_bar_sync(1, 128);
        {
            const float invN = sc[0], ns = sc[0] * sc[1];
#pragma unroll
            for (int i = 0; i < 8; ++i) {                                   // this thread's dAGG row, 8 x 16 bytes, swizzled
                uint32_t pk[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    float r2[2];
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int c = 8 * i + 2 * k + e;
                        const float ic = s_ic[c], ah = a[c] * ic;
                        r2[e] = ((d[c] - ah * ns) * invN - ah * s_dot[c]) * ic;
                    }
                    __nv_bfloat162 h2 = __floats2bfloat162_rn(r2[0], r2[1]);
                    pk[k] = *reinterpret_cast<uint32_t*>(&h2);
                }
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sDA + (uint32_t)l * 128u + (uint32_t)((i ^ (l & 7)) << 4)),
                             "r"(pk[0]), "r"(pk[1]), "r"(pk[2]), "r"(pk[3]) : "memory");
            }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(daready);
        {   // token part of the norm backward (the MMAs of the first chunk run meanwhile)
            const float invN = sc[0], ns = sc[0] * sc[1], invt = sc[2], dt_dot = sc[3];
            bf16* dT = args.dTF + (size_t)b * args.g;
            for (int i = tid; i < args.g; i += 128) {
                const float th = __ldg(t + i) * invt;
                const float dz = (__bfloat162float(dy[i]) - th * ns) * invN;
                const float dti = (dz - th * dt_dot) * invt;
                dT[i] = __float2bfloat16_rn(dti);
                if (args.g_bias_tok != nullptr) atomicAdd(args.g_bias_tok + i, dti);
            }
        }
        // ------------------------------ epilogue: thread = token row of the chunk ------------------------------
        const uint32_t sbuf = sStage + (uint32_t)warp * GB_WBOX;
        for (int ch = 0; ch < chunks; ++ch) {
            mbar_wait(tfull, (uint32_t)ch & 1u);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
            const int row0 = ch * 128 + warp * 32;
            // dF chunk: 128 bf16 columns = two 64-column boxes
#pragma unroll 1
            for (int bx = 0; bx < 2; ++bx) {
                if (lane == 0) bulk_wait_read<0>();
                __syncwarp();
#pragma unroll 1
                for (int c = 0; c < 2; ++c) {
                    uint32_t r[32];
                    tmem_ld32(taddr + GB_TM_DF + bx * 64 + c * 32, r);
                    tmem_ld_wait();
#pragma unroll
                    for (int jq = 0; jq < 4; ++jq) {
                        uint32_t pk[4];
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            __nv_bfloat162 h2 = __floats2bfloat162_rn(__uint_as_float(r[jq * 8 + 2 * i]), __uint_as_float(r[jq * 8 + 2 * i + 1]));
                            pk[i] = *reinterpret_cast<uint32_t*>(&h2);
                        }
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sbuf + (uint32_t)lane * 128u + (uint32_t)(((c * 4 + jq) ^ (lane & 7)) << 4)),
                                     "r"(pk[0]), "r"(pk[1]), "r"(pk[2]), "r"(pk[3]) : "memory");
                    }
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) { tma_store_3d(&tmDF, sbuf, bx * 64, row0, b); bulk_commit(); }
            }
            // dP chunk: 64 fp32 columns = two 32-column boxes
#pragma unroll 1
            for (int bx = 0; bx < 2; ++bx) {
                if (lane == 0) bulk_wait_read<0>();
                __syncwarp();
                uint32_t r[32];
                tmem_ld32(taddr + GB_TM_DP + bx * 32, r);
                tmem_ld_wait();
#pragma unroll
                for (int jq = 0; jq < 8; ++jq)
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sbuf + (uint32_t)lane * 128u + (uint32_t)((jq ^ (lane & 7)) << 4)),
                                 "r"(r[4 * jq]), "r"(r[4 * jq + 1]), "r"(r[4 * jq + 2]), "r"(r[4 * jq + 3]) : "memory");
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) { tma_store_3d(&tmDP, sbuf, bx * 32, row0, b); bulk_commit(); }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty);
        }
        if (lane == 0) bulk_wait_read<0>();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == GB_W_MMA) {
        tc_fence_after();
        tmem_dealloc(tmem_base, GB_TM_COLS);
    }
}

bool agg_bwd_supported(int dtype, int l, int m, int g) {
    static int enabled = -1;
    if (enabled < 0) {
        const char* e = getenv("VPR_AGG_FUSED");
        const char* e2 = getenv("VPR_AGG_BWD_FUSED");
        enabled = ((e != nullptr && e[0] == '0') || (e2 != nullptr && e2[0] == '0')) ? 0 : 1;
    }
    return enabled == 1 && dtype == DT_BF16 && l == 128 && m == 64 && g % 8 == 0;
}

int agg_bwd_run(const AggBwdDesc& d, cudaStream_t st) {
    CUtensorMap tmP, tmF, tmDF, tmDP;
    VPR_TRY(make_map(&tmP, d.P, 2, 64, d.n, d.B, 64, (int64_t)d.n * 64, 64, 128, "P"));
    VPR_TRY(make_map(&tmF, d.F, 2, 128, d.n, d.B, 128, (int64_t)d.n * 128, 64, 128, "F"));
    VPR_TRY(make_map(&tmDF, d.dF, 2, 128, d.n, d.B, 128, (int64_t)d.n * 128, 64, 32, "dF"));
    VPR_TRY(make_map(&tmDP, d.dP, 4, 64, d.n, d.B, 64, (int64_t)d.n * 64, 32, 32, "dP"));
    GbArgs a;
    a.n = d.n; a.g = d.g; a.AGG = d.AGG; a.TF = d.TF; a.d_out = reinterpret_cast<const bf16*>(d.d_out);
    a.dTF = reinterpret_cast<bf16*>(d.dTF); a.g_bias_tok = d.g_bias_tok;
    VPR_SMEM_OK(agg_bwd_kernel, GB_SMEM);
    VPR_CUDA_OK(launch_ex(agg_bwd_kernel, dim3((unsigned)d.B), dim3(GB_THREADS), (size_t)GB_SMEM, st, d.pdl != 0, 1, tmP, tmF, tmDF,
                          tmDP, a));
    return launch_ok("agg_bwd_kernel");
}

}  // namespace vpr
