// SALAD forward, last step, one CTA per image: soft-assignment aggregation on the tensor core fused with
// the descriptor assembly.
//
//   AGG[l, c] = sum_j F[j, l] P[j, c]                    (128 x 64 x n, tcgen05, accumulator in TMEM)
//   out       = normalize( [ normalize(TF) | normalize(AGG over l per cluster c).flatten ] )
//
// F[b] and P[b] arrive by TMA as 128-token row chunks and are consumed as MN-major operands (M = cluster
// feature l, N = cluster c, K = token j), so the product needs no transposed copy.  The 128 epilogue
// threads each own one feature row l of the 128 x 64 accumulator: they store AGG (fp32, the backward
// needs it), take per-cluster sums of squares over l with a register butterfly, and write their 64
// normalised values as one contiguous 128-byte row of the descriptor.  Replaces a batched GEMM launch
// plus a separate norm kernel (14 us -> one launch).
//
//   warps 0-3: epilogue (one TMEM lane quarter each)     warp 4: TMA producer     warp 5: MMA issuer
#include "common.cuh"
#include "gemm.h"

#include <stdlib.h>

namespace vpr {

int make_map(CUtensorMap* map, const void* ptr, int esize, int64_t inner, int64_t rows, int64_t batch, int64_t ld,
             int64_t bs, int box_inner, int box_rows, const char* what);

static constexpr int AG_THREADS = 192;
static constexpr int AG_W_PROD = 4, AG_W_MMA = 5;
static constexpr int AG_BOX = 128 * 128;                 // 128 token rows x 128 B
static constexpr int AG_STAGE = 3 * AG_BOX;              // F chunk (2 boxes: l 0-63, 64-127) + P chunk (1 box)
static constexpr int AG_STAGES = 2;
static constexpr int AG_SMEM = 1024 + AG_STAGES * AG_STAGE + 256;
#define AG_EPS 1e-12f

struct AgArgs {
    int n, g;
    float* AGG;              // [B, 128, 64] fp32
    const float* TF;         // [B, g] fp32
    bf16* out;               // [B, g + 8192]
};

__device__ __forceinline__ float ag_colsum32(float (&v)[32], int lane) {
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

__global__ void __launch_bounds__(AG_THREADS, 2)
agg_fwd_kernel(const __grid_constant__ CUtensorMap tmF, const __grid_constant__ CUtensorMap tmP, const AgArgs args) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ float part[4][64];            // per-warp partial sums of squares per cluster
    __shared__ float inv_c[64];
    __shared__ float wred[4];
    __shared__ float sc[2];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar = base + AG_STAGES * AG_STAGE;
    auto full = [&](int s) { return bar + 8u * s; };
    auto empty = [&](int s) { return bar + 8u * (AG_STAGES + s); };
    const uint32_t tfull = bar + 8u * (2 * AG_STAGES);
    const uint32_t tmem_slot = bar + 8u * (2 * AG_STAGES + 1);
    uint8_t* smem_al = smem_raw + (base - smem_u32(smem_raw));
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_al + (tmem_slot - base));

    // warp index through a shuffle: provably warp-uniform (see gemm.cu: single-thread instructions then take uniform-register
    // operands directly instead of an election loop each)
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    const int b = blockIdx.x;
    const int chunks = (args.n + 127) / 128;

    if (warp == AG_W_PROD && lane == 0) { tma_prefetch_desc(&tmF); tma_prefetch_desc(&tmP); }
    if (warp == AG_W_MMA) {
        if (lane == 0) {
            for (int s = 0; s < AG_STAGES; ++s) { mbar_init(full(s), 1); mbar_init(empty(s), 1); }
            mbar_init(tfull, 1);
            mbar_fence_init();
        }
        __syncwarp();
        tmem_alloc(tmem_slot, 64);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    pdl_wait();                  // everything above overlapped the previous launch's tail
    pdl_launch_dependents();

    if (warp == AG_W_PROD) {
        for (int ch = 0; ch < chunks; ++ch) {                // whole warp, one elected lane issues
            const int s = ch % AG_STAGES;
            mbar_wait(empty(s), (((uint32_t)(ch / AG_STAGES)) & 1u) ^ 1u);
            const uint32_t st = base + (uint32_t)s * AG_STAGE;
            if (elect_one()) {
                mbar_expect_tx(full(s), AG_STAGE);
                tma_load_3d(st, &tmF, full(s), 0, ch * 128, b);                  // rows beyond n are zero-filled
                tma_load_3d(st + AG_BOX, &tmF, full(s), 64, ch * 128, b);
                tma_load_3d(st + 2 * AG_BOX, &tmP, full(s), 0, ch * 128, b);
            }
        }
    } else if (warp == AG_W_MMA) {
        constexpr uint32_t idesc = make_idesc_bf16(128, 64, true, true);
        for (int ch = 0; ch < chunks; ++ch) {
            const int s = ch % AG_STAGES;
            mbar_wait(full(s), ((uint32_t)(ch / AG_STAGES)) & 1u);
            tc_fence_after();
            const uint32_t st = base + (uint32_t)s * AG_STAGE;
            if (elect_one()) {
#pragma unroll
                for (int ks = 0; ks < 8; ++ks)
                    tc_mma_bf16(tmem_base, make_smem_desc(st + ks * 2048, AG_BOX, 1024u),
                                make_smem_desc(st + 2 * AG_BOX + ks * 2048, AG_BOX, 1024u), idesc, (ch > 0 || ks > 0) ? 1u : 0u);
                tc_commit(empty(s));
            }
        }
        if (elect_one()) tc_commit(tfull);
    } else {
        // ------------------------------ epilogue: thread = cluster-feature row l ------------------------------
        const int l = warp * 32 + lane;
        const int tid = threadIdx.x;             // 0..127
        // token part first (independent of the MMA): ||TF||^2
        const float* t = args.TF + (size_t)b * args.g;
        float ts = 0.f;
        for (int i = tid; i < args.g; i += 128) { const float v = __ldg(t + i); ts = fmaf(v, v, ts); }
        ts = warp_sum(ts);
        if (lane == 0) wred[warp] = ts;
        mbar_wait(tfull, 0);
        tc_fence_after();
        float v[64];
        {
            uint32_t r0[32], r1[32];
            const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
            tmem_ld32(taddr, r0);
            tmem_ld32(taddr + 32, r1);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) { v[i] = __uint_as_float(r0[i]); v[32 + i] = __uint_as_float(r1[i]); }
        }
        float* ag = args.AGG + ((size_t)b * 128 + l) * 64;
#pragma unroll
        for (int i = 0; i < 16; ++i) *reinterpret_cast<float4*>(ag + 4 * i) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
        {   // per-cluster sums of squares over the 32 rows of this warp
            float q0[32], q1[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) { q0[i] = v[i] * v[i]; q1[i] = v[32 + i] * v[32 + i]; }
            part[warp][lane] = ag_colsum32(q0, lane);
            part[warp][32 + lane] = ag_colsum32(q1, lane);
        }
        named_bar_sync(1, 128);
        if (warp == 0) {
            float zz = 0.f;
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const int c = lane + 32 * k;
                const float tot = part[0][c] + part[1][c] + part[2][c] + part[3][c];
                const float nc = fmaxf(sqrtf(tot), AG_EPS);
                inv_c[c] = 1.f / nc;
                zz += tot / (nc * nc);
            }
            zz = warp_sum(zz);
            const float tsum = wred[0] + wred[1] + wred[2] + wred[3];
            const float nt = fmaxf(sqrtf(tsum), AG_EPS);
            zz += tsum / (nt * nt);
            const float invN = 1.f / fmaxf(sqrtf(zz), AG_EPS);
            if (lane == 0) { sc[0] = invN; sc[1] = invN / nt; }
        }
        named_bar_sync(1, 128);
        const float invN = sc[0], invt = sc[1];
        bf16* o = args.out + (size_t)b * (args.g + 8192);
        bf16* orow = o + args.g + (size_t)l * 64;
#pragma unroll
        for (int i = 0; i < 8; ++i) {            // 8 x 16 bytes = this row's 64 normalised values
            uint32_t pk[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int c = 8 * i + 2 * k;
                __nv_bfloat162 h2 = __floats2bfloat162_rn(v[c] * inv_c[c] * invN, v[c + 1] * inv_c[c + 1] * invN);
                pk[k] = *reinterpret_cast<uint32_t*>(&h2);
            }
            *reinterpret_cast<uint4*>(orow + 8 * i) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        }
        for (int i = tid; i < args.g; i += 128) o[i] = __float2bfloat16_rn(__ldg(t + i) * invt);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == AG_W_MMA) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 64);
    }
}

bool agg_fwd_supported(int dtype, int l, int m, int g, int n) {
    static int enabled = -1;
    if (enabled < 0) {
        const char* e = getenv("VPR_AGG_FUSED");
        enabled = (e != nullptr && e[0] == '0') ? 0 : 1;
    }
    return enabled == 1 && dtype == DT_BF16 && l == 128 && m == 64 && g % 8 == 0 && n >= 1;
}

int agg_fwd_run(const AggFwdDesc& d, cudaStream_t st) {
    CUtensorMap tmF, tmP;
    VPR_TRY(make_map(&tmF, d.F, 2, 128, d.n, d.B, 128, (int64_t)d.n * 128, 64, 128, "F"));
    VPR_TRY(make_map(&tmP, d.P, 2, 64, d.n, d.B, 64, (int64_t)d.n * 64, 64, 128, "P"));
    AgArgs a;
    a.n = d.n; a.g = d.g; a.AGG = d.AGG; a.TF = d.TF; a.out = reinterpret_cast<bf16*>(d.out);
    VPR_SMEM_OK(agg_fwd_kernel, AG_SMEM);
    VPR_CUDA_OK(launch_ex(agg_fwd_kernel, dim3((unsigned)d.B), dim3(AG_THREADS), (size_t)AG_SMEM, st, d.pdl != 0, 1, tmF, tmP, a));
    return launch_ok("agg_fwd_kernel");
}

}  // namespace vpr
