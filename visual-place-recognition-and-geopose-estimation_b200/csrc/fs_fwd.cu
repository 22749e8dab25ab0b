// SALAD forward, second MLP layer: F = H[:, :h] W1c^T + b1c (bf16) and S = H[:, h:] W1s^T + b1s (fp32 scores).
//
// Both products stream the hidden activation H exactly once (134 MB at B = 256) against a weight matrix
// that is tiny; with generic tiles the weights are re-fetched for every 128-token tile and cost as much
// shared-memory fill traffic as H itself.  Here a CTA owns ONE of the two products for the whole launch:
// its complete weight matrix (W1c: 128 x 512 = 128 KB, W1s: 64 KB) is loaded once and stays resident in
// shared memory, and the CTA only streams H tiles (4-stage TMA ring) -- the launch is bound by the HBM
// read of H.  The token MLP's second layer (TF = HT Wt1^T + bt1, 256 x 256 x 512) rides along on two extra CTAs
// that each keep one 128-row half of Wt1 resident.
//
//   warp 8   : TMA producer (resident weights once, then H k-blocks)
//   warp 9   : tcgen05.mma issuer, accumulators double-buffered in TMEM
//   warps 0-7: epilogue: tcgen05.ld -> + bias -> bf16 / fp32 -> per-warp swizzled staging -> TMA store
#include "common.cuh"
#include "gemm.h"

#include <stdlib.h>

namespace vpr {

int make_map(CUtensorMap* map, const void* ptr, int esize, int64_t inner, int64_t rows, int64_t batch, int64_t ld,
             int64_t bs, int box_inner, int box_rows, const char* what);

static constexpr int FS_THREADS = 320;
static constexpr int FS_W_PROD = 8, FS_W_MMA = 9;
static constexpr int FS_KB = 8;                          // hidden = 512 = 8 k-blocks of 64
static constexpr int FS_A_STAGE = 128 * 128;             // [128 tokens x 64 k] bf16
static constexpr int FS_STAGES_MAX = 8;                  // S CTAs (64 KB of weights) run an 8-deep ring, the others 4
static constexpr int FS_B_BYTES = FS_KB * 128 * 128;     // resident weights, up to [128 n x 512 k] bf16
static constexpr int FS_WBOX = 32 * 128;                 // one warp's 32 rows of an output box
// weights + ring always add up to 192 KB: 128 KB + 4 stages, or 64 KB + 8 stages
static constexpr int FS_SMEM = 1024 + FS_B_BYTES + 4 * FS_A_STAGE + 8 * FS_WBOX + 256;
static_assert(FS_SMEM <= 227 * 1024, "fs_fwd shared memory budget");

struct FsArgs {
    int tiles_m, T, hidden;
    int ctas_f, ctas_s;          // CTAs [0, ctas_f) compute F, the next ctas_s compute S, the last two TF
    int tiles_tok;               // 128-row tiles of the token batch
    int rev;                     // walk the token tiles last-to-first: the rows the H launch wrote last are still in L2
    int l2_hints;                // H loads evict-first, F / S stores evict-last (VPR_FS_L2HINT=0 switches them off)
    const float* bias_f; const float* bias_s; const float* bias_t;
};

__global__ void __launch_bounds__(FS_THREADS, 1)
fs_fwd_kernel(const __grid_constant__ CUtensorMap tmH, const __grid_constant__ CUtensorMap tmW1c,
              const __grid_constant__ CUtensorMap tmW1s, const __grid_constant__ CUtensorMap tmF,
              const __grid_constant__ CUtensorMap tmS, const __grid_constant__ CUtensorMap tmHT,
              const __grid_constant__ CUtensorMap tmWt1, const __grid_constant__ CUtensorMap tmTF, const FsArgs args) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t sB = base;
    // warp index through a shuffle: provably warp-uniform (see gemm.cu: single-thread instructions then take uniform-register
    // operands directly instead of an election loop each)
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    // what this CTA computes: 0 = F (cluster features), 1 = S (scores), 2 = TF (token features, half `thalf`)
    const int cta = (int)blockIdx.x;
    const int kind = cta < args.ctas_f ? 0 : cta < args.ctas_f + args.ctas_s ? 1 : 2;
    const int thalf = kind == 2 ? cta - args.ctas_f - args.ctas_s : 0;
    const int j0 = kind == 0 ? cta : kind == 1 ? cta - args.ctas_f : 0;
    const int nj = kind == 0 ? args.ctas_f : kind == 1 ? args.ctas_s : 1;
    const int tiles_m = kind == 2 ? args.tiles_tok : args.tiles_m;
    const int nrows = kind == 1 ? 64 : 128;              // output columns of this CTA's product
    const int koff = kind == 1 ? args.hidden : 0;        // its half of H
    const int nstages = kind == 1 ? 8 : 4;
    const CUtensorMap* tmA = kind == 2 ? &tmHT : &tmH;
    const CUtensorMap* tmB = kind == 0 ? &tmW1c : kind == 1 ? &tmW1s : &tmWt1;
    const CUtensorMap* tmO = kind == 0 ? &tmF : kind == 1 ? &tmS : &tmTF;
    const uint32_t sA = sB + (uint32_t)(FS_KB * nrows * 128);
    const uint32_t sStage = sB + FS_B_BYTES + 4 * FS_A_STAGE;
    const uint32_t bar = sStage + 8 * FS_WBOX;
    auto afull = [&](int s) { return bar + 8u * s; };
    auto aempty = [&](int s) { return bar + 8u * (FS_STAGES_MAX + s); };
    auto tfull = [&](int s) { return bar + 8u * (2 * FS_STAGES_MAX + s); };
    auto tempty = [&](int s) { return bar + 8u * (2 * FS_STAGES_MAX + 2 + s); };
    const uint32_t bfull = bar + 8u * (2 * FS_STAGES_MAX + 4);
    const uint32_t tmem_slot = bar + 8u * (2 * FS_STAGES_MAX + 5);
    uint8_t* smem_al = smem_raw + (base - smem_u32(smem_raw));
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_al + (tmem_slot - base));

    if (warp == FS_W_PROD && lane == 0) { tma_prefetch_desc(tmA); tma_prefetch_desc(tmB); tma_prefetch_desc(tmO); }
    if (warp == FS_W_MMA) {
        if (lane == 0) {
            for (int s = 0; s < FS_STAGES_MAX; ++s) { mbar_init(afull(s), 1); mbar_init(aempty(s), 1); }
            for (int s = 0; s < 2; ++s) { mbar_init(tfull(s), 1); mbar_init(tempty(s), 8); }
            mbar_init(bfull, 1);
            mbar_fence_init();
        }
        __syncwarp();
        tmem_alloc(tmem_slot, 256);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    // Programmatic dependent launch: everything above overlapped the previous launch's tail; the resident weights come
    // from the parameter pack (two launches back), so the producer fetches them before it waits for H to be complete.
    if (warp != FS_W_PROD) { pdl_wait(); pdl_launch_dependents(); }
    const bool rev = args.rev != 0 && kind != 2;
    auto row_tile = [&](int mt) { return rev ? tiles_m - 1 - mt : mt; };

    if (warp == FS_W_PROD) {
        {   // whole warp, one elected lane issues
            // resident weights, K-major: k-block kb = [nrows x 64 k] at sB + kb * nrows * 128
            if (elect_one()) {
                mbar_expect_tx(bfull, (uint32_t)(FS_KB * nrows * 128));
                for (int kb = 0; kb < FS_KB; ++kb)
                    tma_load_3d(sB + (uint32_t)(kb * nrows * 128), tmB, bfull, kb * 64, thalf * 128, 0);
            }
            __syncwarp();
            pdl_wait();
            pdl_launch_dependents();
            int stage = 0; uint32_t phase = 0;
            // H is read exactly once here (the backward's pass over it is a whole forward tail away): evict-first, so
            // that the 134 MB stream does not push F and S -- which the next two launches read back -- out of L2
            const uint64_t pol_h = l2_evict_first_policy();
            const bool hint = args.l2_hints != 0;
            for (int mt = j0; mt < tiles_m; mt += nj)
                for (int kb = 0; kb < FS_KB; ++kb) {
                    mbar_wait(aempty(stage), phase ^ 1u);
                    if (elect_one()) {
                        mbar_expect_tx(afull(stage), FS_A_STAGE);
                        if (hint) tma_load_3d_hint(sA + (uint32_t)stage * FS_A_STAGE, tmA, afull(stage), koff + kb * 64, row_tile(mt) * 128, 0, pol_h);
                        else tma_load_3d(sA + (uint32_t)stage * FS_A_STAGE, tmA, afull(stage), koff + kb * 64, row_tile(mt) * 128, 0);
                    }
                    if (++stage == nstages) { stage = 0; phase ^= 1u; }
                }
        }
    } else if (warp == FS_W_MMA) {
        {   // whole warp waits, one elected lane issues
            const uint32_t idesc = make_idesc_bf16(128, nrows, false, false);
            mbar_wait(bfull, 0);
            tc_fence_after();
            int stage = 0; uint32_t phase = 0; int it = 0;
            for (int mt = j0; mt < tiles_m; mt += nj, ++it) {
                const int as = it & 1;
                mbar_wait(tempty(as), ((uint32_t)(it >> 1) & 1u) ^ 1u);
                tc_fence_after();
                for (int kb = 0; kb < FS_KB; ++kb) {
                    mbar_wait(afull(stage), phase);
                    tc_fence_after();
                    const uint32_t a_s = sA + (uint32_t)stage * FS_A_STAGE;
                    const uint32_t b_s = sB + (uint32_t)(kb * nrows * 128);
                    if (elect_one()) {
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks)
                            tc_mma_bf16(tmem_base + (uint32_t)(as * 128), make_smem_desc(a_s + ks * 32, 16u, 1024u),
                                        make_smem_desc(b_s + ks * 32, 16u, 1024u), idesc, (kb > 0 || ks > 0) ? 1u : 0u);
                        tc_commit(aempty(stage));
                    }
                    if (++stage == nstages) { stage = 0; phase ^= 1u; }
                }
                if (elect_one()) tc_commit(tfull(as));
            }
        }
    } else {
        const int ew = warp, q = warp & 3, eg = ew >> 2;
        const uint32_t sbuf = sStage + (uint32_t)ew * FS_WBOX;
        const float* bias = kind == 0 ? args.bias_f : kind == 1 ? args.bias_s : args.bias_t + thalf * 128;
        const bool bf = kind == 0;                       // F is stored in bf16, S and TF in fp32
        const uint64_t pol_out = l2_evict_last_policy();  // the Sinkhorn and aggregation launches read F and S right back
        const int box_cols = bf ? 64 : 32;
        const int nbox = nrows / box_cols;               // F: 2, S: 2, TF half: 4
        int it = 0;
        for (int mt = j0; mt < tiles_m; mt += nj, ++it) {
            const int as = it & 1;
            mbar_wait(tfull(as), (uint32_t)(it >> 1) & 1u);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * 128);
            for (int bx = eg; bx < nbox; bx += 2) {      // the two epilogue groups take alternate boxes
                if (lane == 0) bulk_wait_read<0>();
                __syncwarp();
                if (bf) {
#pragma unroll 1
                    for (int c = 0; c < 2; ++c) {
                        uint32_t r[32];
                        tmem_ld32(taddr + bx * 64 + c * 32, r);
                        float bv[32];
#pragma unroll
                        for (int jq = 0; jq < 8; ++jq) {
                            const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias + bx * 64 + c * 32) + jq);
                            bv[4 * jq] = b4.x; bv[4 * jq + 1] = b4.y; bv[4 * jq + 2] = b4.z; bv[4 * jq + 3] = b4.w;
                        }
                        tmem_ld_wait();
#pragma unroll
                        for (int jq = 0; jq < 4; ++jq) {
                            uint32_t pk[4];
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                __nv_bfloat162 h2 = __floats2bfloat162_rn(__uint_as_float(r[jq * 8 + 2 * i]) + bv[jq * 8 + 2 * i],
                                                                          __uint_as_float(r[jq * 8 + 2 * i + 1]) + bv[jq * 8 + 2 * i + 1]);
                                pk[i] = *reinterpret_cast<uint32_t*>(&h2);
                            }
                            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sbuf + (uint32_t)lane * 128u + (uint32_t)(((c * 4 + jq) ^ (lane & 7)) << 4)),
                                         "r"(pk[0]), "r"(pk[1]), "r"(pk[2]), "r"(pk[3]) : "memory");
                        }
                    }
                } else {
                    uint32_t r[32];
                    tmem_ld32(taddr + bx * 32, r);
                    float bv[32];
#pragma unroll
                    for (int jq = 0; jq < 8; ++jq) {
                        const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias + bx * 32) + jq);
                        bv[4 * jq] = b4.x; bv[4 * jq + 1] = b4.y; bv[4 * jq + 2] = b4.z; bv[4 * jq + 3] = b4.w;
                    }
                    tmem_ld_wait();
#pragma unroll
                    for (int jq = 0; jq < 8; ++jq)
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sbuf + (uint32_t)lane * 128u + (uint32_t)((jq ^ (lane & 7)) << 4)),
                                     "r"(__float_as_uint(__uint_as_float(r[4 * jq]) + bv[4 * jq])),
                                     "r"(__float_as_uint(__uint_as_float(r[4 * jq + 1]) + bv[4 * jq + 1])),
                                     "r"(__float_as_uint(__uint_as_float(r[4 * jq + 2]) + bv[4 * jq + 2])),
                                     "r"(__float_as_uint(__uint_as_float(r[4 * jq + 3]) + bv[4 * jq + 3])) : "memory");
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) {
                    if (args.l2_hints != 0) tma_store_3d_hint(tmO, sbuf, thalf * 128 + bx * box_cols, row_tile(mt) * 128 + q * 32, 0, pol_out);
                    else tma_store_3d(tmO, sbuf, thalf * 128 + bx * box_cols, row_tile(mt) * 128 + q * 32, 0);
                    bulk_commit();
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty(as));
        }
        if (lane == 0) bulk_wait_read<0>();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == FS_W_MMA) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 256);
    }
}

bool fs_fwd_supported(int dtype, int l, int m, int hidden) {
    static int enabled = -1;
    if (enabled < 0) {
        const char* e = getenv("VPR_FS_FUSED");
        enabled = (e != nullptr && e[0] == '0') ? 0 : 1;
    }
    return enabled == 1 && dtype == DT_BF16 && l == 128 && m == 64 && hidden == 64 * FS_KB;
}

int fs_fwd_run(const FsFwdDesc& d, cudaStream_t st) {
    VPR_REQUIRE(d.T > 0 && d.T < (1ll << 31), "bad token count");
    VPR_REQUIRE(d.g == 256, "token_dim must be 256 for the fused second-layer launch");
    const int h = d.hidden;
    CUtensorMap tmH, tmW1c, tmW1s, tmF, tmS, tmHT, tmWt1, tmTF;
    VPR_TRY(make_map(&tmH, d.H, 2, 2 * h, d.T, 1, 2 * h, 0, 64, 128, "H"));
    VPR_TRY(make_map(&tmW1c, d.W1c, 2, h, 128, 1, h, 0, 64, 128, "W1c"));
    VPR_TRY(make_map(&tmW1s, d.W1s, 2, h, 64, 1, h, 0, 64, 64, "W1s"));
    VPR_TRY(make_map(&tmF, d.F, 2, 128, d.T, 1, 128, 0, 64, 32, "F"));
    VPR_TRY(make_map(&tmS, d.S, 4, 64, d.T, 1, 64, 0, 32, 32, "S"));
    VPR_TRY(make_map(&tmHT, d.HT, 2, h, d.B, 1, h, 0, 64, 128, "HT"));
    VPR_TRY(make_map(&tmWt1, d.Wt1, 2, h, d.g, 1, h, 0, 64, 128, "Wt1"));
    VPR_TRY(make_map(&tmTF, d.TF, 4, d.g, d.B, 1, d.g, 0, 32, 32, "TF"));
    FsArgs a;
    a.tiles_m = (int)((d.T + 127) / 128);
    a.tiles_tok = (d.B + 127) / 128;
    a.T = (int)d.T;
    a.hidden = h;
    a.bias_f = d.bias_f; a.bias_s = d.bias_s; a.bias_t = d.bias_t;
    {
        static int rev = -1;
        if (rev < 0) { const char* e = getenv("VPR_FS_REVERSE"); rev = (e != nullptr && e[0] == '0') ? 0 : 1; }
        a.rev = rev;
        static int hints = -1;
        if (hints < 0) { const char* e = getenv("VPR_FS_L2HINT"); hints = (e != nullptr && e[0] == '0') ? 0 : 1; }
        a.l2_hints = hints;
    }
    int grid = num_sms();
    if (grid > 2 * a.tiles_m + 2) grid = 2 * a.tiles_m + 2;
    if (grid < 4) grid = 4;
    // both big products stream the same H bytes per tile, but the S CTAs run a ring twice as deep: 60 / 40 split
    a.ctas_f = ((grid - 2) * 3 + 2) / 5;
    a.ctas_s = grid - 2 - a.ctas_f;
    VPR_SMEM_OK(fs_fwd_kernel, FS_SMEM);
    VPR_CUDA_OK(launch_ex(fs_fwd_kernel, dim3((unsigned)grid), dim3(FS_THREADS), (size_t)FS_SMEM, st, d.pdl != 0, 1, tmH, tmW1c,
                          tmW1s, tmF, tmS, tmHT, tmWt1, tmTF, a));
    return launch_ok("fs_fwd_kernel");
}

}  // namespace vpr
