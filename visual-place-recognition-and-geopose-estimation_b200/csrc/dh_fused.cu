// SALAD backward, second MLP layer -- three results from ONE pass over the saved activation H:
//
//   dH  = [dF W1c | dS W1s] * (H > 0) * keep_scale                 (128 x 128 tiles, written by TMA)
//   dW1 = [dF^T H_c ; dS^T H_s]            second-layer weight gradients   (tensor core, TMEM-resident)
//   db0 = column sums of dH                first-layer bias gradients      (tensor core, TMEM-resident)
//   db1 = column sums of dF                cluster_features.3.bias gradient (same trick, first column slice only)
//
// The generic path needs two launches for this (a split-K GEMM that streams H as an operand and a
// masked GEMM that streams H again as the ReLU/dropout mask, with a register butterfly for the column
// sums).  Here every CTA owns ONE 128-column slice of the hidden layer for the whole launch and walks
// token tiles: the W1 slice stays resident in shared memory, and the tiles a token tile needs anyway --
// its dF/dS rows (A operand of the dH MMA) and its H tile (the mask) -- are reused, through MN-major
// shared-memory descriptors over the very same bytes, as the operands of the dW1 MMA
// (dW1[l, h] += sum_t dF[t, l] H[t, h]).  The staged bf16 dH tile, before its TMA store, is the A operand
// of a third MMA against a constant all-ones B (N = 16) that accumulates the column sums.  dW1 and db0
// live in TMEM for the whole launch and are flushed once with fp32 reductions.
//
//   warp 8   : TMA producer (resident W1 slice once; per tile dF/dS rows into a 2-stage ring, H tile x2)
//   warp 9   : tcgen05.mma issuer: dH MMA -> dW1 MMA (same stage) -> column-sum MMA of the previous tile
//   warps 0-7: epilogue: tcgen05.ld -> scale -> bf16 -> mask (H tile in smem) -> swizzled staging -> TMA store
#include "common.cuh"
#include "gemm.h"

#include <stdlib.h>

namespace vpr {

int make_map(CUtensorMap* map, const void* ptr, int esize, int64_t inner, int64_t rows, int64_t batch, int64_t ld,
             int64_t bs, int box_inner, int box_rows, const char* what);

static constexpr int DH_THREADS = 320;
static constexpr int DH_W_PROD = 8, DH_W_MMA = 9;     // above the epilogue warps: the arbiter favours high warp ids
static constexpr int DH_BOX = 128 * 128;                 // 128 rows x 128 B swizzled box (16 KB)
static constexpr int DH_A_STAGE = 2 * DH_BOX;            // [128 tokens x 128 l] bf16 (score half uses one box)
static constexpr int DH_A_STAGES = 2;
static constexpr int DH_B_BYTES = 2 * DH_BOX;            // resident W1 slice: [K <= 128][128 h] bf16
static constexpr int DH_MASK_BYTES = 2 * DH_BOX;         // H tile: [128 tokens x 128 h] bf16
static constexpr int DH_ONES_BYTES = 2048;               // [16 n][64 k] bf16, all ones (every k-step re-reads it)
static constexpr int DH_STAGE_BYTES = 2 * DH_BOX;        // one staged dH tile: two boxes of 64 columns
// the staged tile is double buffered: the epilogue of tile i+1 writes one buffer while the TMA store and the
// column-sum MMA of tile i still read the other (with one buffer the epilogues ran strictly one after the other, each
// behind the previous tile's store + MMA drain: 2.35 us per tile for 0.5 us of epilogue work).  The dynamic shared
// memory base is declared 1024-byte aligned (checked at run time), which pays for the second buffer.
static constexpr int DH_SMEM = DH_B_BYTES + DH_A_STAGES * DH_A_STAGE + 2 * DH_MASK_BYTES + 2 * DH_STAGE_BYTES +
                               DH_ONES_BYTES + 256;
static_assert(DH_SMEM <= 227 * 1024, "dh_fused shared memory budget");
// TMEM columns: dH accumulators 2 x 128, dW1 128, column sums of dH 16, column sums of dF 16
static constexpr uint32_t DH_TM_D2 = 256, DH_TM_D3 = 384, DH_TM_D4 = 400, DH_TM_COLS = 512;

struct DhArgs {
    int tiles_m;             // token tiles (128 tokens each)
    int ct_per_half;         // hidden / 128
    int hidden;
    int T;
    float mask_scale;
    int l2_hints;            // H loads evict-first (its last use in the step), dH stores evict-last (the next two launches read it)
    float* dW1c; float* dW1s; float* db0c; float* db0s; float* db1c;
};

__global__ void __launch_bounds__(DH_THREADS, 1)
dh_fused_kernel(const __grid_constant__ CUtensorMap tmF, const __grid_constant__ CUtensorMap tmS,
                const __grid_constant__ CUtensorMap tmW1c, const __grid_constant__ CUtensorMap tmW1s,
                const __grid_constant__ CUtensorMap tmH, const __grid_constant__ CUtensorMap tmO, const DhArgs args) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t base = smem_u32(smem_raw);
    if ((base & 1023u) != 0u) __trap();                        // SWIZZLE_128B atoms need the 1024-byte alignment
    const uint32_t sB = base;
    const uint32_t sA = sB + DH_B_BYTES;
    const uint32_t sM = sA + DH_A_STAGES * DH_A_STAGE;
    const uint32_t sStage = sM + 2 * DH_MASK_BYTES;            // 2 buffers x two adjacent boxes: columns 0-63, 64-127
    const uint32_t sOnes = sStage + 2 * DH_STAGE_BYTES;
    const uint32_t bar = sOnes + DH_ONES_BYTES;
    auto afull = [&](int s) { return bar + 8u * s; };
    auto aempty = [&](int s) { return bar + 8u * (2 + s); };
    auto mfull = [&](int s) { return bar + 8u * (4 + s); };
    auto mempty = [&](int s) { return bar + 8u * (6 + s); };
    auto tfull = [&](int s) { return bar + 8u * (8 + s); };
    auto tempty = [&](int s) { return bar + 8u * (10 + s); };
    auto sfull = [&](int s) { return bar + 8u * (12 + s); };
    auto sdone = [&](int s) { return bar + 8u * (14 + s); };
    const uint32_t bfull = bar + 8u * 16, accdone = bar + 8u * 17;
    const uint32_t tmem_slot = bar + 8u * 18;
    uint8_t* smem_al = smem_raw + (base - smem_u32(smem_raw));
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_al + (tmem_slot - base));

    // warp index through a shuffle: provably warp-uniform, so the producer and MMA warps run their loops convergently with
    // uniform-register operands and one elected lane issuing (see gemm.cu)
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    // this CTA's hidden-layer column slice and its share of the token tiles
    const int nct = 2 * args.ct_per_half;
    const int ct = (int)blockIdx.x % nct;
    const int half = ct / args.ct_per_half;                  // 0 = cluster MLP, 1 = score MLP
    const int cn = ct - half * args.ct_per_half;
    const int col0 = half * args.hidden + cn * 128;          // column of H / dH
    const int j0 = (int)blockIdx.x / nct;
    const int nj = ((int)gridDim.x - ct + nct - 1) / nct;    // CTAs that share this column slice
    const int kboxes = half == 0 ? 2 : 1;                    // K = 128 (l) or 64 (m)

    if (warp == DH_W_PROD && lane == 0) {
        tma_prefetch_desc(&tmF); tma_prefetch_desc(&tmS); tma_prefetch_desc(&tmW1c); tma_prefetch_desc(&tmW1s);
        tma_prefetch_desc(&tmH); tma_prefetch_desc(&tmO);
    }
    if (warp == DH_W_MMA) {
        if (lane == 0) {
            for (int s = 0; s < 2; ++s) {
                mbar_init(afull(s), 1); mbar_init(aempty(s), 1);
                mbar_init(mfull(s), 1); mbar_init(mempty(s), 9);      // 8 epilogue warps + the dW1 MMA's commit
                mbar_init(tfull(s), 1); mbar_init(tempty(s), 8);
            }
            for (int s = 0; s < 2; ++s) { mbar_init(sfull(s), 8); mbar_init(sdone(s), 1); }
            mbar_init(bfull, 1); mbar_init(accdone, 1);
            mbar_fence_init();
        }
        __syncwarp();
        tmem_alloc(tmem_slot, DH_TM_COLS);
        tmem_relinquish();
    }
    // the constant all-ones operand of the column-sum MMA (any layout of ones is ones)
    for (int i = threadIdx.x; i < DH_ONES_BYTES / 4; i += DH_THREADS)
        reinterpret_cast<uint32_t*>(smem_al + (sOnes - base))[i] = 0x3f803f80u;
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    // Programmatic dependent launch: everything above overlapped the previous launch's tail.  The resident W1 slice
    // comes from the parameter pack of the forward pass (many launches back), so the producer fetches it before it
    // waits for the previous launch (the Sinkhorn backward, which produces dS) to complete.
    if (warp != DH_W_PROD) { pdl_wait(); pdl_launch_dependents(); }

    if (warp == DH_W_PROD) {
        // ------------------------------ TMA producer ------------------------------
        {   // whole warp, one elected lane issues
            const CUtensorMap* tmA = half == 0 ? &tmF : &tmS;
            const CUtensorMap* tmB = half == 0 ? &tmW1c : &tmW1s;
            // resident W1 slice, MN-major: per 64-row k-block two boxes of 64 h-columns
            if (elect_one()) {
                mbar_expect_tx(bfull, (uint32_t)kboxes * DH_BOX);
                for (int kb = 0; kb < kboxes; ++kb)
                    for (int i = 0; i < 2; ++i)
                        tma_load_3d(sB + (uint32_t)kb * DH_BOX + (uint32_t)i * 8192u, tmB, bfull, cn * 128 + i * 64, kb * 64, 0);
            }
            __syncwarp();
            pdl_wait();
            pdl_launch_dependents();
            const uint64_t pol_h = l2_evict_first_policy();
            int it = 0;
            for (int mt = j0; mt < args.tiles_m; mt += nj, ++it) {
                const int s = it & 1;
                const uint32_t ph = (uint32_t)(it >> 1) & 1u;
                mbar_wait(aempty(s), ph ^ 1u);
                if (elect_one()) {
                    mbar_expect_tx(afull(s), (uint32_t)kboxes * DH_BOX);
                    for (int kb = 0; kb < kboxes; ++kb)
                        tma_load_3d(sA + (uint32_t)s * DH_A_STAGE + (uint32_t)kb * DH_BOX, tmA, afull(s), kb * 64, mt * 128, 0);
                }
                mbar_wait(mempty(s), ph ^ 1u);
                if (elect_one()) {
                    mbar_expect_tx(mfull(s), DH_MASK_BYTES);
                    for (int i = 0; i < 2; ++i)
                        if (args.l2_hints != 0)
                            tma_load_3d_hint(sM + (uint32_t)s * DH_MASK_BYTES + (uint32_t)i * DH_BOX, &tmH, mfull(s), col0 + i * 64, mt * 128, 0,
                                             pol_h);
                        else
                            tma_load_3d(sM + (uint32_t)s * DH_MASK_BYTES + (uint32_t)i * DH_BOX, &tmH, mfull(s), col0 + i * 64, mt * 128, 0);
                }
            }
        }
    } else if (warp == DH_W_MMA) {
        // ------------------------------ MMA issuer ------------------------------
        {   // whole warp waits, one elected lane issues
            constexpr uint32_t idesc1 = make_idesc_bf16(128, 128, false, true);    // dH: A K-major, B MN-major
            constexpr uint32_t idesc2 = make_idesc_bf16(128, 128, true, true);     // dW1: both MN-major views
            constexpr uint32_t idesc3 = make_idesc_bf16(128, 16, true, false);     // column sums: A MN-major, ones K-major
            const bool do_db1 = half == 0 && cn == 0 && args.db1c != nullptr;
            mbar_wait(bfull, 0);
            tc_fence_after();
            int it = 0;
            auto colsum_mma = [&](int tile_it) {          // staged dH tile of iteration tile_it -> D3
                const int sb = tile_it & 1;
                mbar_wait(sfull(sb), (uint32_t)(tile_it >> 1) & 1u);
                tc_fence_after();
                if (elect_one()) {
#pragma unroll
                    for (int ks = 0; ks < 8; ++ks) {
                        const uint64_t ad = make_smem_desc(sStage + (uint32_t)sb * DH_STAGE_BYTES + ks * 2048, DH_BOX, 1024u);
                        const uint64_t bd = make_smem_desc(sOnes + (ks & 3) * 32, 16u, 1024u);
                        tc_mma_bf16(tmem_base + DH_TM_D3, ad, bd, idesc3, (tile_it > 0 || ks > 0) ? 1u : 0u);
                    }
                    tc_commit(sdone(sb));
                }
            };
            for (int mt = j0; mt < args.tiles_m; mt += nj, ++it) {
                const int s = it & 1;
                const uint32_t ph = (uint32_t)(it >> 1) & 1u;
                const uint32_t a_s = sA + (uint32_t)s * DH_A_STAGE;
                mbar_wait(tempty(s), ph ^ 1u);
                mbar_wait(afull(s), ph);
                tc_fence_after();
                // dH tile: D1[token, h] = sum_k A[token, k] W1[k, h]
                if (elect_one()) {
                    for (int kb = 0; kb < kboxes; ++kb)
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks) {
                            const uint64_t ad = make_smem_desc(a_s + (uint32_t)kb * DH_BOX + ks * 32, 16u, 1024u);
                            const uint64_t bd = make_smem_desc(sB + (uint32_t)kb * DH_BOX + ks * 2048, 8192u, 1024u);
                            tc_mma_bf16(tmem_base + (uint32_t)(s * 128), ad, bd, idesc1, (kb > 0 || ks > 0) ? 1u : 0u);
                        }
                    tc_commit(tfull(s));
                }
                // dW1 slice: D2[l, h] += sum_token A[token, l] H[token, h]   (same bytes, MN-major views)
                mbar_wait(mfull(s), ph);
                tc_fence_after();
                const uint32_t m_s = sM + (uint32_t)s * DH_MASK_BYTES;
                if (elect_one()) {
#pragma unroll
                    for (int ks = 0; ks < 8; ++ks) {
                        const uint64_t ad = make_smem_desc(a_s + ks * 2048, DH_BOX, 1024u);
                        const uint64_t bd = make_smem_desc(m_s + ks * 2048, DH_BOX, 1024u);
                        tc_mma_bf16(tmem_base + DH_TM_D2, ad, bd, idesc2, (it > 0 || ks > 0) ? 1u : 0u);
                    }
                    if (do_db1) {                              // D4[l, .] += sum_token dF[token, l]
#pragma unroll
                        for (int ks = 0; ks < 8; ++ks) {
                            const uint64_t ad = make_smem_desc(a_s + ks * 2048, DH_BOX, 1024u);
                            const uint64_t bd = make_smem_desc(sOnes + (ks & 3) * 32, 16u, 1024u);
                            tc_mma_bf16(tmem_base + DH_TM_D4, ad, bd, idesc3, (it > 0 || ks > 0) ? 1u : 0u);
                        }
                    }
                    tc_commit(aempty(s));
                    tc_commit(mempty(s));
                }
                if (it > 0) colsum_mma(it - 1);
            }
            if (it > 0) colsum_mma(it - 1);
            if (elect_one()) tc_commit(accdone);
        }
    } else {
        // ------------------------------ epilogue ------------------------------
        const int ew = warp;
        const int q = warp & 3;
        const int eg = ew >> 2;                      // group 0: columns 0-63, group 1: columns 64-127 of the tile
        const bool issuer = ((ew & 3) == 0) && lane == 0;
        const int row = q * 32 + lane;
        const float ms = args.mask_scale;
        int it = 0;
        for (int mt = j0; mt < args.tiles_m; mt += nj, ++it) {
            const int s = it & 1;
            const uint32_t ph = (uint32_t)(it >> 1) & 1u;
            const uint32_t sbox = sStage + (uint32_t)s * DH_STAGE_BYTES + (uint32_t)eg * DH_BOX;
            mbar_wait(tfull(s), ph);
            mbar_wait(mfull(s), ph);
            if (it > 1) mbar_wait(sdone(s), ph ^ 1u);                    // the column-sum MMA of tile it - 2 has read this buffer
            if (issuer) bulk_wait_read<1>();                             // ... and so has its TMA store (tile it - 1's may be in flight)
            named_bar_sync(1 + eg, 128);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(s * 128 + eg * 64);
            const uint32_t mbox = sM + (uint32_t)s * DH_MASK_BYTES + (uint32_t)eg * DH_BOX;
            const __nv_bfloat162 zero2 = __floats2bfloat162_rn(0.f, 0.f);
#pragma unroll 1
            for (int c = 0; c < 2; ++c) {
                uint32_t r[32];
                tmem_ld32(taddr + c * 32, r);
                tmem_ld_wait();
                uint32_t pk[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    __nv_bfloat162 h2 = __floats2bfloat162_rn(__uint_as_float(r[2 * i]) * ms, __uint_as_float(r[2 * i + 1]) * ms);
                    pk[i] = *reinterpret_cast<uint32_t*>(&h2);
                }
#pragma unroll
                for (int jq = 0; jq < 4; ++jq) {
                    const uint32_t off = (uint32_t)row * 128u + (uint32_t)(((c * 4 + jq) ^ (row & 7)) << 4);
                    uint32_t mw[4];
                    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                                 : "=r"(mw[0]), "=r"(mw[1]), "=r"(mw[2]), "=r"(mw[3]) : "r"(mbox + off) : "memory");
#pragma unroll
                    for (int t = 0; t < 4; ++t)
                        pk[4 * jq + t] &= __hgt2_mask(*reinterpret_cast<__nv_bfloat162*>(&mw[t]), zero2);
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sbox + off), "r"(pk[4 * jq]),
                                 "r"(pk[4 * jq + 1]), "r"(pk[4 * jq + 2]), "r"(pk[4 * jq + 3]) : "memory");
                }
            }
            // accumulator and mask tile are free again; the staged tile is ready for the store and the column-sum MMA
            tc_fence_before();
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(tempty(s));
                mbar_arrive(mempty(s));
                mbar_arrive(sfull(s));
            }
            named_bar_sync(1 + eg, 128);
            if (issuer) {
                if (args.l2_hints != 0) tma_store_3d_hint(&tmO, sbox, col0 + eg * 64, mt * 128, 0, l2_evict_last_policy());
                else tma_store_3d(&tmO, sbox, col0 + eg * 64, mt * 128, 0);
                bulk_commit();
            }
        }
        // ---- flush the launch-long accumulators: dW1 slice (fp32 reductions) and the column sums ----
        mbar_wait(accdone, 0);
        tc_fence_after();
        if (it > 0) {
            float* dW1 = half == 0 ? args.dW1c : args.dW1s;
            const int lrows = half == 0 ? 128 : 64;
#pragma unroll 1
            for (int c = eg; c < 4; c += 2) {
                uint32_t r[32];
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + DH_TM_D2 + c * 32, r);
                tmem_ld_wait();
                if (row < lrows) {
                    float* op = dW1 + (size_t)row * args.hidden + cn * 128 + c * 32;
#pragma unroll
                    for (int jq = 0; jq < 8; ++jq)
                        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(op + 4 * jq),
                                     "f"(__uint_as_float(r[4 * jq])), "f"(__uint_as_float(r[4 * jq + 1])),
                                     "f"(__uint_as_float(r[4 * jq + 2])), "f"(__uint_as_float(r[4 * jq + 3])) : "memory");
                }
            }
            if (eg == 0) {                               // column sums: TMEM lane = column of the slice
                uint32_t cs;
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];"
                             : "=r"(cs) : "r"(tmem_base + ((uint32_t)(q * 32) << 16) + DH_TM_D3) : "memory");
                tmem_ld_wait();
                float* db = half == 0 ? args.db0c : args.db0s;
                atomicAdd(db + cn * 128 + row, __uint_as_float(cs));
            } else if (half == 0 && cn == 0 && args.db1c != nullptr) {      // TMEM lane = cluster feature l
                uint32_t cs;
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];"
                             : "=r"(cs) : "r"(tmem_base + ((uint32_t)(q * 32) << 16) + DH_TM_D4) : "memory");
                tmem_ld_wait();
                atomicAdd(args.db1c + row, __uint_as_float(cs));
            }
        }
        if (issuer) bulk_wait_read<0>();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == DH_W_MMA) {
        tc_fence_after();
        tmem_dealloc(tmem_base, DH_TM_COLS);
    }
}

bool dh_fused_supported(int dtype, int l, int m, int hidden) {
    static int enabled = -1;
    if (enabled < 0) {
        const char* e = getenv("VPR_DH_FUSED");
        enabled = (e != nullptr && e[0] == '0') ? 0 : 1;
    }
    return enabled == 1 && dtype == DT_BF16 && l == 128 && m == 64 && hidden % 128 == 0 && hidden >= 128;
}

int dh_fused_run(const DhFusedDesc& d, cudaStream_t st) {
    VPR_REQUIRE(d.T > 0 && d.T < (1ll << 31), "bad token count");
    CUtensorMap tmF, tmS, tmW1c, tmW1s, tmH, tmO;
    const int h = d.hidden;
    VPR_TRY(make_map(&tmF, d.dF, 2, 128, d.T, 1, 128, 0, 64, 128, "dF"));
    VPR_TRY(make_map(&tmS, d.dS, 2, 64, d.T, 1, 64, 0, 64, 128, "dS"));
    VPR_TRY(make_map(&tmW1c, d.W1c, 2, h, 128, 1, h, 0, 64, 64, "W1c"));
    VPR_TRY(make_map(&tmW1s, d.W1s, 2, h, 64, 1, h, 0, 64, 64, "W1s"));
    VPR_TRY(make_map(&tmH, d.H, 2, 2 * h, d.T, 1, 2 * h, 0, 64, 128, "H"));
    VPR_TRY(make_map(&tmO, d.dH, 2, 2 * h, d.T, 1, 2 * h, 0, 64, 128, "dH"));
    DhArgs a;
    a.tiles_m = (int)((d.T + 127) / 128);
    a.ct_per_half = h / 128;
    a.hidden = h;
    a.T = (int)d.T;
    a.mask_scale = d.mask_scale;
    {
        static int hints = -1;
        if (hints < 0) { const char* e = getenv("VPR_DH_L2HINT"); hints = (e != nullptr && e[0] == '0') ? 0 : 1; }
        a.l2_hints = hints;
    }
    a.dW1c = d.dW1c; a.dW1s = d.dW1s; a.db0c = d.db0c; a.db0s = d.db0s; a.db1c = d.db1c;
    const int nct = 2 * a.ct_per_half;
    long long want = (long long)a.tiles_m * nct;
    int reserve = d.reserve_sms < 0 ? 0 : d.reserve_sms;
    if (reserve > num_sms() / 2) reserve = num_sms() / 2;
    int grid = (int)(want < num_sms() - reserve ? want : num_sms() - reserve);
    if (grid < nct) grid = nct;                      // every column slice needs at least one CTA
    VPR_SMEM_OK(dh_fused_kernel, DH_SMEM);
    VPR_CUDA_OK(launch_ex(dh_fused_kernel, dim3((unsigned)grid), dim3(DH_THREADS), (size_t)DH_SMEM, st, d.pdl != 0, 1, tmF, tmS,
                          tmW1c, tmW1s, tmH, tmO, a));
    return launch_ok("dh_fused_kernel");
}

}  // namespace vpr
