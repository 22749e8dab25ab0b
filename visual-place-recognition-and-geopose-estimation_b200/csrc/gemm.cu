// Persistent, warp-specialised, grouped bf16 GEMM for sm_100a.
//   warp 8      : TMA producer (cp.async.bulk.tensor, SWIZZLE_128B boxes into an mbarrier ring; in
//                 mask mode also the relu-mask tile of every output tile, double buffered)
//   warp 9      : TMEM allocator + single-thread tcgen05.mma issuer (fp32 accumulators in TMEM,
//                 double-buffered so the epilogue of tile i overlaps the MMAs of tile i+1)
//   warps 0..7  : epilogue, two groups of four warps (one warp per TMEM lane quarter) that take
//                 alternate 128-byte-wide output boxes: tcgen05.ld -> bias / relu / dropout / relu-mask
//                 / column sums in registers -> the warp's own 32-row SWIZZLE_128B staging box -> its
//                 lane 0 issues the TMA tile store (edges are clipped by the tensor map); no
//                 cross-warp barrier anywhere in the epilogue.  Split-K uses vector fp32 reductions
//                 straight from registers
// Operands may be K-major or MN-major in global memory (all four combinations), batched through
// the third tensor-map dimension and split along K.  One launch runs up to MAX_GROUP independent
// problems of the same operand layout / tile shape (the token-MLP GEMMs ride along with the
// per-patch GEMMs instead of paying their own launch + pipeline fill).
#include "common.cuh"
#include "gemm.h"

#include <mutex>
#include <stdlib.h>
#include <string.h>

namespace vpr {

// ---------------------------------------------------------------------------
// tcgen05 kernel
// ---------------------------------------------------------------------------
static constexpr int BM = 128;
static constexpr int BK = 64;                   // 64 bf16 = 128 B = one swizzle span
static constexpr int A_STAGE_BYTES = BM * BK * 2;
static constexpr int GEMM_THREADS = 320;        // TMA warp, MMA warp, 8 epilogue warps
static constexpr int EPI_WARPS = 8;
// Roles by warp index.  The SM's warp arbiter favours the highest warp id of a scheduler, so the two
// latency-critical single-thread roles sit ABOVE the epilogue warps they share schedulers with.
static constexpr int W_PROD = 8, W_MMA = 9;
static constexpr int BOX_BYTES = 128 * 128;     // 128 rows x 128 B: one swizzled TMA box
static constexpr int WBOX_BYTES = 32 * 128;     // one epilogue warp's 32 rows of an output box
static constexpr int MAX_GROUP = 4;
static constexpr int SMEM_LIMIT = 227 * 1024;

enum : uint32_t { F_BIAS = 1u, F_RELU = 2u, F_DROP = 4u, F_MASK = 8u, F_ALPHA = 16u, F_COLSUM = 32u };

struct TcArgs {
    int M, N, K, batch;
    int tiles_m, tiles_n, k_splits, kb_total;
    int b_batched;
    int work_begin;            // first work item of this problem inside the group
    int out_kind;
    uint32_t flags;
    void* out;                 // split-K (atomic) epilogue only; stored outputs go through tmO
    long long ldo, out_bs;
    const float* bias;
    float* colsum;             // [N] fp32 += column sums of the stored (masked, rounded) output
    float mask_scale, alpha;
    uint32_t drop_thr; float drop_scale;
    unsigned long long drop_seed;
    const unsigned long long* drop_seed_dev;   // optional device-side offset of the seed
};

struct TcProblem {
    CUtensorMap tmA, tmB, tmO, tmM;     // operands, output tile store, relu-mask tile load
    TcArgs args;
};
struct TcGroup {
    TcProblem p[MAX_GROUP];
    int count;
    int total_work;
    int pdl;                 // launch with the programmatic-dependent-launch attribute
};
static_assert(sizeof(TcArgs) % 8 == 0, "TcArgs is copied to shared memory in 4-byte words");
static constexpr int ARGS_SMEM_BYTES = ((MAX_GROUP * (int)sizeof(TcArgs) + 127) / 128) * 128;
static constexpr int MISC_SMEM_BYTES = 256 + ARGS_SMEM_BYTES;        // barriers + tmem slot, arguments

#ifndef VPR_SBUF256
#define VPR_SBUF256 1      // 256-wide tiles: one staging box per epilogue warp (a second one costs an operand stage)
#endif

template <int BN, bool MASK, bool CTA2, bool MCB = false>
struct TcCfg {
    // CTA2: a CTA pair (cta_group::2) owns a 256 x BN tile; each CTA stages its own 128 rows of A and
    // half of the B tile, so a k-block costs 32 KB per SM instead of 48 KB at BN = 256.
    // MCB: a cluster of two CTAs with independent single-CTA MMAs on two 128-row blocks of the same N tile;
    // each CTA fetches half of the B tile and multicasts it into both, so the B tile crosses the L2->SM fabric once
    // per cluster (32 KB of fetches per SM and k-block instead of 48 KB) while every CTA keeps its own accumulators
    // and epilogue pace.
    static constexpr int TILE_M = (CTA2 || MCB) ? 2 * BM : BM;
    static constexpr int B_ROWS = (CTA2 || MCB) ? BN / 2 : BN;          // B-tile rows (n) fetched by one CTA
    static constexpr int B_STAGE_BYTES = (CTA2 ? BN / 2 : BN) * BK * 2; // B-tile bytes resident per CTA
    static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
    static constexpr int MASK_BOXES = BN / 64;                          // bf16 mask tile, 64 columns per box
    static constexpr int MASK_BYTES = MASK ? 2 * MASK_BOXES * BOX_BYTES : 0;
    static constexpr int SBUF = (BN >= 256) ? VPR_SBUF256 : 2;          // staging boxes per epilogue warp
    static constexpr int STAGING_BYTES = EPI_WARPS * SBUF * WBOX_BYTES;
    static constexpr int RING_BUDGET = SMEM_LIMIT - 1024 - MISC_SMEM_BYTES - MASK_BYTES - STAGING_BYTES;
    static constexpr int STAGES = (RING_BUDGET / STAGE_BYTES) > 8 ? 8 : (RING_BUDGET / STAGE_BYTES);
    static constexpr int TMEM_COLS = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128
                                     : (2 * BN <= 256) ? 256 : 512;
    static constexpr int SMEM_BYTES = 1024 + STAGES * STAGE_BYTES + MASK_BYTES + STAGING_BYTES + MISC_SMEM_BYTES;
    static_assert(!(MASK && (CTA2 || MCB)), "the relu-mask epilogue runs single-CTA tiles");
    static_assert(!(CTA2 && MCB), "pair tiles and multicast clusters are alternatives");
    static_assert(STAGES >= 2, "operand ring too small");
    static_assert(2 * STAGES + 8 <= 31, "barrier area overflow");
    static_assert(SMEM_BYTES <= SMEM_LIMIT, "shared memory budget");
};

// Per-thread epilogue math on one 32-column chunk of one output row: alpha, bias, relu, dropout.
// The 32 bias values of a chunk, fetched while the accumulator chunk is still on its way from TMEM.
__device__ __forceinline__ void epilogue_bias32(const TcArgs& a, uint32_t flags, float (&bv)[32], int n0) {
    if (flags & F_BIAS) {
        if (n0 + 32 <= a.N) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float4 bq = __ldg(reinterpret_cast<const float4*>(a.bias + n0) + j);
                bv[4 * j] = bq.x; bv[4 * j + 1] = bq.y; bv[4 * j + 2] = bq.z; bv[4 * j + 3] = bq.w;
            }
        } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) bv[i] = (n0 + i < a.N) ? __ldg(a.bias + n0 + i) : 0.f;
        }
    }
}

__device__ __forceinline__ void epilogue_math32(const TcArgs& a, uint32_t flags, const uint32_t (&r)[32], const float (&bv)[32],
                                                float (&v)[32], int b, int m, int n0) {
    const int N = a.N;
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
    if (flags & F_ALPHA) {
        const float al = a.alpha;
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] *= al;
    }
    if (flags & F_BIAS) {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] += bv[i];
    }
    if (flags & F_RELU) {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
    }
    if (flags & F_DROP) {
        const unsigned long long base = ((unsigned long long)b * a.M + m) * (unsigned long long)N + n0;
        const uint32_t thr = a.drop_thr;
        const float sc = a.drop_scale;
        const int off = (int)(base & 31ull);
        const unsigned long long seed = a.drop_seed + (a.drop_seed_dev != nullptr ? __ldg(a.drop_seed_dev) : 0ull);
        uint32_t w0 = dropout_mix(seed, base >> 5);
        if (off == 0) {                          // this thread owns a whole run: one multiply-add per element
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const uint32_t w = w0 * drop_lcg_mul(i) + drop_lcg_add(i);
                v[i] = (w >= thr) ? v[i] * sc : 0.f;
            }
        } else {                                 // the chunk straddles two runs: step the generator
            const uint32_t w1 = dropout_mix(seed, (base >> 5) + 1);
            for (int k = 0; k < off; ++k) w0 = w0 * DROP_LCG_A + DROP_LCG_C;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                if (off + i == 32) w0 = w1;
                v[i] = (w0 >= thr) ? v[i] * sc : 0.f;
                w0 = w0 * DROP_LCG_A + DROP_LCG_C;
            }
        }
    }
}

// split-K epilogue: vector fp32 reductions straight from registers
__device__ __forceinline__ void epilogue_atomic32(const TcArgs& a, const float (&v)[32], int b, int m, int n0) {
    float* op = reinterpret_cast<float*>(a.out) + (long long)b * a.out_bs + (long long)m * a.ldo + n0;
    if (n0 + 32 <= a.N && ((a.ldo & 3) == 0)) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(op + 4 * j), "f"(v[4 * j]),
                         "f"(v[4 * j + 1]), "f"(v[4 * j + 2]), "f"(v[4 * j + 3])
                         : "memory");
    } else {
#pragma unroll
        for (int i = 0; i < 32; ++i)
            if (n0 + i < a.N) atomicAdd(op + i, v[i]);
    }
}

// Column sums of a 32-row x 32-column register tile (lane = row, v[i] = column i): a butterfly
// reduce-scatter over the warp leaves the total of column `lane` in lane `lane`.
__device__ __forceinline__ float warp_colsum32(float (&v)[32], int lane) {
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

// 16-byte slot `c16` of row `row` inside a 128-row x 128-byte SWIZZLE_128B box
__device__ __forceinline__ uint32_t box_addr(uint32_t box, int row, int c16) {
    return box + (uint32_t)row * 128u + (uint32_t)((c16 ^ (row & 7)) << 4);
}

template <bool A_MN, bool B_MN, int BN, bool MASK, bool CTA2, bool MCB = false>
__global__ void __launch_bounds__(GEMM_THREADS, 1) gemm_tc_kernel(const __grid_constant__ TcGroup grp) {
    using Cfg = TcCfg<BN, MASK, CTA2, MCB>;
    constexpr bool CL2 = CTA2 || MCB;                       // launched as clusters of two CTAs
    constexpr int STAGES = Cfg::STAGES;
    constexpr int SBUF = Cfg::SBUF;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t mask_base = smem_base + STAGES * Cfg::STAGE_BYTES;
    const uint32_t staging = mask_base + Cfg::MASK_BYTES;
    const uint32_t bar_base = staging + Cfg::STAGING_BYTES;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
    auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + s); };
    auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + 2 + s); };
    auto mfull_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + 4 + s); };
    auto mempty_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + 6 + s); };
    const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 8);
    uint8_t* smem_al = smem_raw + (smem_base - smem_u32(smem_raw));
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_al + (tmem_slot - smem_base));
    // per-problem arguments in shared memory: dynamically indexed by every role without touching param space
    TcArgs* sargs = reinterpret_cast<TcArgs*>(smem_al + (bar_base + 256 - smem_base));

    // warp index through a shuffle: the compiler then knows every branch on it is warp-uniform (uniform datapath, no
    // per-lane election loops around the single-thread instructions of the producer and MMA warps)
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;
    const uint32_t cta_rank = CL2 ? cluster_ctarank() : 0u;      // 0 = leader of the pair (issues the MMAs)
    const int w_first = CL2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
    const int w_step = CL2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;

    if (warp == W_PROD && lane == 0) {
        for (int g = 0; g < grp.count; ++g) {
            tma_prefetch_desc(&grp.p[g].tmA);
            tma_prefetch_desc(&grp.p[g].tmB);
            if (grp.p[g].args.out_kind != OUT_ATOMIC_F32) tma_prefetch_desc(&grp.p[g].tmO);
            if (MASK && (grp.p[g].args.flags & F_MASK)) tma_prefetch_desc(&grp.p[g].tmM);
        }
    }
    if (warp == W_MMA) {
        if (lane == 0) {
            for (int s = 0; s < STAGES; ++s) {
                mbar_init(full_bar(s), 1);
                mbar_init(empty_bar(s), MCB ? 2 : 1);        // MCB: both CTAs' MMAs have read what the peer multicast here
            }
            for (int s = 0; s < 2; ++s) {
                mbar_init(tfull_bar(s), 1);
                mbar_init(tempty_bar(s), CTA2 ? 2 * EPI_WARPS : EPI_WARPS);     // the leader's is the one waited on
                mbar_init(mfull_bar(s), 1);
                mbar_init(mempty_bar(s), EPI_WARPS);
            }
            mbar_fence_init();
        }
        __syncwarp();
        if (CTA2) { tmem_alloc_2sm(tmem_slot, Cfg::TMEM_COLS); tmem_relinquish_2sm(); }
        else      { tmem_alloc(tmem_slot, Cfg::TMEM_COLS); tmem_relinquish(); }
    }
    for (int i = threadIdx.x; i < grp.count * (int)(sizeof(TcArgs) / 4); i += GEMM_THREADS) {
        const int gi = i / (int)(sizeof(TcArgs) / 4), wi = i % (int)(sizeof(TcArgs) / 4);
        reinterpret_cast<uint32_t*>(sargs + gi)[wi] = reinterpret_cast<const uint32_t*>(&grp.p[gi].args)[wi];
    }
    tc_fence_before();
    if (CL2) cluster_sync_all();           // the peer's barriers must exist before anything arrives on them
    else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    const int total_work = grp.total_work;
    // programmatic dependent launch: everything above overlapped the previous launch's tail.  pdl == 2: no operand of
    // this launch comes from the launch right before it (dX after the dW0 group), so the whole kernel may overlap that
    // launch's tail; the wait moves to the end, where it keeps completion transitive for whatever follows.
    const bool pdl_late = grp.pdl == 2;
    if (!pdl_late) { pdl_wait(); pdl_launch_dependents(); }

    // work item -> (problem, batch, m tile, n tile, k-block range)
    auto decode = [&](int w, int& g, int& b, int& mt, int& nt, int& kb0, int& kb1) {
        g = 0;
#pragma unroll
        for (int i = 1; i < MAX_GROUP; ++i)
            if (i < grp.count && w >= sargs[i].work_begin) g = i;
        const TcArgs& a = sargs[g];
        w -= a.work_begin;
        nt = w % a.tiles_n;
        w /= a.tiles_n;
        const int ks = w % a.k_splits;
        w /= a.k_splits;
        mt = w % a.tiles_m;
        b = w / a.tiles_m;
        kb0 = (ks * a.kb_total) / a.k_splits;          // ks < 148, kb_total < 2^20: no overflow
        kb1 = ((ks + 1) * a.kb_total) / a.k_splits;
    };

    if (warp == W_PROD) {
        // ------------------------------ TMA producer ------------------------------
        // The WHOLE warp walks the loop (convergent control flow: addresses and coordinates live in uniform registers);
        // one elected lane issues.  With `if (lane == 0)` around the loop every cp.async.bulk.tensor / tcgen05.mma /
        // tcgen05.commit was wrapped in an ELECT + 5-7 x R2UR + branch sequence (the operands of those instructions must be
        // uniform, and inside a divergent region the compiler has to re-establish that each time): ~70 issue slots per
        // k-block for the producer and ~100 for the MMA warp, against 512 clocks of MMA work per k-block.
        {
            int stage = 0;
            uint32_t phase = 0;
            int mit = 0;                                  // masked tiles so far (mask buffer = mit & 1)
            // CTA pair: both producers credit the leader's full barrier, which expects both CTAs' bytes
            const uint32_t full0 = CTA2 ? mapa_shared(full_bar(0), 0u) : full_bar(0);
            for (int w = w_first; w < total_work; w += w_step) {
                int g, b, mt, nt, kb0, kb1;
                decode(w, g, b, mt, nt, kb0, kb1);
                const CUtensorMap* tmA = &grp.p[g].tmA;
                const CUtensorMap* tmB = &grp.p[g].tmB;
                const int bb = sargs[g].b_batched ? b : 0;
                const int m0 = mt * Cfg::TILE_M + (int)cta_rank * BM;
                const int n0 = nt * BN + (int)cta_rank * (CL2 ? Cfg::B_ROWS : 0);
                if (MASK && (sargs[g].flags & F_MASK)) {
                    const int mb = mit & 1;
                    mbar_wait(mempty_bar(mb), ((uint32_t)(mit >> 1) & 1u) ^ 1u);
                    int nbox = (sargs[g].N - nt * BN + 63) / 64;
                    if (nbox > Cfg::MASK_BOXES) nbox = Cfg::MASK_BOXES;
                    if (elect_one()) {
                        mbar_expect_tx(mfull_bar(mb), (uint32_t)nbox * BOX_BYTES);
                        for (int bx = 0; bx < nbox; ++bx)
                            tma_load_3d(mask_base + (uint32_t)(mb * Cfg::MASK_BOXES + bx) * BOX_BYTES, &grp.p[g].tmM,
                                        mfull_bar(mb), nt * BN + bx * 64, mt * BM, b);
                    }
                    ++mit;
                }
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(empty_bar(stage), phase ^ 1u);
                    const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES;
                    const uint32_t sb = sa + A_STAGE_BYTES;
                    const uint32_t fb = full0 + 8u * stage;
                    if (elect_one()) {
                    if (!CTA2 || cta_rank == 0) mbar_expect_tx(full_bar(stage), Cfg::STAGE_BYTES * (CTA2 ? 2 : 1));
                    auto load = [&](uint32_t dst, const CUtensorMap* tm, int c0, int c1, int c2) {
                        if (CTA2) tma_load_3d_2sm(dst, tm, fb, c0, c1, c2);
                        else tma_load_3d(dst, tm, fb, c0, c1, c2);
                    };
                    // MCB: this CTA's half of the B tile goes to the same place in both CTAs
                    const uint32_t sbh = sb + (MCB ? cta_rank * (uint32_t)(Cfg::B_ROWS * BK * 2) : 0u);
                    auto load_b = [&](uint32_t dst, const CUtensorMap* tm, int c0, int c1, int c2) {
                        if (MCB) tma_load_3d_mc(dst, tm, fb, c0, c1, c2, (uint16_t)3);
                        else load(dst, tm, c0, c1, c2);
                    };
                    if (A_MN) {
#pragma unroll
                        for (int i = 0; i < BM / 64; ++i) load(sa + i * 8192, tmA, m0 + i * 64, kb * BK, b);
                    } else {
                        load(sa, tmA, kb * BK, m0, b);
                    }
                    if (B_MN) {
#pragma unroll
                        for (int i = 0; i < Cfg::B_ROWS / 64; ++i) load_b(sbh + i * 8192, tmB, n0 + i * 64, kb * BK, bb);
                    } else {
                        load_b(sbh, tmB, kb * BK, n0, bb);
                    }
                    }   // elected lane
                    if (++stage == STAGES) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == W_MMA) {
        // ------------------------------ MMA issuer ------------------------------
        if (cta_rank == 0 || MCB) {            // whole warp, one elected lane issues (see the producer)
            constexpr uint32_t idesc = make_idesc_bf16(CTA2 ? 2 * BM : BM, BN, A_MN, B_MN);
            constexpr uint32_t A_LBO = A_MN ? 8192u : 16u;
            constexpr uint32_t B_LBO = B_MN ? 8192u : 16u;
            constexpr uint32_t A_KSTEP = A_MN ? 2048u : 32u;   // bytes per UMMA_K = 16 elements
            constexpr uint32_t B_KSTEP = B_MN ? 2048u : 32u;
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int w = w_first; w < total_work; w += w_step, ++it) {
                int g, b, mt, nt, kb0, kb1;
                decode(w, g, b, mt, nt, kb0, kb1);
                const int as = it & 1;
                const uint32_t aphase = (uint32_t)(it >> 1) & 1u;
                mbar_wait(tempty_bar(as), aphase ^ 1u);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(as * BN);
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(full_bar(stage), phase);
                    tc_fence_after();
                    const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES;
                    const uint32_t sb = sa + A_STAGE_BYTES;
                    if (elect_one()) {
#pragma unroll
                        for (int k = 0; k < BK / 16; ++k) {
                            const uint64_t ad = make_smem_desc(sa + k * A_KSTEP, A_LBO, 1024u);
                            const uint64_t bd = make_smem_desc(sb + k * B_KSTEP, B_LBO, 1024u);
                            if (CTA2) tc_mma_bf16_2sm(d_tmem, ad, bd, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
                            else tc_mma_bf16(d_tmem, ad, bd, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
                        }
                        if (CTA2) tc_commit_2sm(empty_bar(stage), 3);     // frees the stage in both CTAs
                        else if (MCB) tc_commit_mc(empty_bar(stage), 3);  // one of the two releases each CTA's producer waits for
                        else tc_commit(empty_bar(stage));
                    }
                    if (++stage == STAGES) { stage = 0; phase ^= 1u; }
                }
                if (elect_one()) {
                    if (CTA2) tc_commit_2sm(tfull_bar(as), 3);
                    else tc_commit(tfull_bar(as));
                }
            }
        }
    } else {
        // ------------------------------ epilogue ------------------------------
        const int ew = warp;               // 0..7
        const int q = warp & 3;            // TMEM lane quarter this warp may touch
        const int eg = ew >> 2;            // epilogue group: alternate output boxes / 32-column chunks
        const uint32_t my_staging = staging + (uint32_t)(ew * SBUF) * WBOX_BYTES;      // 32 rows x 128 B, private
        const int row = q * 32 + lane;
        int nstore = 0;                    // boxes this group has stored so far
        int it = 0, mit = 0;
        const uint32_t tempty0 = CTA2 ? mapa_shared(tempty_bar(0), 0u) : tempty_bar(0);   // the MMA issuer's barrier
        for (int w = w_first; w < total_work; w += w_step, ++it) {
            int g, b, mt, nt, kb0, kb1;
            decode(w, g, b, mt, nt, kb0, kb1);
            const TcArgs& args = sargs[g];
            const uint32_t flags = args.flags;
            const int as = it & 1;
            const uint32_t aphase = (uint32_t)(it >> 1) & 1u;
            const int m_base = mt * Cfg::TILE_M + (int)cta_rank * BM;
            const int m = m_base + row;
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BN);
            const bool masked = MASK && (flags & F_MASK);
            const int mb = mit & 1;
            mbar_wait(tfull_bar(as), aphase);
            tc_fence_after();
            if (args.out_kind == OUT_ATOMIC_F32) {
#pragma unroll 1
                for (int c = eg; c < BN / 32; c += 2) {
                    const int n0 = nt * BN + c * 32;
                    if (n0 >= args.N) break;
                    uint32_t r[32];
                    float v[32], bv[32];
                    tmem_ld32(taddr + c * 32, r);
                    epilogue_bias32(args, flags, bv, n0);
                    tmem_ld_wait();
                    if (m < args.M) {
                        epilogue_math32(args, flags, r, bv, v, b, m, n0);
                        epilogue_atomic32(args, v, b, m, n0);
                    }
                }
            } else {
                const bool bf = args.out_kind == OUT_BF16;
                const int box_cols = bf ? 64 : 32;                     // 128 bytes of output per box row
                int nbox = (args.N - nt * BN + box_cols - 1) / box_cols;
                if (nbox > BN / box_cols) nbox = BN / box_cols;
                if (masked) mbar_wait(mfull_bar(mb), (uint32_t)(mit >> 1) & 1u);
#pragma unroll 1
                for (int bx = eg; bx < nbox; bx += 2) {
                    const uint32_t sbuf = my_staging + (uint32_t)(nstore % SBUF) * WBOX_BYTES;
                    ++nstore;
                    if (lane == 0) bulk_wait_read<SBUF - 1>();         // the box we are about to overwrite has been read
                    __syncwarp();
                    const int cpb = bf ? 2 : 1;
#pragma unroll 1
                    for (int c = 0; c < cpb; ++c) {
                        const int n0 = nt * BN + bx * box_cols + c * 32;
                        if (n0 >= args.N) break;
                        uint32_t r[32];
                        float v[32], bv[32];
                        tmem_ld32(taddr + bx * box_cols + c * 32, r);
                        epilogue_bias32(args, flags, bv, n0);
                        tmem_ld_wait();
                        epilogue_math32(args, flags, r, bv, v, b, m, n0);
                        if (bf) {
                            if (masked) {
                                const float ms = args.mask_scale;
#pragma unroll
                                for (int i = 0; i < 32; ++i) v[i] *= ms;
                            }
                            uint32_t pk[16];
#pragma unroll
                            for (int i = 0; i < 16; ++i) {
                                __nv_bfloat162 h2 = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
                                pk[i] = *reinterpret_cast<uint32_t*>(&h2);
                            }
                            if (MASK && masked) {                          // keep where the saved activation is > 0
                                const uint32_t mbox = mask_base + (uint32_t)(mb * Cfg::MASK_BOXES + bx) * BOX_BYTES;
                                const __nv_bfloat162 zero2 = __floats2bfloat162_rn(0.f, 0.f);
#pragma unroll
                                for (int j = 0; j < 4; ++j) {
                                    uint32_t mw[4];
                                    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                                                 : "=r"(mw[0]), "=r"(mw[1]), "=r"(mw[2]), "=r"(mw[3])
                                                 : "r"(box_addr(mbox, row, c * 4 + j)) : "memory");
#pragma unroll
                                    for (int t = 0; t < 4; ++t)
                                        pk[4 * j + t] &= __hgt2_mask(*reinterpret_cast<__nv_bfloat162*>(&mw[t]), zero2);
                                }
                            }
                            if (flags & F_COLSUM) {
                                const bool live = m < args.M;
#pragma unroll
                                for (int i = 0; i < 16; ++i) {
                                    v[2 * i] = live ? __uint_as_float(pk[i] << 16) : 0.f;
                                    v[2 * i + 1] = live ? __uint_as_float(pk[i] & 0xFFFF0000u) : 0.f;
                                }
                                const float tot = warp_colsum32(v, lane);
                                if (n0 + lane < args.N) atomicAdd(args.colsum + n0 + lane, tot);
                            }
#pragma unroll
                            for (int j = 0; j < 4; ++j)
                                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(box_addr(sbuf, lane, c * 4 + j)),
                                             "r"(pk[4 * j]), "r"(pk[4 * j + 1]), "r"(pk[4 * j + 2]), "r"(pk[4 * j + 3]) : "memory");
                        } else {
#pragma unroll
                            for (int j = 0; j < 8; ++j)
                                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(box_addr(sbuf, lane, j)),
                                             "r"(__float_as_uint(v[4 * j])), "r"(__float_as_uint(v[4 * j + 1])),
                                             "r"(__float_as_uint(v[4 * j + 2])), "r"(__float_as_uint(v[4 * j + 3])) : "memory");
                        }
                    }
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) {
                        tma_store_3d(&grp.p[g].tmO, sbuf, nt * BN + bx * box_cols, m_base + q * 32, b);
                        bulk_commit();
                    }
                }
            }
            // every tcgen05.ld / mask read of this tile by this warp has completed: release both buffers
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (CTA2) mbar_arrive_cluster(tempty0 + 8u * as);
                else mbar_arrive(tempty_bar(as));
                if (masked) mbar_arrive(mempty_bar(mb));
            }
            if (masked) ++mit;
        }
        if (lane == 0) bulk_wait_read<0>();   // staging must outlive the last tile stores
    }

    if (pdl_late) { pdl_wait(); pdl_launch_dependents(); }
    tc_fence_before();
    if (CL2) cluster_sync_all();           // the leader's MMAs read the peer's shared memory and write its TMEM
    else __syncthreads();
    if (warp == W_MMA) {
        tc_fence_after();
        if (CTA2) tmem_dealloc_2sm(tmem_base, Cfg::TMEM_COLS);
        else tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
    }
}

// ---------------------------------------------------------------------------
// fp32 FMA kernel (parity mode).  64x64 tile, 256 threads, 4x4 per thread, BK = 16.
// ---------------------------------------------------------------------------
struct SimtArgs {
    const float* A; long long a_sm, a_sk, a_sb;
    const float* B; long long b_sn, b_sk, b_sb;
    int M, N, K, k_splits;
    float* out; int out_kind; long long ldo, out_bs;
    const float* bias; int relu;
    const float* mask; long long ld_mask, mask_bs; float mask_scale, alpha;
    uint32_t drop_thr; float drop_scale; unsigned long long drop_seed; const unsigned long long* drop_seed_dev;
};

__global__ void __launch_bounds__(256) gemm_simt_kernel(const SimtArgs a) {
    __shared__ float As[16][64 + 4];
    __shared__ float Bs[16][64 + 4];
    const int b = blockIdx.z / a.k_splits;
    const int ks = blockIdx.z % a.k_splits;
    const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
    const int kchunks = (a.K + 15) / 16;
    const int kc0 = (int)((long long)ks * kchunks / a.k_splits);
    const int kc1 = (int)((long long)(ks + 1) * kchunks / a.k_splits);
    const float* Ab = a.A + (long long)b * a.a_sb;
    const float* Bb = a.B + (long long)b * a.b_sb;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    float acc[4][4] = {};
    for (int kc = kc0; kc < kc1; ++kc) {
        const int k0 = kc * 16;
        for (int e = threadIdx.x; e < 16 * 64; e += 256) {
            int kk, mm;
            if (a.a_sk == 1) { kk = e & 15; mm = e >> 4; } else { mm = e & 63; kk = e >> 6; }
            float va = 0.f;
            if (m0 + mm < a.M && k0 + kk < a.K) va = Ab[(long long)(m0 + mm) * a.a_sm + (long long)(k0 + kk) * a.a_sk];
            As[kk][mm] = va;
            int nn;
            if (a.b_sk == 1) { kk = e & 15; nn = e >> 4; } else { nn = e & 63; kk = e >> 6; }
            float vb = 0.f;
            if (n0 + nn < a.N && k0 + kk < a.K) vb = Bb[(long long)(n0 + nn) * a.b_sn + (long long)(k0 + kk) * a.b_sk];
            Bs[kk][nn] = vb;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) {
            float av[4], bv[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) av[i] = As[kk][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) bv[j] = Bs[kk][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        if (m >= a.M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n >= a.N) continue;
            float v = acc[i][j] * a.alpha;
            if (a.bias) v += a.bias[n];
            if (a.relu) v = fmaxf(v, 0.f);
            if (a.drop_thr != 0u)
                v = dropout_keep(a.drop_seed + (a.drop_seed_dev != nullptr ? __ldg(a.drop_seed_dev) : 0ull),
                                 ((unsigned long long)b * a.M + m) * (unsigned long long)a.N + n, a.drop_thr)
                        ? v * a.drop_scale : 0.f;
            if (a.mask) {
                float mk = a.mask[(long long)b * a.mask_bs + (long long)m * a.ld_mask + n];
                v = mk > 0.f ? v * a.mask_scale : 0.f;
            }
            float* op = a.out + (long long)b * a.out_bs + (long long)m * a.ldo + n;
            if (a.out_kind == OUT_ATOMIC_F32) atomicAdd(op, v); else *op = v;
        }
    }
}

// column sums of an fp32 [batch][M][N] output (parity mode's counterpart of the fused column sums)
__global__ void __launch_bounds__(256) colsum_f32_kernel(const float* __restrict__ X, long long ld, long long bs,
                                                         int batch, int M, int N, float* __restrict__ out) {
    __shared__ float red[8][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int col = blockIdx.x * 32 + tx;
    const long long rows = (long long)batch * M;
    const long long r0 = (long long)blockIdx.y * 256;
    float acc = 0.f;
    if (col < N)
        for (long long r = r0 + ty; r < r0 + 256 && r < rows; r += 8) {
            const long long b = r / M, m = r - b * M;
            acc += X[b * bs + m * ld + col];
        }
    red[ty][tx] = acc;
    __syncthreads();
    if (ty == 0 && col < N) {
        float tot = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) tot += red[i][tx];
        atomicAdd(out + col, tot);
    }
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    return fn;
}

// 3-D tensor map over a row-major [batch][rows][inner] array of 2-byte (bf16) or 4-byte (fp32)
// elements, 128B swizzle; ld / bs in elements.
int make_map(CUtensorMap* map, const void* ptr, int esize, int64_t inner, int64_t rows, int64_t batch, int64_t ld,
             int64_t bs, int box_inner, int box_rows, const char* what) {
    EncodeTiledFn enc = get_encode();
    VPR_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled not available from the driver");
    VPR_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, "%s: pointer must be 16-byte aligned", what);
    VPR_REQUIRE((ld * esize) % 16 == 0, "%s: row stride must be a multiple of 16 bytes (got %lld elements of %d bytes)",
                what, (long long)ld, esize);
    VPR_REQUIRE(inner > 0 && rows > 0 && batch > 0, "%s: empty tensor", what);
    if (batch == 1 || bs == 0) { batch = 1; bs = rows * ld; }
    VPR_REQUIRE((bs * esize) % 16 == 0, "%s: batch stride must be a multiple of 16 bytes (got %lld elements)", what,
                (long long)bs);
    // Encoding a map is a ~2 us driver call and a training step needs ~50 of them over the same buffers
    // (the caller's allocator hands back the same pointers step after step): keep them in a small cache.
    struct Key { const void* ptr; int64_t inner, rows, batch, ld, bs; int esize, box_inner, box_rows; };
    struct Entry { Key key; CUtensorMap map; bool valid; };
    static thread_local Entry cache[512];
    Key key;
    memset(&key, 0, sizeof(key));
    key.ptr = ptr; key.inner = inner; key.rows = rows; key.batch = batch; key.ld = ld; key.bs = bs;
    key.esize = esize; key.box_inner = box_inner; key.box_rows = box_rows;
    uint64_t hsh = 1469598103934665603ull;
    for (size_t i = 0; i < sizeof(Key); ++i) hsh = (hsh ^ reinterpret_cast<const unsigned char*>(&key)[i]) * 1099511628211ull;
    Entry& ce = cache[(hsh >> 20) & 511];
    if (ce.valid && memcmp(&ce.key, &key, sizeof(Key)) == 0) { *map = ce.map; return 0; }
    cuuint64_t dims[3] = {(cuuint64_t)inner, (cuuint64_t)rows, (cuuint64_t)batch};
    cuuint64_t strides[2] = {(cuuint64_t)ld * esize, (cuuint64_t)bs * esize};
    cuuint32_t box[3] = {(cuuint32_t)box_inner, (cuuint32_t)box_rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(map, esize == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3,
                     const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    VPR_REQUIRE(r == CUDA_SUCCESS, "%s: cuTensorMapEncodeTiled failed (%d): inner=%lld rows=%lld batch=%lld ld=%lld bs=%lld",
                what, (int)r, (long long)inner, (long long)rows, (long long)batch, (long long)ld, (long long)bs);
    ce.key = key; ce.map = *map; ce.valid = true;
    return 0;
}

template <bool A_MN, bool B_MN, int BN, bool MASK, bool CTA2, bool MCB = false>
static int launch_tc(const TcGroup& grp, int grid, cudaStream_t st) {
    using Cfg = TcCfg<BN, MASK, CTA2, MCB>;
    static bool attr_set = false;
    if (!attr_set) {
        VPR_SMEM_OK((gemm_tc_kernel<A_MN, B_MN, BN, MASK, CTA2, MCB>), Cfg::SMEM_BYTES);
        attr_set = true;
    }
    VPR_CUDA_OK(launch_ex(gemm_tc_kernel<A_MN, B_MN, BN, MASK, CTA2, MCB>, dim3((unsigned)grid), dim3(GEMM_THREADS),
                          (size_t)Cfg::SMEM_BYTES, st, grp.pdl != 0, (CTA2 || MCB) ? 2 : 1, grp));
    return launch_ok("gemm_tc_kernel");
}

template <int BN>
static int dispatch_major(int a_mn, int b_mn, bool mask, bool cta2, bool mcb, const TcGroup& grp, int grid, cudaStream_t st) {
    if constexpr (BN <= 128) {
        if (mask) {
            VPR_REQUIRE(!a_mn, "the relu-mask epilogue needs a K-major A operand");
            return b_mn ? launch_tc<false, true, BN, true, false>(grp, grid, st)
                        : launch_tc<false, false, BN, true, false>(grp, grid, st);
        }
    }
    if constexpr (BN == 256) {
        if (mcb) return launch_tc<false, false, BN, false, false, true>(grp, grid, st);      // planned only for K-major operands
        if (cta2) {
            if (!a_mn && !b_mn) return launch_tc<false, false, BN, false, true>(grp, grid, st);
            if (!a_mn && b_mn) return launch_tc<false, true, BN, false, true>(grp, grid, st);
            if (a_mn && !b_mn) return launch_tc<true, false, BN, false, true>(grp, grid, st);
            return launch_tc<true, true, BN, false, true>(grp, grid, st);
        }
    }
    if (!a_mn && !b_mn) return launch_tc<false, false, BN, false, false>(grp, grid, st);
    if (!a_mn && b_mn) return launch_tc<false, true, BN, false, false>(grp, grid, st);
    if (a_mn && !b_mn) return launch_tc<true, false, BN, false, false>(grp, grid, st);
    return launch_tc<true, true, BN, false, false>(grp, grid, st);
}

// VPR_GEMM_2CTA=0 forces single-CTA tiles (A/B measurements)
static bool cta_pairs_enabled() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("VPR_GEMM_2CTA");
        v = (e != nullptr && e[0] == '0') ? 0 : (e != nullptr && e[0] == '2') ? 2 : 1;
    }
    return v >= 1;
}
static bool cta_pairs_forced() { cta_pairs_enabled(); const char* e = getenv("VPR_GEMM_2CTA"); return e != nullptr && e[0] == '2'; }

static int auto_bn(int N) { return (N > 128) ? 256 : (N > 64) ? 128 : 64; }

// Tile shape for a group of bf16 problems: the widest N any problem wants; 256 x 256 CTA-pair tiles
// when that does not inflate the MMA work (short-M problems would leave the peer CTA's rows empty).
static bool mcb_enabled() {
    static int v = -1;
    // measured (B200, H = dropout(relu(X W0cat^T + b)) in the graphed step).  Round 1, election loops around every TMA /
    // MMA instruction: 97.6 us with the multicast clusters vs 96.0 us with independent CTAs, so it was opt-in.  With the
    // convergent producer / MMA warps the order flips: 0.3769 vs 0.3783 ms/step (two alternating pairs of runs, same box),
    // so it is on unless VPR_GEMM_MCB=0.
    if (v < 0) { const char* e = getenv("VPR_GEMM_MCB"); v = (e != nullptr && e[0] == '0') ? 0 : 1; }
    return v == 1;
}

static int plan_group(const GemmDesc* gs, int count, int& bn, bool& any_mask, bool& cta2, bool& mcb) {
    VPR_REQUIRE(count >= 1 && count <= MAX_GROUP, "a GEMM group holds 1..%d problems", MAX_GROUP);
    bn = 0;
    any_mask = false;
    for (int i = 0; i < count; ++i) {
        VPR_REQUIRE(gs[i].a_mn == gs[0].a_mn && gs[i].b_mn == gs[0].b_mn, "grouped GEMMs must share the operand layout");
        const int want = gs[i].bn ? gs[i].bn : auto_bn(gs[i].N);
        if (want > bn) bn = want;
        any_mask = any_mask || gs[i].mask != nullptr;
    }
    if (any_mask && bn > 128) bn = 128;          // the mask tile shares shared memory with the operand ring
    VPR_REQUIRE(bn == 64 || bn == 128 || bn == 256, "bad tile N %d", bn);
    cta2 = false;
    bool single = false;
    // measured (tools/gemm_bench.py, B200): with the dropout epilogue the leader's MMA ends up waiting for the
    // slower of two CTAs' epilogues (100 us vs 90 us for H); the gradient GEMMs gain 8-12 % from pair tiles
    for (int i = 0; i < count; ++i) single = single || gs[i].cta_pair == 2 || (gs[i].dropout_p > 0.f && !cta_pairs_forced());
    if (bn == 256 && !any_mask && !single && cta_pairs_enabled()) {
        double c1 = 0, c2 = 0;
        for (int i = 0; i < count; ++i) {
            const double w = (double)gs[i].batch * ((gs[i].N + bn - 1) / bn) * gs[i].K;
            c1 += w * ((gs[i].M + BM - 1) / BM) * BM;
            c2 += w * ((gs[i].M + 2 * BM - 1) / (2 * BM)) * (2 * BM);
        }
        cta2 = c2 <= 1.05 * c1;
    }
    // dropout epilogue (kept off pair tiles, see above): two independent CTAs share each B tile by multicast instead
    mcb = false;
    if (bn == 256 && !any_mask && !cta2 && single && mcb_enabled() && !gs[0].a_mn && !gs[0].b_mn) {
        double c1 = 0, c2 = 0;
        bool forced_single = false;
        for (int i = 0; i < count; ++i) {
            const double w = (double)gs[i].batch * ((gs[i].N + bn - 1) / bn) * gs[i].K;
            c1 += w * ((gs[i].M + BM - 1) / BM) * BM;
            c2 += w * ((gs[i].M + 2 * BM - 1) / (2 * BM)) * (2 * BM);
            forced_single = forced_single || gs[i].cta_pair == 2;
        }
        mcb = !forced_single && c2 <= 1.05 * c1;
    }
    return 0;
}

GemmTilePlan gemm_tile_plan(const GemmDesc* gs, int count) {
    GemmTilePlan p = {BM, 256, num_sms()};
    if (gs[0].dtype != DT_BF16) return p;
    int bn; bool any_mask, cta2, mcb;
    if (plan_group(gs, count, bn, any_mask, cta2, mcb) != 0) return p;
    p.tile_m = (cta2 || mcb) ? 2 * BM : BM;
    p.tile_n = bn;
    p.slots = (cta2 || mcb) ? num_sms() / 2 : num_sms();
    return p;
}

static int gemm_tc_group(const GemmDesc* gs, int count, cudaStream_t st) {
    int bn;
    bool any_mask, cta2, mcb;
    VPR_TRY(plan_group(gs, count, bn, any_mask, cta2, mcb));
    const bool cl2 = cta2 || mcb;
    const int tile_m = cl2 ? 2 * BM : BM;
    TcGroup grp;
    grp.count = count;
    long long total = 0;
    for (int i = 0; i < count; ++i) {
        const GemmDesc& g = gs[i];
        TcProblem& p = grp.p[i];
        if (g.a_mn) VPR_TRY(make_map(&p.tmA, g.A, 2, g.M, g.K, g.batch, g.lda, g.a_bs, 64, 64, "GEMM operand A"));
        else        VPR_TRY(make_map(&p.tmA, g.A, 2, g.K, g.M, g.batch, g.lda, g.a_bs, 64, BM, "GEMM operand A"));
        if (g.b_mn) VPR_TRY(make_map(&p.tmB, g.B, 2, g.N, g.K, g.batch, g.ldb, g.b_bs, 64, 64, "GEMM operand B"));
        else        VPR_TRY(make_map(&p.tmB, g.B, 2, g.K, g.N, g.batch, g.ldb, g.b_bs, 64, cl2 ? bn / 2 : bn, "GEMM operand B"));
        if (g.out_kind == OUT_BF16)
            VPR_TRY(make_map(&p.tmO, g.out, 2, g.N, g.M, g.batch, g.ldo, g.out_bs, 64, 32, "GEMM output"));
        else if (g.out_kind == OUT_F32)
            VPR_TRY(make_map(&p.tmO, g.out, 4, g.N, g.M, g.batch, g.ldo, g.out_bs, 32, 32, "GEMM output"));
        if (g.mask != nullptr)
            VPR_TRY(make_map(&p.tmM, g.mask, 2, g.N, g.M, g.batch, g.ld_mask, g.mask_bs, 64, BM, "GEMM relu-mask"));
        TcArgs& a = p.args;
        a.M = g.M; a.N = g.N; a.K = g.K; a.batch = g.batch;
        a.tiles_m = (g.M + tile_m - 1) / tile_m;
        a.tiles_n = (g.N + bn - 1) / bn;
        a.kb_total = (g.K + BK - 1) / BK;
        a.k_splits = g.k_splits < 1 ? 1 : g.k_splits;
        if (a.k_splits > a.kb_total) a.k_splits = a.kb_total;
        a.b_batched = (g.batch > 1 && g.b_bs != 0) ? 1 : 0;
        a.out = g.out; a.out_kind = g.out_kind; a.ldo = g.ldo; a.out_bs = g.out_bs;
        a.bias = g.bias; a.colsum = g.colsum;
        a.mask_scale = g.mask_scale; a.alpha = g.alpha;
        a.drop_thr = g.dropout_p > 0.f ? dropout_threshold(g.dropout_p) : 0u;
        a.drop_scale = g.dropout_p > 0.f ? 1.f / (1.f - g.dropout_p) : 1.f; a.drop_seed = g.dropout_seed;
        a.drop_seed_dev = reinterpret_cast<const unsigned long long*>(g.dropout_seed_dev);
        a.flags = (g.bias ? F_BIAS : 0u) | (g.relu ? F_RELU : 0u) | (a.drop_thr ? F_DROP : 0u) | (g.mask ? F_MASK : 0u) |
                  (g.alpha != 1.f ? F_ALPHA : 0u) | (g.colsum ? F_COLSUM : 0u);
        VPR_REQUIRE(g.colsum == nullptr || g.out_kind == OUT_BF16, "fused column sums need a bf16 output");
        VPR_REQUIRE(g.mask == nullptr || g.out_kind == OUT_BF16, "the relu-mask epilogue needs a bf16 output");
        a.work_begin = (int)total;
        total += (long long)a.batch * a.tiles_m * a.tiles_n * a.k_splits;
        VPR_REQUIRE(total < (1ll << 31), "too many tiles");
    }
    grp.total_work = (int)total;
    grp.pdl = 0;
    for (int i = 0; i < count; ++i) if (gs[i].pdl > grp.pdl) grp.pdl = gs[i].pdl;
    for (int i = 0; i < count; ++i) if (gs[i].pdl != 2 && grp.pdl == 2) grp.pdl = 1;      // late wait only if every problem allows it
    int reserve = 0;
    for (int i = 0; i < count; ++i) if (gs[i].reserve_sms > reserve) reserve = gs[i].reserve_sms;
    if (reserve > num_sms() / 2) reserve = num_sms() / 2;
    int grid;
    if (cl2) {
        const int pairs = (num_sms() - reserve) / 2;
        grid = 2 * (int)(total < pairs ? total : pairs);
    } else {
        grid = (int)(total < num_sms() - reserve ? total : num_sms() - reserve);
    }
    if (bn == 256) return dispatch_major<256>(gs[0].a_mn, gs[0].b_mn, any_mask, cta2, mcb, grp, grid, st);
    if (bn == 128) return dispatch_major<128>(gs[0].a_mn, gs[0].b_mn, any_mask, cta2, mcb, grp, grid, st);
    return dispatch_major<64>(gs[0].a_mn, gs[0].b_mn, any_mask, cta2, mcb, grp, grid, st);
}

static int gemm_simt(const GemmDesc& g, cudaStream_t st) {
    SimtArgs a;
    a.A = reinterpret_cast<const float*>(g.A);
    a.a_sm = g.a_mn ? 1 : g.lda; a.a_sk = g.a_mn ? g.lda : 1; a.a_sb = g.a_bs;
    a.B = reinterpret_cast<const float*>(g.B);
    a.b_sn = g.b_mn ? 1 : g.ldb; a.b_sk = g.b_mn ? g.ldb : 1; a.b_sb = g.b_bs;
    a.M = g.M; a.N = g.N; a.K = g.K;
    int kchunks = (g.K + 15) / 16;
    a.k_splits = g.k_splits < 1 ? 1 : (g.k_splits > kchunks ? kchunks : g.k_splits);
    VPR_REQUIRE(g.out_kind != OUT_BF16, "fp32 GEMM cannot write bf16");
    a.out = reinterpret_cast<float*>(g.out); a.out_kind = g.out_kind; a.ldo = g.ldo; a.out_bs = g.out_bs;
    a.bias = g.bias; a.relu = g.relu;
    a.mask = reinterpret_cast<const float*>(g.mask); a.ld_mask = g.ld_mask; a.mask_bs = g.mask_bs;
    a.mask_scale = g.mask_scale; a.alpha = g.alpha;
    a.drop_thr = g.dropout_p > 0.f ? dropout_threshold(g.dropout_p) : 0u;
    a.drop_scale = g.dropout_p > 0.f ? 1.f / (1.f - g.dropout_p) : 1.f; a.drop_seed = g.dropout_seed;
    a.drop_seed_dev = reinterpret_cast<const unsigned long long*>(g.dropout_seed_dev);
    dim3 grid((g.N + 63) / 64, (g.M + 63) / 64, g.batch * a.k_splits);
    VPR_REQUIRE(grid.y < 65536 && grid.z < 65536, "fp32 GEMM grid too large");
    gemm_simt_kernel<<<grid, 256, 0, st>>>(a);
    VPR_TRY(launch_ok("gemm_simt_kernel"));
    if (g.colsum != nullptr) {
        VPR_REQUIRE(g.out_kind == OUT_F32, "column sums need a stored output");
        const long long rows = (long long)g.batch * g.M;
        dim3 cg((g.N + 31) / 32, (unsigned)((rows + 255) / 256));
        colsum_f32_kernel<<<cg, 256, 0, st>>>(a.out, g.ldo, g.out_bs, g.batch, g.M, g.N, g.colsum);
        VPR_TRY(launch_ok("colsum_f32_kernel"));
    }
    return 0;
}

int gemm_run_group(const GemmDesc* gs, int count, cudaStream_t st) {
    for (int i = 0; i < count; ++i) {
        const GemmDesc& g = gs[i];
        VPR_REQUIRE(g.M > 0 && g.N > 0 && g.K > 0 && g.batch > 0, "empty GEMM %d x %d x %d (batch %d)", g.M, g.N, g.K, g.batch);
        VPR_REQUIRE(g.k_splits <= 1 || g.out_kind == OUT_ATOMIC_F32, "split-K needs the atomic epilogue");
        VPR_REQUIRE(g.dtype == gs[0].dtype, "grouped GEMMs must share the dtype");
    }
    if (gs[0].dtype == DT_BF16) return gemm_tc_group(gs, count, st);
    for (int i = 0; i < count; ++i) VPR_TRY(gemm_simt(gs[i], st));
    return 0;
}

int gemm_run(const GemmDesc& g, cudaStream_t st) { return gemm_run_group(&g, 1, st); }

}  // namespace vpr
