// Persistent, warp-specialised, grouped bf16 GEMM for sm_100a.
//   warp 0      : TMA producer (cp.async.bulk.tensor, SWIZZLE_128B boxes into an mbarrier ring)
//   warp 1      : TMEM allocator + single-thread tcgen05.mma issuer (fp32 accumulators in TMEM,
//                 double-buffered so the epilogue of tile i overlaps the MMAs of tile i+1)
//   warps 2..9  : epilogue: tcgen05.ld -> bias / relu / dropout -> swizzled smem staging tile ->
//                 coalesced 16-byte global stores (relu-mask applied and column sums taken on the
//                 way out), or vector fp32 reductions for split-K
// Operands may be K-major or MN-major in global memory (all four combinations), batched through
// the third tensor-map dimension and split along K.  One launch runs up to MAX_GROUP independent
// problems of the same operand layout / tile shape (the token-MLP GEMMs ride along with the
// per-patch GEMMs instead of paying their own launch + pipeline fill).
#include "common.cuh"
#include "gemm.h"

#include <mutex>

namespace vpr {

// ---------------------------------------------------------------------------
// tcgen05 kernel
// ---------------------------------------------------------------------------
static constexpr int BM = 128;
static constexpr int BK = 64;                   // 64 bf16 = 128 B = one swizzle span
static constexpr int A_STAGE_BYTES = BM * BK * 2;
static constexpr int GEMM_THREADS = 320;        // TMA warp, MMA warp, 8 epilogue warps
static constexpr int EPI_WARPS = 8;
static constexpr int SMEM_BUDGET = 196608;      // operand ring
static constexpr int STAGING_BYTES = 128 * 256; // epilogue staging: 128 rows x 256 B
static constexpr int MAX_GROUP = 4;

struct TcArgs {
    int M, N, K, batch;
    int tiles_m, tiles_n, k_splits, kb_total;
    int b_batched;
    int work_begin;            // first work item of this problem inside the group
    void* out;
    int out_kind;
    long long ldo, out_bs;
    const float* bias;
    int relu;
    const bf16* mask;
    long long ld_mask, mask_bs;
    float mask_scale, alpha;
    uint32_t drop_thr; float drop_scale; unsigned long long drop_seed;
    float* colsum;             // [N] fp32 += column sums of the stored (masked, rounded) output
};

struct TcProblem {
    CUtensorMap tmA, tmB;
    TcArgs args;
};
struct TcGroup {
    TcProblem p[MAX_GROUP];
    int count;
    int total_work;
};
static_assert(sizeof(TcArgs) % 8 == 0, "TcArgs is copied to shared memory in 4-byte words");
static constexpr int ARGS_SMEM_BYTES = ((MAX_GROUP * (int)sizeof(TcArgs) + 127) / 128) * 128;

template <int BN>
struct TcCfg {
    static constexpr int B_STAGE_BYTES = BN * BK * 2;
    static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
    static constexpr int STAGES = SMEM_BUDGET / STAGE_BYTES;
    static constexpr int TMEM_COLS = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128
                                     : (2 * BN <= 256) ? 256 : 512;
    // ring + staging + barriers/tmem slot (256) + column-sum scratch (512) + alignment slack
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + STAGING_BYTES + 256 + 512 + ARGS_SMEM_BYTES + 1024;
};

// Per-thread epilogue math on one 32-column chunk of one output row.
__device__ __forceinline__ void epilogue_math32(const TcArgs& a, const uint32_t (&r)[32], float (&v)[32], int b,
                                                int m, int n0) {
    const int N = a.N;
    const bool full = (n0 + 32 <= N);
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]) * a.alpha;
    if (a.bias != nullptr) {
        if (full) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float4 bq = __ldg(reinterpret_cast<const float4*>(a.bias + n0) + j);
                v[4 * j] += bq.x; v[4 * j + 1] += bq.y; v[4 * j + 2] += bq.z; v[4 * j + 3] += bq.w;
            }
        } else {
#pragma unroll
            for (int i = 0; i < 32; ++i)
                if (n0 + i < N) v[i] += __ldg(a.bias + n0 + i);
        }
    }
    if (a.relu) {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
    }
    if (a.drop_thr != 0u) {
        const unsigned long long base = ((unsigned long long)b * a.M + m) * (unsigned long long)N + n0;
        if ((base & 3ull) == 0ull) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const unsigned long long h = dropout_hash64(a.drop_seed, (base >> 2) + j);
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    v[4 * j + i] = ((uint32_t)(h >> (16 * i)) & 0xFFFFu) >= a.drop_thr ? v[4 * j + i] * a.drop_scale : 0.f;
            }
        } else {
#pragma unroll
            for (int i = 0; i < 32; ++i)
                v[i] = dropout_keep(a.drop_seed, base + i, a.drop_thr) ? v[i] * a.drop_scale : 0.f;
        }
    }
    if (a.mask != nullptr) {       // the mask's scale; its zeroing happens in the coalesced flush
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] *= a.mask_scale;
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

// Staging tile: 128 rows x 256 B, 16-byte slots XOR-swizzled with (row & 7) inside each 128 B line so
// that both the row-per-thread writes and the row-per-half-warp reads are bank-conflict free.
__device__ __forceinline__ uint32_t stage_addr(uint32_t base, int row, int i16) {
    return base + (uint32_t)row * 256u + (uint32_t)(i16 >> 3) * 128u + (uint32_t)(((i16 & 7) ^ (row & 7)) << 4);
}

// Phase B of the epilogue for one staged round: every warp owns 16 rows of the tile, a row segment
// (<= 256 B) is spread over LPR consecutive lanes -> fully coalesced global traffic.
template <int ESIZE, int LPR>
__device__ __forceinline__ void flush_round(const TcArgs& a, uint32_t staging, float* cs, int ew, int lane, int b,
                                            int mt, int col0, int seg_n) {
    constexpr int RPS = 32 / LPR;            // rows per warp step
    constexpr int STEPS = 16 / RPS;
    constexpr int EPL = 16 / ESIZE;          // elements per lane (8 bf16 / 4 fp32)
    const int i16 = lane % LPR;
    const int ncol = i16 * EPL;
    const int rsub = lane / LPR;
    const bool col_ok = ncol < seg_n;
    const bool vec_ok = (ncol + EPL <= seg_n);
    const bool use_mask = (ESIZE == 2) && a.mask != nullptr;
    uint4 mq[STEPS];
    if (use_mask) {                          // all mask loads in flight before anything depends on them
        const bool mvec = vec_ok && ((a.ld_mask & 7) == 0);
#pragma unroll
        for (int s = 0; s < STEPS; ++s) {
            const int mm = mt * BM + ew * 16 + s * RPS + rsub;
            mq[s] = make_uint4(0x3f803f80u, 0x3f803f80u, 0x3f803f80u, 0x3f803f80u);   // bf16 1.0
            if (mm < a.M && col_ok) {
                const bf16* mp = a.mask + (long long)b * a.mask_bs + (long long)mm * a.ld_mask + col0 + ncol;
                if (mvec) {
                    mq[s] = __ldg(reinterpret_cast<const uint4*>(mp));
                } else {
                    bf16* mh = reinterpret_cast<bf16*>(&mq[s]);
                    for (int i = 0; i < 8; ++i)
                        if (ncol + i < seg_n) mh[i] = mp[i];
                }
            }
        }
    }
    float csum[EPL];
#pragma unroll
    for (int i = 0; i < EPL; ++i) csum[i] = 0.f;
#pragma unroll
    for (int s = 0; s < STEPS; ++s) {
        const int rr = ew * 16 + s * RPS + rsub;
        const int mm = mt * BM + rr;
        if (mm >= a.M || !col_ok) continue;
        uint4 d;
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(d.x), "=r"(d.y), "=r"(d.z), "=r"(d.w)
                     : "r"(stage_addr(staging, rr, i16)) : "memory");
        const long long goff = (long long)b * a.out_bs + (long long)mm * a.ldo + col0 + ncol;
        if (ESIZE == 2) {
            bf16* dv = reinterpret_cast<bf16*>(&d);
            if (use_mask) {
                const bf16* mh = reinterpret_cast<const bf16*>(&mq[s]);
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    if (!(__bfloat162float(mh[i]) > 0.f)) dv[i] = __float2bfloat16_rn(0.f);
            }
            if (a.colsum != nullptr) {
#pragma unroll
                for (int i = 0; i < 8; ++i) csum[i] += __bfloat162float(dv[i]);
            }
            bf16* op = reinterpret_cast<bf16*>(a.out) + goff;
            if (vec_ok && ((a.ldo & 7) == 0)) {
                *reinterpret_cast<uint4*>(op) = d;
            } else {
                for (int i = 0; i < 8; ++i)
                    if (ncol + i < seg_n) op[i] = dv[i];
            }
        } else {
            float* op = reinterpret_cast<float*>(a.out) + goff;
            if (vec_ok && ((a.ldo & 3) == 0)) {
                *reinterpret_cast<uint4*>(op) = d;
            } else {
                const float* dv = reinterpret_cast<const float*>(&d);
                for (int i = 0; i < 4; ++i)
                    if (ncol + i < seg_n) op[i] = dv[i];
            }
        }
    }
    if (ESIZE == 2 && a.colsum != nullptr) {
        // lanes that share a column group (same i16) first, then one shared-memory add per column
#pragma unroll
        for (int o = LPR; o < 32; o <<= 1)
#pragma unroll
            for (int i = 0; i < EPL; ++i) csum[i] += __shfl_xor_sync(0xffffffffu, csum[i], o);
        if (rsub == 0 && col_ok) {
#pragma unroll
            for (int i = 0; i < EPL; ++i)
                if (ncol + i < seg_n) atomicAdd(cs + ncol + i, csum[i]);
        }
    }
}

template <bool A_MN, bool B_MN, int BN>
__global__ void __launch_bounds__(GEMM_THREADS, 1) gemm_tc_kernel(const __grid_constant__ TcGroup grp) {
    using Cfg = TcCfg<BN>;
    constexpr int STAGES = Cfg::STAGES;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t staging = smem_base + STAGES * Cfg::STAGE_BYTES;
    const uint32_t bar_base = staging + STAGING_BYTES;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
    auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + s); };
    auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + 2 + s); };
    const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 4);
    uint8_t* smem_al = smem_raw + (smem_base - smem_u32(smem_raw));
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_al + (tmem_slot - smem_base));
    float* cs = reinterpret_cast<float*>(smem_al + (bar_base + 256 - smem_base));      // [128] column-sum scratch
    // per-problem arguments in shared memory: dynamically indexed by every role without touching param space
    TcArgs* sargs = reinterpret_cast<TcArgs*>(smem_al + (bar_base + 256 + 512 - smem_base));

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        for (int g = 0; g < grp.count; ++g) {
            tma_prefetch_desc(&grp.p[g].tmA);
            tma_prefetch_desc(&grp.p[g].tmB);
        }
    }
    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < STAGES; ++s) {
                mbar_init(full_bar(s), 1);
                mbar_init(empty_bar(s), 1);
            }
            for (int s = 0; s < 2; ++s) {
                mbar_init(tfull_bar(s), 1);
                mbar_init(tempty_bar(s), EPI_WARPS);
            }
            mbar_fence_init();
        }
        __syncwarp();
        tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
        tmem_relinquish();
    }
    if (warp >= 2 && threadIdx.x - 64 < 128) cs[threadIdx.x - 64] = 0.f;
    for (int i = threadIdx.x; i < grp.count * (int)(sizeof(TcArgs) / 4); i += GEMM_THREADS) {
        const int gi = i / (int)(sizeof(TcArgs) / 4), wi = i % (int)(sizeof(TcArgs) / 4);
        reinterpret_cast<uint32_t*>(sargs + gi)[wi] = reinterpret_cast<const uint32_t*>(&grp.p[gi].args)[wi];
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    const int total_work = grp.total_work;

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

    if (warp == 0) {
        // ------------------------------ TMA producer ------------------------------
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
                int g, b, mt, nt, kb0, kb1;
                decode(w, g, b, mt, nt, kb0, kb1);
                const CUtensorMap* tmA = &grp.p[g].tmA;
                const CUtensorMap* tmB = &grp.p[g].tmB;
                const int bb = sargs[g].b_batched ? b : 0;
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(empty_bar(stage), phase ^ 1u);
                    const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES;
                    const uint32_t sb = sa + A_STAGE_BYTES;
                    mbar_expect_tx(full_bar(stage), Cfg::STAGE_BYTES);
                    if (A_MN) {
#pragma unroll
                        for (int i = 0; i < BM / 64; ++i)
                            tma_load_3d(sa + i * 8192, tmA, full_bar(stage), mt * BM + i * 64, kb * BK, b);
                    } else {
                        tma_load_3d(sa, tmA, full_bar(stage), kb * BK, mt * BM, b);
                    }
                    if (B_MN) {
#pragma unroll
                        for (int i = 0; i < BN / 64; ++i)
                            tma_load_3d(sb + i * 8192, tmB, full_bar(stage), nt * BN + i * 64, kb * BK, bb);
                    } else {
                        tma_load_3d(sb, tmB, full_bar(stage), kb * BK, nt * BN, bb);
                    }
                    if (++stage == STAGES) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------ MMA issuer ------------------------------
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_bf16(BM, BN, A_MN, B_MN);
            constexpr uint32_t A_LBO = A_MN ? 8192u : 16u;
            constexpr uint32_t B_LBO = B_MN ? 8192u : 16u;
            constexpr uint32_t A_KSTEP = A_MN ? 2048u : 32u;   // bytes per UMMA_K = 16 elements
            constexpr uint32_t B_KSTEP = B_MN ? 2048u : 32u;
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int w = blockIdx.x; w < total_work; w += gridDim.x, ++it) {
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
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) {
                        const uint64_t ad = make_smem_desc(sa + k * A_KSTEP, A_LBO, 1024u);
                        const uint64_t bd = make_smem_desc(sb + k * B_KSTEP, B_LBO, 1024u);
                        tc_mma_bf16(d_tmem, ad, bd, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
                    }
                    tc_commit(empty_bar(stage));
                    if (++stage == STAGES) { stage = 0; phase ^= 1u; }
                }
                tc_commit(tfull_bar(as));
            }
        }
    } else {
        // ------------------------------ epilogue ------------------------------
        const int ew = warp - 2;           // 0..7
        const int q = warp & 3;            // TMEM lane quarter this warp may touch
        const int chalf = ew >> 2;         // two warps share a quarter: even / odd 32-column chunks
        const int et = threadIdx.x - 64;   // 0..255
        int it = 0;
        for (int w = blockIdx.x; w < total_work; w += gridDim.x, ++it) {
            int g, b, mt, nt, kb0, kb1;
            decode(w, g, b, mt, nt, kb0, kb1);
            const TcArgs& args = sargs[g];
            const int as = it & 1;
            const uint32_t aphase = (uint32_t)(it >> 1) & 1u;
            mbar_wait(tfull_bar(as), aphase);
            tc_fence_after();
            const int row = q * 32 + lane;
            const int m = mt * BM + row;
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BN);
            if (args.out_kind == OUT_ATOMIC_F32) {
#pragma unroll 1
                for (int c = chalf; c < BN / 32; c += 2) {
                    const int n0 = nt * BN + c * 32;
                    if (n0 >= args.N) break;
                    uint32_t r[32];
                    float v[32];
                    tmem_ld32(taddr + c * 32, r);
                    tmem_ld_wait();
                    if (m < args.M) {
                        epilogue_math32(args, r, v, b, m, n0);
                        epilogue_atomic32(args, v, b, m, n0);
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(tempty_bar(as));
                continue;
            }
            const int esize = args.out_kind == OUT_BF16 ? 2 : 4;
            const int seg_cols = 256 / esize;                                // columns staged per round
            const int rounds = (BN + seg_cols - 1) / seg_cols;
            const int round_chunks = (BN < seg_cols ? BN : seg_cols) / 32;    // 32-column chunks per round
            int valid_rounds = (args.N - nt * BN + seg_cols - 1) / seg_cols;      // rounds that hold real columns
            if (valid_rounds > rounds) valid_rounds = rounds;
#pragma unroll 1
            for (int rd = 0; rd < valid_rounds; ++rd) {
                const int col0 = nt * BN + rd * seg_cols;
                // ---- phase A: TMEM -> registers -> epilogue math -> swizzled staging tile ----
#pragma unroll 1
                for (int c = chalf; c < round_chunks; c += 2) {
                    const int n0 = col0 + c * 32;
                    if (n0 >= args.N) break;
                    uint32_t r[32];
                    float v[32];
                    tmem_ld32(taddr + rd * seg_cols + c * 32, r);
                    tmem_ld_wait();
                    epilogue_math32(args, r, v, b, m, n0);
                    if (esize == 2) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            uint32_t pk[4];
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                __nv_bfloat162 h2 = __floats2bfloat162_rn(v[j * 8 + 2 * i], v[j * 8 + 2 * i + 1]);
                                pk[i] = *reinterpret_cast<uint32_t*>(&h2);
                            }
                            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stage_addr(staging, row, c * 4 + j)),
                                         "r"(pk[0]), "r"(pk[1]), "r"(pk[2]), "r"(pk[3]) : "memory");
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stage_addr(staging, row, c * 8 + j)),
                                         "r"(__float_as_uint(v[4 * j])), "r"(__float_as_uint(v[4 * j + 1])),
                                         "r"(__float_as_uint(v[4 * j + 2])), "r"(__float_as_uint(v[4 * j + 3])) : "memory");
                    }
                }
                if (rd == valid_rounds - 1) {
                    // every tcgen05.ld of this tile has completed: hand the accumulator back to the MMA warp
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(tempty_bar(as));
                }
                asm volatile("bar.sync 1, 256;" ::: "memory");
                // ---- phase B: coalesced flush ----
                int seg_n = args.N - col0;                                       // valid columns of this round
                if (seg_n > round_chunks * 32) seg_n = round_chunks * 32;
                const int lpr = (round_chunks * 32 * esize) >> 4;                 // 16-byte lanes per row
                if (esize == 2) {
                    if (lpr == 16) flush_round<2, 16>(args, staging, cs, ew, lane, b, mt, col0, seg_n);
                    else if (lpr == 8) flush_round<2, 8>(args, staging, cs, ew, lane, b, mt, col0, seg_n);
                    else flush_round<2, 4>(args, staging, cs, ew, lane, b, mt, col0, seg_n);
                } else {
                    if (lpr == 16) flush_round<4, 16>(args, staging, cs, ew, lane, b, mt, col0, seg_n);
                    else flush_round<4, 8>(args, staging, cs, ew, lane, b, mt, col0, seg_n);
                }
                asm volatile("bar.sync 1, 256;" ::: "memory");   // staging tile reusable, column sums complete
                if (args.colsum != nullptr && et < seg_n) {
                    const float sv = cs[et];
                    cs[et] = 0.f;
                    if (sv != 0.f) atomicAdd(args.colsum + col0 + et, sv);
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
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
    uint32_t drop_thr; float drop_scale; unsigned long long drop_seed;
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
                v = dropout_keep(a.drop_seed, ((unsigned long long)b * a.M + m) * (unsigned long long)a.N + n, a.drop_thr)
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

// 3-D bf16 tensor map over a row-major [batch][rows][inner] operand, 128B swizzle.
static int make_map(CUtensorMap* map, const void* ptr, int64_t inner, int64_t rows, int64_t batch, int64_t ld,
                    int64_t bs, int box_inner, int box_rows) {
    EncodeTiledFn enc = get_encode();
    VPR_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled not available from the driver");
    VPR_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, "TMA operand pointer must be 16-byte aligned");
    VPR_REQUIRE((ld % 8) == 0, "TMA operand row stride must be a multiple of 8 bf16 (got %lld)", (long long)ld);
    VPR_REQUIRE(inner > 0 && rows > 0 && batch > 0, "empty TMA operand");
    if (batch == 1 || bs == 0) { batch = 1; bs = rows * ld; }
    VPR_REQUIRE((bs % 8) == 0, "TMA operand batch stride must be a multiple of 8 bf16 (got %lld)", (long long)bs);
    cuuint64_t dims[3] = {(cuuint64_t)inner, (cuuint64_t)rows, (cuuint64_t)batch};
    cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)bs * 2};
    cuuint32_t box[3] = {(cuuint32_t)box_inner, (cuuint32_t)box_rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    VPR_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d): inner=%lld rows=%lld batch=%lld ld=%lld bs=%lld",
                (int)r, (long long)inner, (long long)rows, (long long)batch, (long long)ld, (long long)bs);
    return 0;
}

template <bool A_MN, bool B_MN, int BN>
static int launch_tc(const TcGroup& grp, int grid, cudaStream_t st) {
    using Cfg = TcCfg<BN>;
    static bool attr_set = false;
    if (!attr_set) {
        VPR_CUDA_OK(cudaFuncSetAttribute(gemm_tc_kernel<A_MN, B_MN, BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg::SMEM_BYTES));
        attr_set = true;
    }
    gemm_tc_kernel<A_MN, B_MN, BN><<<grid, GEMM_THREADS, Cfg::SMEM_BYTES, st>>>(grp);
    return launch_ok("gemm_tc_kernel");
}

template <int BN>
static int dispatch_major(int a_mn, int b_mn, const TcGroup& grp, int grid, cudaStream_t st) {
    if (!a_mn && !b_mn) return launch_tc<false, false, BN>(grp, grid, st);
    if (!a_mn && b_mn) return launch_tc<false, true, BN>(grp, grid, st);
    if (a_mn && !b_mn) return launch_tc<true, false, BN>(grp, grid, st);
    return launch_tc<true, true, BN>(grp, grid, st);
}

static int auto_bn(int N) { return (N > 128) ? 256 : (N > 64) ? 128 : 64; }

static int gemm_tc_group(const GemmDesc* gs, int count, cudaStream_t st) {
    VPR_REQUIRE(count >= 1 && count <= MAX_GROUP, "a GEMM group holds 1..%d problems", MAX_GROUP);
    int bn = 0;
    for (int i = 0; i < count; ++i) {
        VPR_REQUIRE(gs[i].a_mn == gs[0].a_mn && gs[i].b_mn == gs[0].b_mn, "grouped GEMMs must share the operand layout");
        const int want = gs[i].bn ? gs[i].bn : auto_bn(gs[i].N);
        if (want > bn) bn = want;
    }
    VPR_REQUIRE(bn == 64 || bn == 128 || bn == 256, "bad tile N %d", bn);
    TcGroup grp;
    grp.count = count;
    long long total = 0;
    for (int i = 0; i < count; ++i) {
        const GemmDesc& g = gs[i];
        TcProblem& p = grp.p[i];
        if (g.a_mn) VPR_TRY(make_map(&p.tmA, g.A, g.M, g.K, g.batch, g.lda, g.a_bs, 64, 64));
        else        VPR_TRY(make_map(&p.tmA, g.A, g.K, g.M, g.batch, g.lda, g.a_bs, 64, BM));
        if (g.b_mn) VPR_TRY(make_map(&p.tmB, g.B, g.N, g.K, g.batch, g.ldb, g.b_bs, 64, 64));
        else        VPR_TRY(make_map(&p.tmB, g.B, g.K, g.N, g.batch, g.ldb, g.b_bs, 64, bn));
        TcArgs& a = p.args;
        a.M = g.M; a.N = g.N; a.K = g.K; a.batch = g.batch;
        a.tiles_m = (g.M + BM - 1) / BM;
        a.tiles_n = (g.N + bn - 1) / bn;
        a.kb_total = (g.K + BK - 1) / BK;
        a.k_splits = g.k_splits < 1 ? 1 : g.k_splits;
        if (a.k_splits > a.kb_total) a.k_splits = a.kb_total;
        a.b_batched = (g.batch > 1 && g.b_bs != 0) ? 1 : 0;
        a.out = g.out; a.out_kind = g.out_kind; a.ldo = g.ldo; a.out_bs = g.out_bs;
        a.bias = g.bias; a.relu = g.relu;
        a.mask = reinterpret_cast<const bf16*>(g.mask); a.ld_mask = g.ld_mask; a.mask_bs = g.mask_bs;
        a.mask_scale = g.mask_scale; a.alpha = g.alpha;
        a.drop_thr = g.dropout_p > 0.f ? dropout_threshold(g.dropout_p) : 0u;
        a.drop_scale = g.dropout_p > 0.f ? 1.f / (1.f - g.dropout_p) : 1.f; a.drop_seed = g.dropout_seed;
        a.colsum = g.colsum;
        VPR_REQUIRE(g.colsum == nullptr || g.out_kind == OUT_BF16, "fused column sums need a bf16 output");
        VPR_REQUIRE(g.mask == nullptr || g.out_kind == OUT_BF16, "the relu-mask epilogue needs a bf16 output");
        a.work_begin = (int)total;
        total += (long long)a.batch * a.tiles_m * a.tiles_n * a.k_splits;
        VPR_REQUIRE(total < (1ll << 31), "too many tiles");
    }
    grp.total_work = (int)total;
    int grid = (int)(total < num_sms() ? total : num_sms());
    if (bn == 256) return dispatch_major<256>(gs[0].a_mn, gs[0].b_mn, grp, grid, st);
    if (bn == 128) return dispatch_major<128>(gs[0].a_mn, gs[0].b_mn, grp, grid, st);
    return dispatch_major<64>(gs[0].a_mn, gs[0].b_mn, grp, grid, st);
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
