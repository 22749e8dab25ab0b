// Descriptor kNN: scores = Q G^T on tcgen05 with a fused per-row top-k epilogue (SURVEY.md
// section 8 row a10).  The [Q, G] score matrix is never written: each epilogue thread owns one
// query row of the 128-row tile, keeps that row's current k-th best score in a register as a
// threshold, and the warp inserts the (rare) survivors cooperatively into sorted lists in shared
// memory (lane = list slot).
//
// Work decomposition: one work item = (128-query tile, gallery chunk); a CTA streams every
// 256-row gallery tile of its chunk through the TMA/MMA pipeline, then writes the chunk-local
// top-k of its 128 queries to the workspace; a small merge kernel combines the chunks (and, in
// the sharded case, the per-rank results after the allgather).
#include "common.cuh"

#include <stdlib.h>
#include "../../include/vpr_b200.h"

#include <math.h>

namespace vpr {

// A CTA pair (cta_group::2) owns a 256-query x 256-gallery-row tile: each CTA stages its own 128
// query rows and half of the gallery tile (32 KB per k-block per SM instead of 48 KB), the leader
// issues the 256 x 256 x 16 MMAs, and each CTA runs the top-k epilogue for its own 128 queries.
static constexpr int KBM = 128, KBN = 256, KBK = 64;
static constexpr int K_TILE_M = 2 * KBM;                         // query rows per CTA pair
static constexpr int K_A_BYTES = KBM * KBK * 2, K_B_BYTES = (KBN / 2) * KBK * 2;
static constexpr int K_STAGE_BYTES = K_A_BYTES + K_B_BYTES;
static constexpr int K_STAGES = 6;
static constexpr int KMAX = 32;   // list capacity per row
static constexpr int K_LIST_BYTES = KMAX * KBM * 8;
static constexpr int K_SMEM_BYTES = K_STAGES * K_STAGE_BYTES + K_LIST_BYTES + 1024 + 256;
static_assert(K_SMEM_BYTES <= 227 * 1024, "kNN shared memory budget");
static constexpr int K_THREADS = 192;
static constexpr int K_W_PROD = 4, K_W_MMA = 5;        // above the 4 epilogue warps: the arbiter favours high warp ids

struct KnnArgs {
    int Q, G, D, k;
    int tiles_m, chunks, tiles_per_chunk, tiles_n, kb_total;
    int q_group;           // query tiles that walk the same gallery chunk at the same time (L2 reuse)
    int relaxed;           // epilogue warps sleep between polls while they wait for a tile
    int throttle;          // > 0: the q_group producers of a chunk keep within `throttle` k-blocks of each other
    int* sync;             // [query groups][chunks][8] progress words (zeroed by the host), used when throttle > 0
    float* ws_score;       // [Q, chunks, k]
    int* ws_idx;           // [Q, chunks, k] local gallery row, -1 = empty
};

// L2 eviction policies for the two operand streams: the group's query tiles are re-read for every gallery tile
// (keep them), a gallery tile is dead once the group's CTA pairs have pulled it through (drop it first)
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void tma_load_3d_2sm_hint(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2,
                                                     uint64_t policy) {
    asm volatile(
        "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%3, %4, %5}], [%2], %6;"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "l"(policy)
        : "memory");
}

__global__ void __launch_bounds__(K_THREADS, 1)
knn_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmG, const KnnArgs args) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_al = smem_raw + (smem_base - smem_u32(smem_raw));
    float* lscore = reinterpret_cast<float*>(smem_al + K_STAGES * K_STAGE_BYTES);      // [128][KMAX] sorted, descending
    int* lidx = reinterpret_cast<int*>(lscore + KMAX * KBM);                             // [128][KMAX]
    const uint32_t bar_base = smem_base + K_STAGES * K_STAGE_BYTES + K_LIST_BYTES;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (K_STAGES + s); };
    auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * K_STAGES + s); };
    auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * K_STAGES + 2 + s); };
    const uint32_t tmem_slot = bar_base + 8u * (2 * K_STAGES + 4);
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

    // warp index through a shuffle: branches on it are provably warp-uniform, so the producer and MMA warps below run their
    // loops convergently with uniform-register operands and one elected lane issuing (see gemm.cu)
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    const uint32_t cta_rank = cluster_ctarank();                 // 0 = leader of the pair
    const int w_first = (int)(blockIdx.x >> 1), w_step = (int)(gridDim.x >> 1);
    if (warp == K_W_PROD && lane == 0) { tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmG); }
    if (warp == K_W_MMA) {
        if (lane == 0) {
            for (int s = 0; s < K_STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
            for (int s = 0; s < 2; ++s) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), 8); }   // both CTAs' warps
            mbar_fence_init();
        }
        __syncwarp();
        tmem_alloc_2sm(tmem_slot, 512);
        tmem_relinquish_2sm();
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    const int total_work = args.tiles_m * args.chunks;

    // Work order: groups of q_group query tiles; inside a group chunk-major, so the CTAs that run
    // concurrently are (q_group query tiles) x (a few gallery chunks) walking their gallery tiles in
    // lock step: every gallery tile is fetched from HBM once per group and shared through L2, and the
    // group's query tiles (q_group * 128 * D * 2 bytes) stay L2-resident.
    int grp_idx = 0, grp_members = 1;               // of the last decoded work item (the producer's throttle uses them)
    auto tile_range = [&](int w, int& mt, int& ch, int& nt0, int& nt1) {
        const int per_group = args.q_group * args.chunks;
        const int qg = w / per_group;
        const int r = w - qg * per_group;
        int qt = args.tiles_m - qg * args.q_group;
        if (qt > args.q_group) qt = args.q_group;
        grp_idx = qg; grp_members = qt;
        ch = r / qt;
        mt = qg * args.q_group + (r - ch * qt);
        nt0 = ch * args.tiles_per_chunk;
        nt1 = nt0 + args.tiles_per_chunk;
        if (nt1 > args.tiles_n) nt1 = args.tiles_n;
    };

    if (warp == K_W_PROD) {
        {
            int stage = 0; uint32_t phase = 0;
            const uint32_t full0 = mapa_shared(full_bar(0), 0u);     // both producers credit the leader's barrier
            const uint64_t pol_q = l2_policy_evict_last(), pol_g = l2_policy_evict_first();
            for (int w = w_first; w < total_work; w += w_step) {
                int mt, ch, nt0, nt1;
                tile_range(w, mt, ch, nt0, nt1);
                const int q0 = mt * K_TILE_M + (int)cta_rank * KBM;
                // Lock step: the q_group CTA pairs that scan this gallery chunk (one per query tile of the group; the host
                // plans q_group x chunks == CTA pairs, so they all run now) pull the same gallery k-blocks through L2.
                // Left alone they drift apart -- top-k insertions are data dependent -- by more than L2 holds, and every
                // pair ends up streaming the gallery from HBM on its own (measured: 32 x the gallery at Q = 8192).  So the
                // leader's producer publishes its progress every `throttle` k-blocks and waits until every member of the
                // group has published the step before; the 6-stage operand ring hides the wait.
                const bool paced = args.throttle > 0 && cta_rank == 0 && grp_members > 1;
                volatile int* prog = paced ? args.sync + ((size_t)grp_idx * args.chunks + ch) * 8 : nullptr;
                const int me = mt - grp_idx * args.q_group;
                int kbi = 0;
                for (int nt = nt0; nt < nt1; ++nt)
                    for (int kb = 0; kb < args.kb_total; ++kb, ++kbi) {
                        if (paced && kbi % args.throttle == 0) {
                          if (elect_one()) {
                            const int step = kbi / args.throttle;
                            prog[me] = step + 1;                                     // "I am issuing step `step`"
                            if (step > 0) {
                                long long t0 = 0;
                                int spins = 0;
                                for (int m = 0; m < grp_members; ++m)                // everyone is at least issuing step - 1
                                    while (prog[m] < step) {
                                        if (++spins == 64) t0 = clock64();
                                        if (spins > 64) {
                                            __nanosleep(100);
                                            if ((spins & 1023) == 0 && clock64() - t0 > 4000000000LL) __trap();
                                        }
                                    }
                            }
                          }
                          __syncwarp();                                              // the other lanes wait for the paced one
                        }
                        mbar_wait(empty_bar(stage), phase ^ 1u);
                        const uint32_t sa = smem_base + stage * K_STAGE_BYTES, sb = sa + K_A_BYTES;
                        if (elect_one()) {
                            if (cta_rank == 0) mbar_expect_tx(full_bar(stage), 2 * K_STAGE_BYTES);
                            tma_load_3d_2sm_hint(sa, &tmQ, full0 + 8u * stage, kb * KBK, q0, 0, pol_q);
                            tma_load_3d_2sm_hint(sb, &tmG, full0 + 8u * stage, kb * KBK, nt * KBN + (int)cta_rank * (KBN / 2), 0, pol_g);
                        }
                        if (++stage == K_STAGES) { stage = 0; phase ^= 1u; }
                    }
            }
        }
    } else if (warp == K_W_MMA) {
        if (cta_rank == 0) {
            constexpr uint32_t idesc = make_idesc_bf16(K_TILE_M, KBN, false, false);
            int stage = 0; uint32_t phase = 0; int it = 0;
            for (int w = w_first; w < total_work; w += w_step) {
                int mt, ch, nt0, nt1;
                tile_range(w, mt, ch, nt0, nt1);
                for (int nt = nt0; nt < nt1; ++nt, ++it) {
                    const int as = it & 1;
                    const uint32_t aphase = (uint32_t)(it >> 1) & 1u;
                    mbar_wait(tempty_bar(as), aphase ^ 1u);
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + (uint32_t)(as * KBN);
                    for (int kb = 0; kb < args.kb_total; ++kb) {
                        mbar_wait(full_bar(stage), phase);
                        tc_fence_after();
                        const uint32_t sa = smem_base + stage * K_STAGE_BYTES, sb = sa + K_A_BYTES;
                        if (elect_one()) {
#pragma unroll
                            for (int k = 0; k < KBK / 16; ++k)
                                tc_mma_bf16_2sm(d_tmem, make_smem_desc(sa + k * 32, 16u, 1024u),
                                                make_smem_desc(sb + k * 32, 16u, 1024u), idesc, (kb > 0 || k > 0) ? 1u : 0u);
                            tc_commit_2sm(empty_bar(stage), 3);
                        }
                        if (++stage == K_STAGES) { stage = 0; phase ^= 1u; }
                    }
                    if (elect_one()) tc_commit_2sm(tfull_bar(as), 3);
                }
            }
        }
    } else {
        // Top-k epilogue.  Lane = one query row: it keeps the row's current k-th best score in a register and
        // tests its 32 scores of a chunk against it (2-3 instructions per score).  Survivors are rare
        // (k ln(G/k) per row), but a lane-private insertion loop would stall the other 31 lanes for the whole
        // shift, so insertion is warp-cooperative: the warp takes the candidates one at a time, lane s holds
        // slot s of that row's sorted list, one ballot finds the position and one shuffle does the shift.
        const int q = warp & 3;
        const int row = q * 32 + lane;              // this thread's query row within the CTA's 128
        const int kk = args.k;
        float* wl_s = lscore + (size_t)(q * 32) * KMAX;      // this warp's 32 lists: [row][KMAX]
        int* wl_i = lidx + (size_t)(q * 32) * KMAX;
        int it = 0;
        const uint32_t tempty0 = mapa_shared(tempty_bar(0), 0u);     // the MMA issuer's barrier
        for (int w = w_first; w < total_work; w += w_step) {
            int mt, ch, nt0, nt1;
            tile_range(w, mt, ch, nt0, nt1);
            for (int s = 0; s < KMAX; ++s) { wl_s[lane * KMAX + s] = -INFINITY; wl_i[lane * KMAX + s] = -1; }
            __syncwarp();
            float thr = -INFINITY;                   // current k-th best of this row
            for (int nt = nt0; nt < nt1; ++nt, ++it) {
                const int as = it & 1;
                const uint32_t aphase = (uint32_t)(it >> 1) & 1u;
                if (args.relaxed) mbar_wait_relaxed(tfull_bar(as), aphase, 256);   // a tile takes ~100 us: do not spin against the MMA warp
                else mbar_wait(tfull_bar(as), aphase);
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * KBN);
#pragma unroll 1
                for (int c = 0; c < KBN / 32; ++c) {
                    const int n0 = nt * KBN + c * 32;
                    if (n0 >= args.G) break;
                    uint32_t r[32];
                    tmem_ld32(taddr + c * 32, r);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const float v = __uint_as_float(r[i]);
                        unsigned cand = __ballot_sync(0xffffffffu, v > thr && n0 + i < args.G);
                        while (cand) {
                            const int src = __ffs(cand) - 1;
                            cand &= cand - 1;
                            const float cv = __shfl_sync(0xffffffffu, v, src);
                            float* ls = wl_s + src * KMAX;
                            int* li = wl_i + src * KMAX;
                            const float cur = ls[lane];
                            const int curi = li[lane];
                            // entries that stay ahead: score >= cv (equal scores keep the earlier gallery row first)
                            const int pos = __popc(__ballot_sync(0xffffffffu, lane < kk && cur >= cv));
                            const float up = __shfl_up_sync(0xffffffffu, cur, 1);
                            const int upi = __shfl_up_sync(0xffffffffu, curi, 1);
                            const float nv = lane > pos ? up : (lane == pos ? cv : cur);
                            if (lane < kk && lane >= pos) { ls[lane] = nv; li[lane] = lane == pos ? n0 + i : upi; }
                            const float nthr = __shfl_sync(0xffffffffu, nv, kk - 1);
                            if (lane == src) thr = nthr;
                            __syncwarp();
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(tempty0 + 8u * as);
            }
            const int qrow = mt * K_TILE_M + (int)cta_rank * KBM + row;
            if (qrow < args.Q) {
                float* os = args.ws_score + ((size_t)qrow * args.chunks + ch) * kk;
                int* oi = args.ws_idx + ((size_t)qrow * args.chunks + ch) * kk;
                for (int s = 0; s < kk; ++s) { os[s] = wl_s[lane * KMAX + s]; oi[s] = wl_i[lane * KMAX + s]; }
            }
            __syncwarp();
        }
    }
    tc_fence_before();
    cluster_sync_all();          // the leader's MMAs read the peer's shared memory and write its TMEM
    if (warp == K_W_MMA) { tc_fence_after(); tmem_dealloc_2sm(tmem_base, 512); }
}

// Merge R sorted-or-not candidate lists of k entries per query into the global top-k
// (score descending, ties -> lower index).  One warp per query, k selection rounds.
template <typename IdxT>
__global__ void __launch_bounds__(256) knn_merge_kernel(const float* __restrict__ scores, const IdxT* __restrict__ idx,
                                                        int R, long long Q, int k, long long q_stride_lists,
                                                        long long list_stride, long long idx_offset,
                                                        float* __restrict__ out_score, long long* __restrict__ out_idx) {
    // candidate (r, s) of query q lives at q*q_stride_lists + r*list_stride + s
    const long long q = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (q >= Q) return;
    const int total = R * k;
    float prev_s = INFINITY;
    long long prev_i = -1;
    for (int sel = 0; sel < k; ++sel) {
        float best_s = -INFINITY;
        long long best_i = 0x7fffffffffffffffLL;
        for (int e = lane; e < total; e += 32) {
            const int r = e / k, s = e - r * k;
            const long long off = q * q_stride_lists + (long long)r * list_stride + s;
            const long long ii = (long long)idx[off];
            if (ii < 0) continue;
            const float sc = scores[off];
            // strictly after (prev_s, prev_i) in (score desc, idx asc) order
            const bool after = (sc < prev_s) || (sc == prev_s && ii > prev_i);
            if (!after) continue;
            if (sc > best_s || (sc == best_s && ii < best_i)) { best_s = sc; best_i = ii; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float os = __shfl_xor_sync(0xffffffffu, best_s, o);
            const long long oi = __shfl_xor_sync(0xffffffffu, best_i, o);
            if (os > best_s || (os == best_s && oi < best_i)) { best_s = os; best_i = oi; }
        }
        if (lane == 0) {
            const bool found = best_i != 0x7fffffffffffffffLL;
            out_score[q * k + sel] = found ? best_s : -INFINITY;
            out_idx[q * k + sel] = found ? best_i + idx_offset : -1;
        }
        prev_s = best_s;
        prev_i = best_i;
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int knn_map(CUtensorMap* map, const void* ptr, int64_t rows, int D, int box_rows) {
    static EncodeTiledFn enc = nullptr;
    if (!enc) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
            qres != cudaDriverEntryPointSuccess)
            return fail(1, "cuTensorMapEncodeTiled not available from the driver");
        enc = reinterpret_cast<EncodeTiledFn>(p);
    }
    VPR_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, "kNN operand must be 16-byte aligned");
    cuuint64_t dims[3] = {(cuuint64_t)D, (cuuint64_t)rows, 1};
    cuuint64_t strides[2] = {(cuuint64_t)D * 2, (cuuint64_t)D * 2 * (cuuint64_t)rows};
    cuuint32_t box[3] = {64, (cuuint32_t)box_rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    VPR_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d) for kNN operand rows=%lld D=%d", (int)r,
                (long long)rows, D);
    return 0;
}

// Schedule: q_group query tiles x `chunks` gallery chunks = the CTA pairs of one launch wave, i.e. every pair keeps ONE
// gallery chunk for the whole launch and walks it once per group of query tiles, in lock step with the q_group - 1
// pairs that scan the same chunk for the other query tiles of the group.  DRAM traffic: (query tiles / q_group) gallery
// passes; the group's query tiles (q_group x 256 x D x 2 bytes, evict-last) stay in L2.
static int knn_q_group(int D, int tiles_m) {
    static long long budget_mb = -1;
    if (budget_mb < 0) { const char* e = getenv("VPR_KNN_L2_MB"); budget_mb = (e != nullptr && atoi(e) > 0) ? atoi(e) : 40; }
    long long qmax = (budget_mb << 20) / ((long long)K_TILE_M * D * 2);
    if (qmax > 8) qmax = 8;                     // beyond 8 the gallery passes saved no longer pay for the idle pairs
    if (qmax < 1) qmax = 1;
    if (qmax > tiles_m) qmax = tiles_m;
    const int pairs = num_sms() / 2;
    // among the group sizes within budget prefer the one that leaves the fewest pairs idle, then the largest
    int best = 1, best_fill = 0;
    for (int qg = 1; qg <= (int)qmax; ++qg) {
        const int fill = qg * (pairs / qg);
        if (fill > best_fill || (fill >= best_fill - 2 && qg > best)) { best_fill = fill > best_fill ? fill : best_fill; best = qg; }
    }
    return best;
}

static void knn_plan(int64_t Q, int64_t G, int D, int& tiles_m, int& tiles_n, int& chunks, int& tiles_per_chunk) {
    tiles_m = (int)((Q + K_TILE_M - 1) / K_TILE_M);
    tiles_n = (int)((G + KBN - 1) / KBN);
    const int pairs = num_sms() / 2;
    const int qg = knn_q_group(D, tiles_m);
    chunks = pairs / qg;                        // q_group x chunks work items = one wave of CTA pairs
    if (chunks > tiles_n) chunks = tiles_n;
    if (chunks < 1) chunks = 1;
    tiles_per_chunk = (tiles_n + chunks - 1) / chunks;
    chunks = (tiles_n + tiles_per_chunk - 1) / tiles_per_chunk;
}

}  // namespace vpr

using namespace vpr;

extern "C" {

size_t vpr_knn_workspace_bytes(int64_t Q, int64_t G, int D, int k) {
    int tm, tn, ch, tpc;
    knn_plan(Q, G, D, tm, tn, ch, tpc);
    size_t off = 0;
    carve_size(off, (size_t)Q * ch * k * 4);
    carve_size(off, (size_t)Q * ch * k * 4);
    carve_size(off, (size_t)(tm + 1) * ch * 8 * sizeof(int));   // lock-step progress words, 8 per (query group, chunk)
    return (off + 255) & ~size_t(255);
}

int vpr_knn_topk(const void* q, const void* g, int64_t Q, int64_t G, int D, int k, int64_t g_offset, float* out_score,
                 int64_t* out_idx, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
    VPR_TRY(::vpr::check_device());
    VPR_REQUIRE(q && g && out_score && out_idx && workspace, "null argument");
    VPR_REQUIRE(Q > 0 && G > 0, "empty kNN problem (Q=%lld, G=%lld)", (long long)Q, (long long)G);
    VPR_REQUIRE(k >= 1 && k <= KMAX, "k must be in [1, %d]", KMAX);
    VPR_REQUIRE(D % 8 == 0, "descriptor width must be a multiple of 8");
    VPR_REQUIRE(G < (1ll << 31) - KBN && Q < (1ll << 31) - KBM, "shard too large for 32-bit local indices");
    VPR_REQUIRE(workspace_bytes >= vpr_knn_workspace_bytes(Q, G, D, k), "workspace too small");
    KnnArgs a;
    a.Q = (int)Q; a.G = (int)G; a.D = D; a.k = k;
    knn_plan(Q, G, D, a.tiles_m, a.tiles_n, a.chunks, a.tiles_per_chunk);
    a.kb_total = (D + KBK - 1) / KBK;
    a.q_group = knn_q_group(D, a.tiles_m);
    {
        static int relaxed = -1;
        if (relaxed < 0) { const char* e = getenv("VPR_KNN_RELAXED"); relaxed = (e != nullptr && e[0] == '0') ? 0 : 1; }
        a.relaxed = relaxed;
    }
    size_t off = 0;
    a.ws_score = (float*)((char*)workspace + carve_size(off, (size_t)Q * a.chunks * k * 4));
    a.ws_idx = (int*)((char*)workspace + carve_size(off, (size_t)Q * a.chunks * k * 4));
    a.sync = (int*)((char*)workspace + carve_size(off, (size_t)(a.tiles_m + 1) * a.chunks * 8 * sizeof(int)));
    CUtensorMap tq, tg;
    VPR_TRY(knn_map(&tq, q, Q, D, KBM));
    VPR_TRY(knn_map(&tg, g, G, D, KBN / 2));
    static bool attr = false;
    if (!attr) {
        VPR_SMEM_OK(knn_tc_kernel, K_SMEM_BYTES);
        attr = true;
    }
    long long total = (long long)a.tiles_m * a.chunks;
    const int pairs = num_sms() / 2;
    // one wave = q_group x chunks pairs: then the members of a chunk group run at the same time and can be paced
    long long launch_pairs = (long long)a.q_group * a.chunks;
    if (launch_pairs > pairs || launch_pairs > total) launch_pairs = total < pairs ? total : pairs;
    {
        static int thr_env = -2;                   // VPR_KNN_THROTTLE: k-blocks between lock-step points, 0 = off
        if (thr_env == -2) { const char* e = getenv("VPR_KNN_THROTTLE"); thr_env = e ? atoi(e) : -1; }
        // measured (B200, Q = 8192, G = 1.25 M, D = 8448, 132 k-blocks per gallery tile): lock-step points every
        // 11 / 22 / 44 / 66 / 132 / 264 k-blocks -> 841 / 976 / 1097 / 1135 / 1186 / 1184 TFLOP/s; no pacing 991.
        // Once per gallery tile is enough to keep the group's pairs within L2 reach and cheap enough not to drain the ring.
        int per_tile = a.kb_total >= 64 ? a.kb_total : ((64 + a.kb_total - 1) / a.kb_total) * a.kb_total;
        int thr = thr_env >= 0 ? thr_env : per_tile;
        a.throttle = (launch_pairs == (long long)a.q_group * a.chunks && a.q_group > 1) ? thr : 0;
    }
    if (a.throttle > 0)
        VPR_CUDA_OK(cudaMemsetAsync(a.sync, 0, (size_t)(a.tiles_m + 1) * a.chunks * 8 * sizeof(int), stream));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(2 * launch_pairs));
    cfg.blockDim = dim3(K_THREADS);
    cfg.dynamicSmemBytes = K_SMEM_BYTES;
    cfg.stream = stream;
    cudaLaunchAttribute attrs[1];
    attrs[0].id = cudaLaunchAttributeClusterDimension;
    attrs[0].val.clusterDim.x = 2; attrs[0].val.clusterDim.y = 1; attrs[0].val.clusterDim.z = 1;
    cfg.attrs = attrs;
    cfg.numAttrs = 1;
    VPR_CUDA_OK(cudaLaunchKernelEx(&cfg, knn_tc_kernel, tq, tg, a));
    VPR_TRY(launch_ok("knn_tc_kernel"));
    const long long warps = Q;
    knn_merge_kernel<int><<<(unsigned)((warps * 32 + 255) / 256), 256, 0, stream>>>(
        a.ws_score, a.ws_idx, a.chunks, Q, k, (long long)a.chunks * k, k, g_offset, out_score, (long long*)out_idx);
    return launch_ok("knn_merge_kernel");
}

int vpr_knn_merge(const float* scores, const int64_t* idx, int R, int64_t Q, int k, float* out_score, int64_t* out_idx,
                  cudaStream_t stream) {
    VPR_TRY(::vpr::check_device());
    VPR_REQUIRE(scores && idx && out_score && out_idx && R > 0 && Q > 0 && k > 0, "bad argument");
    // input layout [R, Q, k]
    knn_merge_kernel<long long><<<(unsigned)((Q * 32 + 255) / 256), 256, 0, stream>>>(
        scores, (const long long*)idx, R, Q, k, (long long)k, (long long)Q * k, 0, out_score, (long long*)out_idx);
    return launch_ok("knn_merge_kernel");
}

}  // extern "C"
