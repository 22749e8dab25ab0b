// Persistent, warp-specialised bf16 GEMM for sm_100a:
//   warp 0      : TMA producer (cp.async.bulk.tensor, SWIZZLE_128B tiles into a smem ring)
//   warp 1      : TMEM allocator + single-thread tcgen05.mma issuer (fp32 accumulators in TMEM,
//                 double-buffered so the epilogue of tile i overlaps the MMAs of tile i+1)
//   warps 2..5  : epilogue (tcgen05.ld -> bias / relu / relu-mask -> bf16 | fp32 | fp32 atomics)
// Operands may be K-major or MN-major in global memory (all four combinations), batched
// through the third tensor-map dimension, and split along K for the weight-gradient shapes.
#include "common.cuh"
#include "gemm.h"

#include <mutex>

namespace vpr {

// ---------------------------------------------------------------------------
// tcgen05 kernel
// ---------------------------------------------------------------------------
static constexpr int BM = 128;
static constexpr int BK = 64;                 // 64 bf16 = 128 B = one swizzle span
static constexpr int A_STAGE_BYTES = BM * BK * 2;
static constexpr int GEMM_THREADS = 320;      // TMA warp, MMA warp, 8 epilogue warps
static constexpr int EPI_WARPS = 8;
static constexpr int SMEM_BUDGET = 196608;    // operand ring; barriers live above it

struct TcArgs {
    int M, N, K, batch;
    int tiles_m, tiles_n, k_splits, kb_total;
    int b_batched;
    void* out;
    int out_kind;
    long long ldo, out_bs;
    const float* bias;
    int relu;
    const bf16* mask;
    long long ld_mask, mask_bs;
    float mask_scale, alpha;
    uint32_t drop_thr; float drop_scale; unsigned long long drop_seed;
};

template <int BN>
struct TcCfg {
    static constexpr int B_STAGE_BYTES = BN * BK * 2;
    static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
    static constexpr int STAGES = SMEM_BUDGET / STAGE_BYTES;
    static constexpr int TMEM_COLS = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128
                                     : (2 * BN <= 256) ? 256 : 512;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
};

__device__ __forceinline__ void epilogue_store32(const TcArgs& a, const uint32_t (&r)[32], int b,
                                                 int m, int n0) {
    // one thread = one output row m, 32 consecutive columns starting at n0
    const int N = a.N;
    const bool full = (n0 + 32 <= N);
    float v[32];
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
    if (a.mask != nullptr) {
        const bf16* mp = a.mask + (long long)b * a.mask_bs + (long long)m * a.ld_mask + n0;
        if (full && ((a.ld_mask & 7) == 0)) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                uint4 q = __ldg(reinterpret_cast<const uint4*>(mp) + j);
                const bf16* h = reinterpret_cast<const bf16*>(&q);
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    v[j * 8 + i] = (__bfloat162float(h[i]) > 0.f) ? v[j * 8 + i] * a.mask_scale : 0.f;
            }
        } else {
#pragma unroll
            for (int i = 0; i < 32; ++i)
                if (n0 + i < N) v[i] = (__bfloat162float(mp[i]) > 0.f) ? v[i] * a.mask_scale : 0.f;
        }
    }
    const long long off = (long long)b * a.out_bs + (long long)m * a.ldo + n0;
    if (a.out_kind == OUT_BF16) {
        bf16* op = reinterpret_cast<bf16*>(a.out) + off;
        if (full && ((a.ldo & 7) == 0)) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                uint4 q;
                __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&q);
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    h[i] = __floats2bfloat162_rn(v[j * 8 + 2 * i], v[j * 8 + 2 * i + 1]);
                reinterpret_cast<uint4*>(op)[j] = q;
            }
        } else {
#pragma unroll
            for (int i = 0; i < 32; ++i)
                if (n0 + i < N) op[i] = __float2bfloat16_rn(v[i]);
        }
    } else if (a.out_kind == OUT_F32) {
        float* op = reinterpret_cast<float*>(a.out) + off;
        if (full && ((a.ldo & 3) == 0)) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
                reinterpret_cast<float4*>(op)[j] =
                    make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        } else {
#pragma unroll
            for (int i = 0; i < 32; ++i)
                if (n0 + i < N) op[i] = v[i];
        }
    } else {
        float* op = reinterpret_cast<float*>(a.out) + off;
        if (full && ((a.ldo & 3) == 0)) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(op + 4 * j), "f"(v[4 * j]),
                             "f"(v[4 * j + 1]), "f"(v[4 * j + 2]), "f"(v[4 * j + 3])
                             : "memory");
        } else {
#pragma unroll
            for (int i = 0; i < 32; ++i)
                if (n0 + i < N) atomicAdd(op + i, v[i]);
        }
    }
}

template <bool A_MN, bool B_MN, int BN>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const TcArgs args) {
    using Cfg = TcCfg<BN>;
    constexpr int STAGES = Cfg::STAGES;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar_base = smem_base + STAGES * Cfg::STAGE_BYTES;
    // barrier addresses
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
    auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + s); };
    auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + 2 + s); };
    const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 4);
    volatile uint32_t* tmem_slot_ptr =
        reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
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
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    const int total_work = args.batch * args.tiles_m * args.tiles_n * args.k_splits;

    auto decode = [&](int w, int& b, int& mt, int& nt, int& kb0, int& kb1) {
        nt = w % args.tiles_n;
        w /= args.tiles_n;
        int ks = w % args.k_splits;
        w /= args.k_splits;
        mt = w % args.tiles_m;
        b = w / args.tiles_m;
        kb0 = (int)((long long)ks * args.kb_total / args.k_splits);
        kb1 = (int)((long long)(ks + 1) * args.kb_total / args.k_splits);
    };

    if (warp == 0) {
        // ------------------------------ TMA producer ------------------------------
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
                int b, mt, nt, kb0, kb1;
                decode(w, b, mt, nt, kb0, kb1);
                const int bb = args.b_batched ? b : 0;
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(empty_bar(stage), phase ^ 1u);
                    const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES;
                    const uint32_t sb = sa + A_STAGE_BYTES;
                    mbar_expect_tx(full_bar(stage), Cfg::STAGE_BYTES);
                    if (A_MN) {
#pragma unroll
                        for (int i = 0; i < BM / 64; ++i)
                            tma_load_3d(sa + i * 8192, &tmA, full_bar(stage), mt * BM + i * 64, kb * BK, b);
                    } else {
                        tma_load_3d(sa, &tmA, full_bar(stage), kb * BK, mt * BM, b);
                    }
                    if (B_MN) {
#pragma unroll
                        for (int i = 0; i < BN / 64; ++i)
                            tma_load_3d(sb + i * 8192, &tmB, full_bar(stage), nt * BN + i * 64, kb * BK, bb);
                    } else {
                        tma_load_3d(sb, &tmB, full_bar(stage), kb * BK, nt * BN, bb);
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
                int b, mt, nt, kb0, kb1;
                decode(w, b, mt, nt, kb0, kb1);
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
        const int q = warp & 3;            // TMEM lane quarter this warp may touch
        const int chalf = (warp - 2) >> 2;   // two warps share a quarter: even / odd 32-column chunks
        int it = 0;
        for (int w = blockIdx.x; w < total_work; w += gridDim.x, ++it) {
            int b, mt, nt, kb0, kb1;
            decode(w, b, mt, nt, kb0, kb1);
            const int as = it & 1;
            const uint32_t aphase = (uint32_t)(it >> 1) & 1u;
            mbar_wait(tfull_bar(as), aphase);
            tc_fence_after();
            const int m = mt * BM + q * 32 + lane;
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BN);
#pragma unroll 1
            for (int c = chalf; c < BN / 32; c += 2) {
                const int n0 = nt * BN + c * 32;
                if (n0 >= args.N) break;
                uint32_t r[32];
                tmem_ld32(taddr + c * 32, r);
                tmem_ld_wait();
                if (m < args.M) epilogue_store32(args, r, b, m, n0);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar(as));
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
static int launch_tc(const CUtensorMap& ta, const CUtensorMap& tb, const TcArgs& args, int grid, cudaStream_t st) {
    using Cfg = TcCfg<BN>;
    static bool attr_set = false;
    if (!attr_set) {
        VPR_CUDA_OK(cudaFuncSetAttribute(gemm_tc_kernel<A_MN, B_MN, BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg::SMEM_BYTES));
        attr_set = true;
    }
    gemm_tc_kernel<A_MN, B_MN, BN><<<grid, GEMM_THREADS, Cfg::SMEM_BYTES, st>>>(ta, tb, args);
    return launch_ok("gemm_tc_kernel");
}

template <int BN>
static int dispatch_major(int a_mn, int b_mn, const CUtensorMap& ta, const CUtensorMap& tb, const TcArgs& args,
                          int grid, cudaStream_t st) {
    if (!a_mn && !b_mn) return launch_tc<false, false, BN>(ta, tb, args, grid, st);
    if (!a_mn && b_mn) return launch_tc<false, true, BN>(ta, tb, args, grid, st);
    if (a_mn && !b_mn) return launch_tc<true, false, BN>(ta, tb, args, grid, st);
    return launch_tc<true, true, BN>(ta, tb, args, grid, st);
}

static int gemm_tc(const GemmDesc& g, cudaStream_t st) {
    int bn = g.bn;
    if (bn == 0) bn = (g.N > 128) ? 256 : (g.N > 64) ? 128 : 64;
    VPR_REQUIRE(bn == 64 || bn == 128 || bn == 256, "bad tile N %d", bn);
    CUtensorMap ta, tb;
    if (g.a_mn) VPR_TRY(make_map(&ta, g.A, g.M, g.K, g.batch, g.lda, g.a_bs, 64, 64));
    else        VPR_TRY(make_map(&ta, g.A, g.K, g.M, g.batch, g.lda, g.a_bs, 64, BM));
    if (g.b_mn) VPR_TRY(make_map(&tb, g.B, g.N, g.K, g.batch, g.ldb, g.b_bs, 64, 64));
    else        VPR_TRY(make_map(&tb, g.B, g.K, g.N, g.batch, g.ldb, g.b_bs, 64, bn));
    TcArgs a;
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
    long long total = (long long)a.batch * a.tiles_m * a.tiles_n * a.k_splits;
    VPR_REQUIRE(total < (1ll << 31), "too many tiles");
    int grid = (int)(total < num_sms() ? total : num_sms());
    if (bn == 256) return dispatch_major<256>(g.a_mn, g.b_mn, ta, tb, a, grid, st);
    if (bn == 128) return dispatch_major<128>(g.a_mn, g.b_mn, ta, tb, a, grid, st);
    return dispatch_major<64>(g.a_mn, g.b_mn, ta, tb, a, grid, st);
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
    return launch_ok("gemm_simt_kernel");
}

int gemm_run(const GemmDesc& g, cudaStream_t st) {
    VPR_REQUIRE(g.M > 0 && g.N > 0 && g.K > 0 && g.batch > 0, "empty GEMM %d x %d x %d (batch %d)", g.M, g.N, g.K, g.batch);
    VPR_REQUIRE(g.k_splits <= 1 || g.out_kind == OUT_ATOMIC_F32, "split-K needs the atomic epilogue");
    if (g.dtype == DT_BF16) return gemm_tc(g, st);
    return gemm_simt(g, st);
}

}  // namespace vpr
