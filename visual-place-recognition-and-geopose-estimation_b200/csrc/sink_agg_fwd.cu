// SALAD forward, last two steps in ONE launch, one CTA per image (n <= 256 tokens, m = 64, l = 128, bf16):
//
//   P   = sinkhorn(S)                                     register-tile solver of sinkhorn_rt.cuh
//   AGG = F^T P                                           128 x 64 x n, tcgen05, accumulator in TMEM
//   out = normalize([ normalize(TF) | normalize(AGG over l per cluster).flatten ])
//
// The plan never comes back from memory: the row owners of the Sinkhorn tile write it, as bf16, straight into the
// SWIZZLE_128B shared-memory layout that is the aggregation MMA's B operand (over the score tile, which is dead
// once the last iteration's barrier has passed), and to global memory only because the backward needs it.  The first
// 128-token chunk of F is requested by TMA before the first Sinkhorn load and lands while the iterations run; the
// second one goes into the dead half of the tile as soon as the iterations are over, under the plan's arithmetic.
// Compared with sinkhorn_fwd_rt + agg_fwd_kernel this removes one launch ramp, the plan's round trip and the aggregation
// kernel's own wait for its operands, and keeps two CTAs per SM (106 KB each).  Measured, it is NOT faster at the
// benchmarked batch (see sink_agg_fwd_supported below), so it is opt-in.
//
// Epilogue on all 16 warps (64 registers per thread do not hold a 64-value accumulator row): warp w drains TMEM lane
// quarter w % 4 (cluster-feature rows 32 (w % 4) ..) for clusters 16 (w / 4) .. + 15.
//
// Range guard: an image whose Sinkhorn factors leave fp32 range goes through the log-domain body as in the stand-alone
// kernel; its plan is then re-read from global memory into the operand layout.
#include "sinkhorn_rt.cuh"
#include "gemm.h"

#include <stdlib.h>

namespace vpr {

int make_map(CUtensorMap* map, const void* ptr, int esize, int64_t inner, int64_t rows, int64_t batch, int64_t ld,
             int64_t bs, int box_inner, int box_rows, const char* what);

static constexpr int SA_NT = 512;
static constexpr int SA_W_ASYNC = 4;                                   // lane 0 of this warp issues the TMA loads and the MMAs
static constexpr int SA_BOX = 128 * 128;                               // 128 token rows x 128 B
static constexpr int SA_TILE_BYTES = RT_ROWS * RT_PITCH * 4;           // score tile; later P operand (2 boxes) + F chunk 1 (2 boxes)
static constexpr int SA_VEC_BYTES = 3072;                              // beta, dcol, alpha, misc of the solver
static constexpr int SA_F0 = SA_TILE_BYTES + SA_VEC_BYTES;             // F chunk 0: two boxes (l 0-63, 64-127)
static constexpr int SA_BAR = SA_F0 + 2 * SA_BOX;
static constexpr int SA_SMEM = 1024 + SA_BAR + 64;
#define SA_EPS 1e-12f
static_assert(SA_F0 % 1024 == 0, "TMA boxes / UMMA operands need 1024-byte alignment");
static_assert(4 * SA_BOX <= SA_TILE_BYTES, "P operand + F chunk 1 must fit in the dead score tile");
static_assert(sinkrt_fwd_floats() * 4 <= (size_t)SA_F0, "solver vectors must end before F chunk 0");
static_assert(sinkfast_fwd_floats<64>(RT_ROWS) * 4 <= (size_t)SA_F0, "the guard path's scratch must not reach F chunk 0");
static_assert(2 * (SA_SMEM + 2048) <= 227 * 1024, "two CTAs per SM");

struct SaArgs {
    const float* S; const float* dust_bin;
    bf16* P; float* U; float* V;            // plan (kept for the backward) and per-iteration potentials
    float* AGG;                             // [B, 128, 64] fp32
    const float* TF;                        // [B, g] fp32
    bf16* out;                              // [B, g + 8192]
    int n, g, iters;
    SinkConst2 k;
};

__device__ __forceinline__ void tmem_ld16(uint32_t addr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32"
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(addr)
        : "memory");
}

// Sum v[0..16) over the 32 lanes; every lane returns the total of the value with index sa_col_of_lane(lane).
__device__ __forceinline__ int sa_col_of_lane(int lane) { return (lane >> 1) & 15; }
__device__ __forceinline__ float sa_colsum16(float (&v)[16], int lane) {
#pragma unroll
    for (int w = 8, bit = 16; w >= 1; w >>= 1, bit >>= 1) {
        const bool up = (lane & bit) != 0;
#pragma unroll
        for (int i = 0; i < w; ++i) {
            const float keep = up ? v[i + w] : v[i];
            const float send = up ? v[i] : v[i + w];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, bit);
        }
    }
    return v[0] + __shfl_xor_sync(0xffffffffu, v[0], 1);
}

// The Sinkhorn phase as a function of its own (not inlined): inside the kernel body the register allocator, with the
// epilogue's and the operand bookkeeping's values also live, kept the 32-entry tile row in local memory and re-read it in
// every iteration; here it gets the same 64 registers the stand-alone kernel has.  Iterates, then (unless the range guard
// fired) requests F chunk 1 into the dead half of the tile and writes the plan to global memory and into the MMA operand.
__device__ __noinline__ bool sa_solve_and_place(float* sk, const SaArgs& a, const CUtensorMap* tmF, uint32_t base, int chunks) {
    RtFwdState<SA_NT> st;
    if (rt_fwd_iterate<SA_NT>(sk, a.S, a.dust_bin, a.U, a.V, a.n, a.iters, a.k, st)) return true;
    // the tile's second half is handed from the generic proxy (the solver's stores) to the async one (TMA writes)
    fence_proxy_async_smem();
    __syncthreads();
    if (chunks > 1 && threadIdx.x == SA_W_ASYNC * 32) {
        const uint32_t fullF1 = base + SA_BAR + 8;
        mbar_expect_tx(fullF1, 2 * SA_BOX);
        tma_load_3d(base + 2 * SA_BOX, tmF, fullF1, 0, 128, (int)blockIdx.x);
        tma_load_3d(base + 3 * SA_BOX, tmF, fullF1, 64, 128, (int)blockIdx.x);
    }
    rt_fwd_plan<bf16, SA_NT, true>(sk + RT_ROWS * RT_PITCH + 2 * RT_ROWS, st, a.P, a.n, reinterpret_cast<uint8_t*>(sk), chunks * 128);
    return false;
}

__global__ void __launch_bounds__(SA_NT, 2)
sink_agg_fwd_kernel(const __grid_constant__ CUtensorMap tmF, const __grid_constant__ SaArgs a) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ float part[4][64];            // per lane-quarter partial sums of squares per cluster
    __shared__ float inv_c[64];
    __shared__ float wred[SA_NT / 32];
    __shared__ float sc[2];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_al = smem_raw + (base - smem_u32(smem_raw));
    float* sk = reinterpret_cast<float*>(smem_al);
    const uint32_t fullF0 = base + SA_BAR, fullF1 = fullF0 + 8, tfull = fullF0 + 16, tmem_slot = fullF0 + 24;
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_al + SA_BAR + 24);

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);      // provably warp-uniform (see gemm.cu)
    const int b = blockIdx.x;
    const int chunks = (a.n + 127) / 128;

    if (warp == SA_W_ASYNC) {
        if (lane == 0) {
            tma_prefetch_desc(&tmF);
            mbar_init(fullF0, 1); mbar_init(fullF1, 1); mbar_init(tfull, 1);
            mbar_fence_init();
        }
        __syncwarp();
        tmem_alloc(tmem_slot, 64);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    pdl_wait();                  // everything above overlapped the previous launch's tail
    pdl_launch_dependents();

    if (warp == SA_W_ASYNC && lane == 0) {                      // F chunk 0 lands while the Sinkhorn iterations run
        mbar_expect_tx(fullF0, 2 * SA_BOX);
        tma_load_3d(base + SA_F0, &tmF, fullF0, 0, 0, b);       // rows beyond n are zero-filled
        tma_load_3d(base + SA_F0 + SA_BOX, &tmF, fullF0, 64, 0, b);
    }
    const bool guard = sa_solve_and_place(sk, a, &tmF, base, chunks);
    if (guard) {                                                 // block-uniform: out of fp32 range -> exact path
        uint8_t* op = smem_al;                                   // P operand: [chunk][128 tokens][128 B], over the dead tile
        __syncthreads();
        sinkhorn_fwd_logdomain<bf16, 64, SA_NT>(sk, a.S, a.dust_bin, a.P, a.U, a.V, a.n, a.iters, a.k);
        __syncthreads();                                         // this image's plan is complete in global memory
        bf16* Pb = a.P + (size_t)b * a.n * 64;
        for (int idx = tid; idx < chunks * 128 * 8; idx += SA_NT) {
            const int j = idx >> 3, q = idx & 7;
            uint4 v = make_uint4(0u, 0u, 0u, 0u);
            if (j < a.n) v = *reinterpret_cast<uint4*>(Pb + (size_t)j * 64 + q * 8);
            *reinterpret_cast<uint4*>(op + rt_plan_operand_offset(j, q)) = v;
        }
        fence_proxy_async_smem();
        __syncthreads();
        if (chunks > 1 && warp == SA_W_ASYNC && lane == 0) {
            mbar_expect_tx(fullF1, 2 * SA_BOX);
            tma_load_3d(base + 2 * SA_BOX, &tmF, fullF1, 0, 128, b);
            tma_load_3d(base + 3 * SA_BOX, &tmF, fullF1, 64, 128, b);
        }
    }
    fence_proxy_async_smem();                                    // the operand bytes become visible to the tensor core
    __syncthreads();
    const uint32_t tmem_base = *tmem_slot_ptr;                   // (read here: nothing extra stays live across the iterations)
    if (warp == SA_W_ASYNC && lane == 0) {
        constexpr uint32_t idesc = make_idesc_bf16(128, 64, true, true);
        tc_fence_after();
        for (int ch = 0; ch < chunks; ++ch) {
            mbar_wait(ch == 0 ? fullF0 : fullF1, 0);
            tc_fence_after();
            const uint32_t fa = ch == 0 ? base + SA_F0 : base + 2 * SA_BOX;
            const uint32_t pb = base + (uint32_t)ch * SA_BOX;
#pragma unroll
            for (int ks = 0; ks < 8; ++ks)
                tc_mma_bf16(tmem_base, make_smem_desc(fa + ks * 2048, SA_BOX, 1024u), make_smem_desc(pb + ks * 2048, SA_BOX, 1024u),
                            idesc, (ch > 0 || ks > 0) ? 1u : 0u);
        }
        tc_commit(tfull);
    }
    __syncwarp();
    // ------------------------------ epilogue: thread = (cluster-feature row l, 16 clusters) ------------------------------
    {
        const int quarter = warp & 3, cgp = warp >> 2;
        const int l = quarter * 32 + lane, c0 = cgp * 16;
        // token part first (independent of the MMA): ||TF||^2
        const float* t = a.TF + (size_t)b * a.g;
        float ts = 0.f;
        for (int i = tid; i < a.g; i += SA_NT) { const float x = __ldg(t + i); ts = fmaf(x, x, ts); }
        ts = warp_sum(ts);
        if (lane == 0) wred[warp] = ts;
        mbar_wait(tfull, 0);
        tc_fence_after();
        float v[16];
        {
            uint32_t r[16];
            tmem_ld16(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)c0, r);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
        }
        float* ag = a.AGG + ((size_t)b * 128 + l) * 64 + c0;
#pragma unroll
        for (int i = 0; i < 4; ++i) *reinterpret_cast<float4*>(ag + 4 * i) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
        {   // per-cluster sums of squares over the 32 rows of this warp
            float q[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) q[i] = v[i] * v[i];
            const float tot = sa_colsum16(q, lane);
            if ((lane & 1) == 0) part[quarter][c0 + sa_col_of_lane(lane)] = tot;
        }
        __syncthreads();
        if (warp == 0) {
            float zz = 0.f;
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const int c = lane + 32 * k;
                const float tot = part[0][c] + part[1][c] + part[2][c] + part[3][c];
                const float nc = fmaxf(sqrtf(tot), SA_EPS);
                inv_c[c] = 1.f / nc;
                zz += tot / (nc * nc);
            }
            zz = warp_sum(zz);
            float tsum = 0.f;
#pragma unroll
            for (int w = 0; w < SA_NT / 32; ++w) tsum += wred[w];
            const float nt = fmaxf(sqrtf(tsum), SA_EPS);
            zz += tsum / (nt * nt);
            const float invN = 1.f / fmaxf(sqrtf(zz), SA_EPS);
            if (lane == 0) { sc[0] = invN; sc[1] = invN / nt; }
        }
        __syncthreads();
        const float invN = sc[0], invt = sc[1];
        bf16* o = a.out + (size_t)b * (a.g + 8192);
        bf16* orow = o + a.g + (size_t)l * 64 + c0;
#pragma unroll
        for (int i = 0; i < 2; ++i) {            // 2 x 16 bytes = this thread's 16 normalised values
            uint32_t pk[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int c = 8 * i + 2 * k;
                __nv_bfloat162 h2 = __floats2bfloat162_rn(v[c] * inv_c[c0 + c] * invN, v[c + 1] * inv_c[c0 + c + 1] * invN);
                pk[k] = *reinterpret_cast<uint32_t*>(&h2);
            }
            *reinterpret_cast<uint4*>(orow + 8 * i) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        }
        for (int i = tid; i < a.g; i += SA_NT) o[i] = __float2bfloat16_rn(__ldg(t + i) * invt);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == SA_W_ASYNC) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 64);
    }
}

bool sink_agg_fwd_supported(int dtype, int l, int m, int g, int n, int iters) {
    // Opt-in (VPR_SINKAGG_FUSED=1, read at every call): measured on B200 at B = 256, n = 256 inside the step, this launch
    // takes 29.2 us against 14.9 + 10.9 us for sinkhorn_fwd_rt + agg_fwd_kernel (profiles/r2_kernels_ncu_summary.md): the
    // F chunk requested up front competes with the score load every CTA is waiting on, and the aggregation tail on 64
    // registers per thread is no shorter than the stand-alone kernel's.  Kept for batches below one wave and as the
    // parity-tested statement of "the plan never leaves the SM".
    const char* e = getenv("VPR_SINKAGG_FUSED");
    const bool enabled = e != nullptr && e[0] == '1';
    const char* r = getenv("VPR_SINK_RT");
    const bool rt512 = !(r != nullptr && (r[0] == '0' || r[0] == '2'));
    return enabled && rt512 && dtype == DT_BF16 && l == 128 && m == 64 && g % 8 == 0 && n > m && n <= RT_ROWS && iters >= 1 &&
           iters <= 8;
}

int sink_agg_fwd_run(const SinkAggFwdDesc& d, cudaStream_t st) {
    CUtensorMap tmF;
    VPR_TRY(make_map(&tmF, d.F, 2, 128, d.n, d.B, 128, (int64_t)d.n * 128, 64, 128, "F"));
    SaArgs a;
    a.S = d.S; a.dust_bin = d.dust_bin; a.P = reinterpret_cast<bf16*>(d.P); a.U = d.U; a.V = d.V;
    a.AGG = d.AGG; a.TF = d.TF; a.out = reinterpret_cast<bf16*>(d.out);
    a.n = d.n; a.g = d.g; a.iters = d.iters;
    a.k.norm2 = (float)(-log2((double)(d.n + 64)));
    a.k.log_a_m2 = a.k.norm2 + (float)log2((double)(d.n - 64));
    VPR_SMEM_OK(sink_agg_fwd_kernel, SA_SMEM);
    VPR_CUDA_OK(launch_ex(sink_agg_fwd_kernel, dim3((unsigned)d.B), dim3(SA_NT), (size_t)SA_SMEM, st, d.pdl != 0, 1, tmF, a));
    return launch_ok("sink_agg_fwd_kernel");
}

}  // namespace vpr
