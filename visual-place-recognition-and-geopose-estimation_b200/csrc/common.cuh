// Shared device/host helpers for the sm_100a kernels: PTX wrappers for mbarrier,
// TMA (cp.async.bulk.tensor), tcgen05 (alloc / mma / commit / ld) and the
// library-wide error channel.  Everything here is written for sm_100a only.
#pragma once

#include <cuda.h>          // CUtensorMap (types only; the encoder is fetched at run time)
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace vpr {

typedef __nv_bfloat16 bf16;

// ---------------------------------------------------------------------------
// error channel (api.cu)
// ---------------------------------------------------------------------------
int fail(int code, const char* fmt, ...);
int num_sms();
// One process drives one device: the first compute entry binds the library (cached SM count, kernel attributes,
// side stream) to the current device; later calls fail if another device is current.
int check_device();
void note_launch();   // counts kernel launches (vpr_launch_count)

#define VPR_CUDA_OK(expr)                                                          \
    do {                                                                           \
        cudaError_t e__ = (expr);                                                  \
        if (e__ != cudaSuccess)                                                    \
            return ::vpr::fail(2, "%s:%d %s -> %s", __FILE__, __LINE__, #expr,     \
                               cudaGetErrorString(e__));                           \
    } while (0)

#define VPR_TRY(expr)                  \
    do {                               \
        int r__ = (expr);              \
        if (r__ != 0) return r__;      \
    } while (0)

#define VPR_REQUIRE(cond, ...)                                   \
    do {                                                         \
        if (!(cond)) return ::vpr::fail(1, __VA_ARGS__);         \
    } while (0)

// cudaFuncAttributeMaxDynamicSharedMemorySize only has to grow: remember the largest value set per kernel
// so that steady-state launches skip the driver call (api.cu)
int ensure_dynamic_smem(const void* kernel, size_t bytes);
#define VPR_SMEM_OK(kernel, bytes) VPR_TRY(::vpr::ensure_dynamic_smem(reinterpret_cast<const void*>(kernel), (bytes)))

// A side stream + fork/join events (thread-local, created on first use) for small launches that are off the
// critical path of a call: fork after the producer on the caller's stream, join before the consumer.  Works
// under stream capture (the side stream joins the capture and becomes a parallel branch of the graph).
bool pdl_enabled();          // VPR_PDL=0 switches programmatic dependent launches off

// kernel launch with optional cluster width and programmatic-dependent-launch attribute
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_ex(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl,
                                    int cluster, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    int na = 0;
    if (cluster > 1) {
        attr[na].id = cudaLaunchAttributeClusterDimension;
        attr[na].val.clusterDim.x = (unsigned)cluster;
        attr[na].val.clusterDim.y = 1;
        attr[na].val.clusterDim.z = 1;
        ++na;
    }
    if (pdl && pdl_enabled()) {
        attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    cfg.attrs = attr;
    cfg.numAttrs = (unsigned)na;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

struct SideBranch { cudaStream_t stream; cudaEvent_t fork, join, fork2, zeroed; };
int side_branch(SideBranch** out);

static inline int launch_ok(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(3, "launch %s: %s", what, cudaGetErrorString(e));
    note_launch();
    return 0;
}

template <typename T>
static inline T* carve(char*& p, size_t count) {
    uintptr_t a = (reinterpret_cast<uintptr_t>(p) + 255) & ~uintptr_t(255);
    T* r = reinterpret_cast<T*>(a);
    p = reinterpret_cast<char*>(a + count * sizeof(T));
    return r;
}
static inline size_t carve_size(size_t& off, size_t bytes) {
    off = (off + 255) & ~size_t(255);
    size_t r = off;
    off += bytes;
    return r;
}

// Counter-based dropout.  Element indices are grouped in runs of 32; a 32-bit avalanche mix of
// (seed, run) seeds the full-period LCG x -> A x + C (mod 2^32) and element i of the run is kept iff
// the LCG's i-th state is >= p * 2^32.  State i is the affine map A^i x0 + C_i of the run's seed, so
// a thread that owns a whole run pays one multiply-add per element (the GEMM epilogue), while
// anything else can evaluate single elements.
static constexpr uint32_t DROP_LCG_A = 1664525u, DROP_LCG_C = 1013904223u;
__host__ __device__ constexpr uint32_t drop_lcg_mul(int i) {
    uint32_t a = 1u;
    for (int k = 0; k < i; ++k) a *= DROP_LCG_A;
    return a;
}
__host__ __device__ constexpr uint32_t drop_lcg_add(int i) {
    uint32_t c = 0u;
    for (int k = 0; k < i; ++k) c = c * DROP_LCG_A + DROP_LCG_C;
    return c;
}
__host__ __device__ inline uint32_t dropout_mix(uint64_t seed, uint64_t run) {
    uint32_t h = (uint32_t)run * 0x9E3779B1u + (uint32_t)seed;
    h ^= h >> 16; h *= 0x85EBCA6Bu;
    h ^= (uint32_t)(run >> 32) * 0xC2B2AE35u + (uint32_t)(seed >> 32);
    h ^= h >> 13; h *= 0xC2B2AE35u;
    h ^= h >> 16;
    return h;
}
__host__ __device__ inline bool dropout_keep(uint64_t seed, uint64_t idx, uint32_t thr32) {
    uint32_t w = dropout_mix(seed, idx >> 5);
    const int i = (int)(idx & 31);
    for (int k = 0; k < i; ++k) w = w * DROP_LCG_A + DROP_LCG_C;
    return w >= thr32;
}
static inline uint32_t dropout_threshold(float p) {
    double t = (double)p * 4294967296.0;
    if (t < 1.0) t = 1.0;                    // 0 is reserved for "no dropout"
    if (t > 4294967295.0) t = 4294967295.0;
    return (uint32_t)t;
}

#ifdef __CUDACC__
// ---------------------------------------------------------------------------
// small device utilities
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31; }

__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b32 r;\n\t"
        "elect.sync r|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(pred));
    return pred != 0;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// ---------------------------------------------------------------------------
// mbarrier
// ---------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// Bounded wait: a wedged pipeline traps (the launch fails with an error) instead of
// hanging the GPU box.  ~4e9 SM cycles is two orders of magnitude above any
// legitimate wait in these kernels.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    long long t0 = 0;
    int spins = 0;
    while (true) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}\n"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (done) return;
        if (++spins == 64) t0 = clock64();
        if (spins > 64 && (spins & 1023) == 0 && clock64() - t0 > 4000000000LL) __trap();
    }
}

// Same wait for warps that expect to wait long (epilogue warps waiting ~100 us for a kNN tile): after a few failed
// polls they sleep between polls instead of competing with the MMA / TMA warps for issue slots and power.
__device__ __forceinline__ void mbar_wait_relaxed(uint32_t bar, uint32_t parity, unsigned sleep_ns) {
    uint32_t done;
    long long t0 = 0;
    int spins = 0;
    while (true) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}\n"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (done) return;
        if (++spins > 4) __nanosleep(sleep_ns);
        if (spins == 64) t0 = clock64();
        if (spins > 64 && (spins & 255) == 0 && clock64() - t0 > 4000000000LL) __trap();
    }
}

// ---------------------------------------------------------------------------
// TMA
// ---------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* m, uint32_t bar,
                                            int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cta.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}

// L2 eviction policies for streams whose reuse is known: evict-first for bytes that are read once, evict-last for
// bytes the next launch reads back.
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_evict_last_policy() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void tma_load_3d_hint(uint32_t dst, const CUtensorMap* m, uint32_t bar,
                                                 int c0, int c1, int c2, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cta.global.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%3, %4, %5}], [%2], %6;"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "l"(policy)
        : "memory");
}

// CTA-pair variant: executed by both CTAs of the pair for their own half of the operands; the
// transaction bytes are credited to `bar`, a shared::cluster address (the leader CTA's barrier).
__device__ __forceinline__ void tma_load_3d_2sm(uint32_t dst, const CUtensorMap* m, uint32_t bar,
                                                int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
// multicast load: the box lands at the same shared-memory offset in every CTA of `cta_mask`, and each of those
// CTAs' barriers (same offset) is credited with the bytes
__device__ __forceinline__ void tma_load_3d_mc(uint32_t dst, const CUtensorMap* m, uint32_t bar,
                                               int c0, int c1, int c2, uint16_t cta_mask) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
        " [%0], [%1, {%3, %4, %5}], [%2], %6;"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "h"(cta_mask)
        : "memory");
}
// ---------------------------------------------------------------------------
// programmatic dependent launch: a kernel launched with the programmatic-serialization attribute may become resident
// while its predecessor in the stream is still running; everything it does before pdl_wait() (barrier init, TMEM
// allocation, descriptor prefetch) overlaps the predecessor's tail.  pdl_wait() returns once the predecessor grid has
// completed and its writes are visible (a no-op for a normally launched kernel).  Every kernel here calls
// pdl_launch_dependents() only AFTER its own pdl_wait(), so completion is transitive along a chain.
// ---------------------------------------------------------------------------
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ---------------------------------------------------------------------------
// thread-block clusters
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
// shared::cta address of this CTA -> shared::cluster address of the same location in CTA `rank`
__device__ __forceinline__ uint32_t mapa_shared(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on an mbarrier that lives in another CTA of the cluster (shared::cluster address)
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_bar) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_bar) : "memory");
}

// smem -> global tile store; completion is tracked by the issuing thread's bulk async-groups
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* m, uint32_t src, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
        ::"l"(reinterpret_cast<uint64_t>(m)), "r"(src), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tma_store_3d_hint(const CUtensorMap* m, uint32_t src, int c0, int c1, int c2, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.tensor.3d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%2, %3, %4}], [%1], %5;"
        ::"l"(reinterpret_cast<uint64_t>(m)), "r"(src), "r"(c0), "r"(c1), "r"(c2), "l"(policy)
        : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {      // at most N groups still reading shared memory
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
// generic-proxy writes to shared memory -> visible to the async proxy (TMA store)
__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int threads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

// ---------------------------------------------------------------------------
// tcgen05 / TMEM
// ---------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem),
                 "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// tcgen05.commit: arrives (count 1) on the mbarrier once all MMAs issued so far by
// this thread have completed.  Implies fence::before_thread_sync.
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
                 : "memory");
}
// CTA-pair (cta_group::2) flavours.  alloc / dealloc are executed by the same warp of BOTH CTAs;
// the MMA and its commits are issued by the leader CTA only and cover both CTAs' shared memory / TMEM.
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void tmem_relinquish_2sm() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t addr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols) : "memory");
}
// arrives (count 1) on the barrier at this shared-memory offset in every CTA of `cta_mask`
// single-CTA MMAs, arrival delivered to the barrier at this offset in every CTA of `cta_mask`
__device__ __forceinline__ void tc_commit_mc(uint32_t bar, uint16_t cta_mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"(cta_mask)
                 : "memory");
}
__device__ __forceinline__ void tc_commit_2sm(uint32_t bar, uint16_t cta_mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"(cta_mask)
                 : "memory");
}
__device__ __forceinline__ void tc_mma_bf16_2sm(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                                uint32_t idesc, uint32_t accumulate) {
    const uint32_t z = 0u;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}\n"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(z)
        : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc], kind::f16 (bf16 in, fp32 accumulate).
__device__ __forceinline__ void tc_mma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                            uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// 32 lanes x 32 consecutive fp32 columns -> 32 registers per thread (thread i = lane i of
// the warp's TMEM lane quarter).
__device__ __forceinline__ void tmem_ld32(uint32_t addr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32"
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15,"
        " %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31},"
        " [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
          "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
          "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]),
          "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]),
          "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(addr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// Shared-memory matrix descriptor for tcgen05.mma, SWIZZLE_128B, 16-bit elements.
//  K-major operand : rows of 128 B (64 elements of K), 8-row groups 1024 B apart (SBO);
//                    LBO unused for swizzled K-major (encoded 1 like CUTLASS does).
//  MN-major operand: rows of 128 B hold 64 consecutive M/N elements for one k; 8 k-rows
//                    form a 1024 B group (SBO = 1024); the next 64 M/N elements start LBO
//                    bytes further (one TMA box = 64 k-rows * 128 B = 8192 B).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes,
                                                   uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= uint64_t((smem_addr & 0x3FFFF) >> 4);
    d |= uint64_t((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= uint64_t((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= uint64_t(1) << 46;   // descriptor version (Blackwell)
    d |= uint64_t(2) << 61;   // SWIZZLE_128B
    return d;
}
// Instruction descriptor, kind::f16: bf16 x bf16 -> fp32.
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N, bool a_mn, bool b_mn) {
    return (1u << 4) | (1u << 7) | (1u << 10) | (uint32_t(a_mn) << 15) | (uint32_t(b_mn) << 16) |
           (uint32_t(N >> 3) << 17) | (uint32_t(M >> 4) << 24);
}
#endif  // __CUDACC__

}  // namespace vpr
