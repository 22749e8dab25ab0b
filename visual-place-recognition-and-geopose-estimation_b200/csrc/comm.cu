// Sum all-reduce of fp32 gradient buffers across the GPUs of one NVLink/NVSwitch box, written for the
// fine-tuning step of SURVEY.md section 8e: two messages per step (17 MB of geopose-head gradients that must hide
// behind the SALAD backward, 5.6 MB of SALAD gradients whose tail is exposed), 2..8 ranks, replayed from a CUDA graph.
//
// Every rank holds the same symmetric allocation [ data | staging | flags ] and knows every peer's address of it
// (vpr_b200/parallel.py sets it up with torch's symmetric memory or CUDA IPC).  One launch, two shots:
//   1. push     rank r copies region p of its data into slot r of rank p's staging area   (NVLink stores only)
//   2. reduce   rank p sums the world slots of its region in rank order (its own straight from `data`) and stores the
//               sum into region p of EVERY rank's data                                   (NVLink stores only)
// With a multicast (NVLS) mapping of the allocation the switch does both halves instead: multimem.ld_reduce pulls the
// summed region p, multimem.st broadcasts it.
// All cross-GPU traffic is posted writes or switch-side reductions.  The multimem loads are round trips through the
// switch (~5 us under load), so the kernel lives on memory-level parallelism: 1024 threads per CTA, four independent
// 16-byte accesses in flight per thread.  It runs on SMs of its own: the persistent kernels of the backward it overlaps
// are launched `reserve_sms` CTAs short (vpr_salad_bwd), and a dummy shared-memory request keeps this kernel's CTAs
// off the SMs those kernels occupy.  (Small CTAs that share SMs with the GEMM CTAs were measured first: 40 registers x
// 256 threads do fit beside 168 x 320, but a persistent GEMM with a static tile schedule then waits for the SMs it
// shares -- dW0 79 -> 141 us at N=2 -- which costs far more than giving up 8 of 148 SMs.)
// Synchronisation is per CTA index: CTA c of every rank owns slice c of every region, so a flag from peer q's CTA c
// covers exactly the data CTA c consumes.  Flags carry a call counter kept in device memory (the launch arguments of a
// captured graph never change), never need resetting, and every wait is bounded (trap after ~20 s instead of a hung box).
#include "common.cuh"
#include "../../include/vpr_b200.h"

#include <stdlib.h>

namespace vpr {

static constexpr int COMM_THREADS = 1024;
static constexpr int COMM_DUMMY_SMEM = 64 * 1024;      // > 227 KB minus the big kernels' footprint: no SM sharing with them
static constexpr int COMM_MAX_WORLD = VPR_COMM_MAX_WORLD;
static constexpr int COMM_MAX_CTAS = VPR_COMM_MAX_CTAS;
// flags area of one rank, in 32-bit words: arrive[2 phases][world][ctas] then epoch[ctas]
__host__ __device__ constexpr size_t comm_flag_words() { return (size_t)2 * COMM_MAX_WORLD * COMM_MAX_CTAS + COMM_MAX_CTAS; }

struct CommArgs {
    float* data[COMM_MAX_WORLD];        // peer addresses of the data area (data[rank] is local)
    float* staging[COMM_MAX_WORLD];     // peer addresses of the staging area
    uint32_t* flags[COMM_MAX_WORLD];    // peer addresses of the flags area
    float* mc_data;                     // multicast address of the data area, or null
    long long off4, count4, chunk4;     // message = float4 [off4, off4 + count4) of data; region p = chunk4 float4s
    int rank, world;
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float4 ld_f4_sys(const float4* p) {      // data another GPU wrote: never from L1
    float4 v;
    asm volatile("ld.relaxed.sys.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_f4_sys(float4* p, const float4 v) {
    asm volatile("st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float4 multimem_ld_reduce_f4(const float4* mc) {
    float4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(mc) : "memory");
    return v;
}
__device__ __forceinline__ void multimem_st_f4(float4* mc, const float4 v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}

// All of this CTA's earlier writes (any thread) are published, then flag word `slot` of every peer is set to `epoch`;
// then wait until every peer has set ours.  Threads 0..world-1 each talk to one peer.
__device__ __forceinline__ void comm_barrier(const CommArgs& a, int phase, int cta, uint32_t epoch) {
    __syncthreads();
    const int t = threadIdx.x;
    if (t < a.world && t != a.rank) {
        // st.release.sys orders this thread's earlier writes AND (cumulativity through the bar.sync above) the rest of the
        // CTA's before the flag; a separate fence.sc.sys in front of it would only add a second NVLink round trip
        const size_t mine = ((size_t)phase * COMM_MAX_WORLD + a.rank) * COMM_MAX_CTAS + cta;      // written by me, at peer t
        st_release_sys(a.flags[t] + mine, epoch);
        const size_t theirs = ((size_t)phase * COMM_MAX_WORLD + t) * COMM_MAX_CTAS + cta;         // written by t, here
        const uint32_t* f = a.flags[a.rank] + theirs;
        long long t0 = 0;
        int spins = 0;
        while ((int)(ld_acquire_sys(f) - epoch) < 0) {
            if (++spins == 256) t0 = clock64();
            if (spins > 256) {
                __nanosleep(64);
                if ((spins & 1023) == 0 && clock64() - t0 > 40000000000LL) __trap();     // ~20 s: a peer is gone
            }
        }
    }
    __syncthreads();
}

template <bool MULTIMEM>
__global__ void __launch_bounds__(COMM_THREADS, 2) allreduce_kernel(const CommArgs a) {
    const int cta = blockIdx.x, nct = gridDim.x, tid = threadIdx.x;
    const int NT = blockDim.x;                      // <= COMM_THREADS (tuning knob VPR_COMM_THREADS)
    __shared__ uint32_t epoch_s;
    uint32_t* my_flags = a.flags[a.rank];
    uint32_t* epoch_p = my_flags + (size_t)2 * COMM_MAX_WORLD * COMM_MAX_CTAS + cta;
    if (tid == 0) epoch_s = *epoch_p + 1u;
    __syncthreads();
    const uint32_t epoch = epoch_s;
    // slice of a region this CTA owns: float4 [s0, s1) relative to the region start (the same on every rank); all
    // loop indices are 32-bit offsets from per-region base pointers (the kernel is capped at 32 registers so that its
    // CTAs fit beside a half-occupied SM of the kernels it overlaps)
    const int chunk = (int)a.chunk4;
    const int per = (chunk + nct - 1) / nct;
    const int s0 = per * cta;
    auto slice_end = [&](int region) {              // exclusive end of this CTA's slice inside `region`, clipped to the message
        long long lim = a.count4 - (long long)chunk * region;
        int e = s0 + per < chunk ? s0 + per : chunk;
        if (lim < e) e = lim < 0 ? 0 : (int)lim;
        return e;
    };
    float4* mine = reinterpret_cast<float4*>(a.data[a.rank]) + a.off4;
    if (MULTIMEM) {
        comm_barrier(a, 0, cta, epoch);                                   // every rank's data is final
        float4* mc = reinterpret_cast<float4*>(a.mc_data) + a.off4 + (long long)chunk * a.rank;
        const int e1 = slice_end(a.rank);
        int i = s0 + tid;
        for (; i + 3 * NT < e1; i += 4 * NT) {                            // four switch round trips in flight per thread
            const float4 v0 = multimem_ld_reduce_f4(mc + i), v1 = multimem_ld_reduce_f4(mc + i + NT),
                         v2 = multimem_ld_reduce_f4(mc + i + 2 * NT), v3 = multimem_ld_reduce_f4(mc + i + 3 * NT);
            multimem_st_f4(mc + i, v0);
            multimem_st_f4(mc + i + NT, v1);
            multimem_st_f4(mc + i + 2 * NT, v2);
            multimem_st_f4(mc + i + 3 * NT, v3);
        }
        for (; i < e1; i += NT) multimem_st_f4(mc + i, multimem_ld_reduce_f4(mc + i));
        comm_barrier(a, 1, cta, epoch);                                   // every region has landed here
    } else {
        // ---- shot 1: my slice of region p -> slot `rank` of p's staging ----
        for (int k = 1; k < a.world; ++k) {
            const int p = (a.rank + k) % a.world;                         // staggered: no two ranks start on the same peer
            const float4* src = mine + (long long)chunk * p;
            float4* dst = reinterpret_cast<float4*>(a.staging[p]) + (long long)chunk * a.rank;
            const int e1 = slice_end(p);
            int i = s0 + tid;
            for (; i + 3 * NT < e1; i += 4 * NT) {                        // four independent 16-byte copies in flight
                const float4 v0 = src[i], v1 = src[i + NT], v2 = src[i + 2 * NT], v3 = src[i + 3 * NT];
                st_f4_sys(dst + i, v0);
                st_f4_sys(dst + i + NT, v1);
                st_f4_sys(dst + i + 2 * NT, v2);
                st_f4_sys(dst + i + 3 * NT, v3);
            }
            for (; i < e1; i += NT) st_f4_sys(dst + i, src[i]);
        }
        comm_barrier(a, 0, cta, epoch);                                   // all contributions to my region are here
        // ---- shot 2: sum my region in rank order, store it everywhere ----
        const float4* own = mine + (long long)chunk * a.rank;
        const float4* stg = reinterpret_cast<const float4*>(a.staging[a.rank]);
        const long long out_off = a.off4 + (long long)chunk * a.rank;
        const int e1 = slice_end(a.rank);
        for (int i = s0 + tid; i < e1; i += NT) {
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int q = 0; q < a.world; ++q) {
                const float4 v = q == a.rank ? own[i] : ld_f4_sys(stg + (long long)chunk * q + i);
                acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
            }
            for (int k = 0; k < a.world; ++k) {
                const int p = (a.rank + k) % a.world;
                st_f4_sys(reinterpret_cast<float4*>(a.data[p]) + out_off + i, acc);
            }
        }
        comm_barrier(a, 1, cta, epoch);                                   // every region has landed here
    }
    if (tid == 0) *epoch_p = epoch;
}

}  // namespace vpr

using namespace vpr;

extern "C" {

size_t vpr_comm_flag_bytes(void) { return comm_flag_words() * sizeof(uint32_t); }

size_t vpr_comm_staging_elems(size_t max_count, int world) {
    if (world < 1) return 0;
    const size_t c4 = (max_count + 3) / 4;
    const size_t chunk4 = (c4 + world - 1) / world;
    return chunk4 * 4 * (size_t)world;
}

int vpr_allreduce_sum_f32(const VprCommPeers* peers, size_t offset_elems, size_t count, int ctas, cudaStream_t stream) {
    VPR_REQUIRE(peers != nullptr, "null argument");
    VPR_TRY(check_device());
    VPR_REQUIRE(peers->world >= 1 && peers->world <= COMM_MAX_WORLD && peers->rank >= 0 && peers->rank < peers->world,
                "bad rank/world (%d/%d; at most %d ranks)", peers->rank, peers->world, COMM_MAX_WORLD);
    VPR_REQUIRE(offset_elems % 4 == 0 && count % 4 == 0, "offset and count must be multiples of 4 floats");
    if (count == 0 || peers->world == 1) return 0;
    if (ctas < 1) ctas = 16;
    if (ctas > COMM_MAX_CTAS) ctas = COMM_MAX_CTAS;
    CommArgs a;
    for (int r = 0; r < COMM_MAX_WORLD; ++r) {
        const bool in = r < peers->world;
        a.data[r] = in ? reinterpret_cast<float*>(peers->data[r]) : nullptr;
        a.staging[r] = in ? reinterpret_cast<float*>(peers->staging[r]) : nullptr;
        a.flags[r] = in ? reinterpret_cast<uint32_t*>(peers->flags[r]) : nullptr;
        VPR_REQUIRE(!in || (a.data[r] && a.staging[r] && a.flags[r]), "null peer pointer for rank %d", r);
        VPR_REQUIRE(!in || ((reinterpret_cast<uintptr_t>(a.data[r]) | reinterpret_cast<uintptr_t>(a.staging[r])) & 15) == 0,
                    "peer buffers must be 16-byte aligned");
    }
    a.mc_data = reinterpret_cast<float*>(peers->mc_data);
    a.rank = peers->rank; a.world = peers->world;
    a.off4 = (long long)(offset_elems / 4);
    a.count4 = (long long)(count / 4);
    a.chunk4 = (a.count4 + a.world - 1) / a.world;
    VPR_REQUIRE((size_t)a.chunk4 * 4 * a.world <= peers->staging_elems, "staging area too small: %zu < %zu floats",
                peers->staging_elems, (size_t)a.chunk4 * 4 * a.world);
    if ((long long)ctas > a.chunk4) ctas = (int)a.chunk4;
    static int threads = 0;
    if (threads == 0) {
        const char* e = getenv("VPR_COMM_THREADS");
        threads = e ? atoi(e) : COMM_THREADS;
        if (threads < 32 || threads > COMM_THREADS || threads % 32) threads = COMM_THREADS;
    }
    if (a.mc_data != nullptr) {
        VPR_SMEM_OK(allreduce_kernel<true>, COMM_DUMMY_SMEM);
        allreduce_kernel<true><<<ctas, threads, COMM_DUMMY_SMEM, stream>>>(a);
    } else {
        VPR_SMEM_OK(allreduce_kernel<false>, COMM_DUMMY_SMEM);
        allreduce_kernel<false><<<ctas, threads, COMM_DUMMY_SMEM, stream>>>(a);
    }
    return launch_ok("allreduce_kernel");
}

}  // extern "C"
