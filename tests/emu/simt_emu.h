// Host-side SIMT emulation -- TEST INFRASTRUCTURE ONLY (nothing under the package includes this).
//
// Lets a CUDA kernel written against a small subset of the device API (threadIdx / blockIdx, dynamic shared
// memory, __syncthreads, full-mask __shfl_xor_sync, __ldg, atomicAdd(float), float4 / uint4 vectors, bf16 packing)
// compile as plain C++ and run on the CPU: one OS thread per CUDA thread, one thread block at a time, pthread
// barriers for the block and for each warp.  The CPU test suite uses it to run the device source of the
// register-tile Sinkhorn kernels (csrc/sinkhorn_rt.cuh) against the fp64 oracle without a GPU.
#pragma once
#define VPR_SIMT_EMU 1
#include <pthread.h>

#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)

struct float4 { float x, y, z, w; };
struct uint2 { uint32_t x, y; };
struct uint4 { uint32_t x, y, z, w; };
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
struct __nv_bfloat16 { uint16_t v; };
struct __nv_bfloat162 { __nv_bfloat16 x, y; };
static inline uint16_t simt_f2bf(float f) {       // round to nearest even, like cvt.rn.bf16.f32
    uint32_t u;
    std::memcpy(&u, &f, 4);
    if ((u & 0x7fffffffu) > 0x7f800000u) return 0x7fc0;
    u += 0x7fffu + ((u >> 16) & 1u);
    return (uint16_t)(u >> 16);
}
static inline __nv_bfloat162 __floats2bfloat162_rn(float a, float b) {
    __nv_bfloat162 r;
    r.x.v = simt_f2bf(a);
    r.y.v = simt_f2bf(b);
    return r;
}

namespace simt_emu {
struct Idx { unsigned x, y, z; };
struct Warp {
    pthread_barrier_t bar;
    float buf[32];
};
struct Block {
    pthread_barrier_t bar;
    std::vector<Warp> warps;
    float* smem;
};
struct Tls {
    Block* block;
    Warp* warp;
    int lane;
};
inline thread_local Tls tls;
inline std::atomic<int> guard_count{0};
inline std::mutex atomic_mutex;
inline void note_guard() {
    if (tls.lane == 0 && tls.warp == &tls.block->warps[0]) guard_count.fetch_add(1);
}
inline float* dyn_smem() { return tls.block->smem; }

template <typename F>
void launch(int grid, int nthreads, size_t smem_bytes, F kernel_body);
}  // namespace simt_emu

inline thread_local simt_emu::Idx threadIdx, blockIdx;

static inline void __syncthreads() { pthread_barrier_wait(&simt_emu::tls.block->bar); }
static inline float __shfl_xor_sync(unsigned, float v, int o) {
    simt_emu::Warp& w = *simt_emu::tls.warp;
    const int l = simt_emu::tls.lane;
    w.buf[l] = v;
    pthread_barrier_wait(&w.bar);
    const float r = w.buf[l ^ o];
    pthread_barrier_wait(&w.bar);
    return r;
}
template <typename T>
static inline T __ldg(const T* p) { return *p; }
static inline float atomicAdd(float* p, float v) {
    std::lock_guard<std::mutex> g(simt_emu::atomic_mutex);
    const float old = *p;
    *p = old + v;
    return old;
}

template <typename F>
void simt_emu::launch(int grid, int nthreads, size_t smem_bytes, F kernel_body) {
    const int nwarps = (nthreads + 31) / 32;
    for (int b = 0; b < grid; ++b) {
        Block blk;
        pthread_barrier_init(&blk.bar, nullptr, nthreads);
        blk.warps = std::vector<Warp>(nwarps);
        for (int w = 0; w < nwarps; ++w) {
            const int lanes = (w == nwarps - 1 && nthreads % 32) ? nthreads % 32 : 32;
            pthread_barrier_init(&blk.warps[w].bar, nullptr, lanes);
        }
        // poison shared memory: a read of something never written shows up as NaN in the results
        blk.smem = static_cast<float*>(std::aligned_alloc(128, (smem_bytes + 127) / 128 * 128));
        std::memset(blk.smem, 0xff, smem_bytes);
        std::vector<std::thread> threads;
        threads.reserve(nthreads);
        for (int t = 0; t < nthreads; ++t) {
            threads.emplace_back([&, t, b]() {
                tls.block = &blk;
                tls.warp = &blk.warps[t / 32];
                tls.lane = t % 32;
                threadIdx = Idx{(unsigned)t, 0, 0};
                blockIdx = Idx{(unsigned)b, 0, 0};
                kernel_body();
            });
        }
        for (auto& th : threads) th.join();
        for (int w = 0; w < nwarps; ++w) pthread_barrier_destroy(&blk.warps[w].bar);
        pthread_barrier_destroy(&blk.bar);
        std::free(blk.smem);
    }
}

// ---- the device helpers csrc/sinkhorn_rt.cuh takes from the CUDA headers of the package ----
#define VPR_L2E 1.4426950408889634f
#define VPR_LN2 0.6931471805599453f
#define VPR_RT_SMEM(name) float* name = ::simt_emu::dyn_smem()
namespace vpr {
struct SinkConst2 {
    float norm2;
    float log_a_m2;
};
static inline float fex2(float x) { return exp2f(x); }
static inline float flg2(float x) { return log2f(x); }
static inline float rt_rcp(float x) { return 1.0f / x; }
static inline void pdl_wait() {}
static inline void pdl_launch_dependents() {}
}  // namespace vpr
