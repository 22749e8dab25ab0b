// Library-wide state: the error channel, device properties and the C-ABI test hook for the GEMM.
#include "common.cuh"

#include <stdlib.h>
#include "gemm.h"
#include "../../include/vpr_b200.h"

#include <atomic>
#include <mutex>
#include <stdarg.h>
#include <stdio.h>

namespace vpr {

static thread_local char g_err[1024] = "";

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

static std::atomic<unsigned long long> g_launches{0};
void note_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

static std::atomic<int> g_bound_device{-1};

int check_device() {
    int dev = 0;
    VPR_CUDA_OK(cudaGetDevice(&dev));
    int bound = g_bound_device.load(std::memory_order_relaxed);
    if (bound < 0) {
        int expected = -1;
        g_bound_device.compare_exchange_strong(expected, dev);
        bound = g_bound_device.load();
    }
    VPR_REQUIRE(bound == dev,
                "libvpr_b200 is bound to CUDA device %d (one process per GPU) but device %d is current: make the tensors' "
                "device current (torch.cuda.set_device / with torch.cuda.device(...)) or use one process per GPU",
                bound, dev);
    return 0;
}

int num_sms() {
    static int n = 0;
    if (n == 0) {
        int dev = g_bound_device.load();
        if (dev < 0) cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    }
    return n;
}

int ensure_dynamic_smem(const void* kernel, size_t bytes) {
    struct Slot { const void* fn; size_t bytes; };
    static Slot slots[128];
    static int used = 0;
    static std::mutex mu;
    std::lock_guard<std::mutex> lock(mu);
    Slot* sl = nullptr;
    for (int i = 0; i < used; ++i)
        if (slots[i].fn == kernel) { sl = &slots[i]; break; }
    if (sl != nullptr && sl->bytes >= bytes) return 0;
    VPR_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    if (sl == nullptr && used < 128) { sl = &slots[used++]; sl->fn = kernel; sl->bytes = 0; }
    if (sl != nullptr) sl->bytes = bytes;
    return 0;
}

bool pdl_enabled() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("VPR_PDL"); v = (e != nullptr && e[0] == '0') ? 0 : 1; }
    return v == 1;
}

int side_branch(SideBranch** out) {
    static thread_local SideBranch sb = {nullptr, nullptr, nullptr, nullptr, nullptr};
    if (sb.stream == nullptr) {
        VPR_CUDA_OK(cudaStreamCreateWithFlags(&sb.stream, cudaStreamNonBlocking));
        VPR_CUDA_OK(cudaEventCreateWithFlags(&sb.fork, cudaEventDisableTiming));
        VPR_CUDA_OK(cudaEventCreateWithFlags(&sb.join, cudaEventDisableTiming));
        VPR_CUDA_OK(cudaEventCreateWithFlags(&sb.fork2, cudaEventDisableTiming));
        VPR_CUDA_OK(cudaEventCreateWithFlags(&sb.zeroed, cudaEventDisableTiming));
    }
    *out = &sb;
    return 0;
}

}  // namespace vpr

extern "C" {

const char* vpr_last_error(void) { return vpr::g_err; }

int vpr_abi_version(void) { return VPR_ABI_VERSION; }

unsigned long long vpr_launch_count(void) { return vpr::g_launches.load(); }

int vpr_gemm(int dtype, const void* A, int a_mn, int64_t lda, int64_t a_bs, const void* B, int b_mn, int64_t ldb,
             int64_t b_bs, int M, int N, int K, int batch, void* out, int out_kind, int64_t ldo, int64_t out_bs,
             const float* bias, int relu, const void* mask, int64_t ld_mask, int64_t mask_bs, float mask_scale,
             float alpha, int k_splits, int bn, float* colsum, float dropout_p, uint64_t dropout_seed,
             cudaStream_t stream) {
    VPR_TRY(::vpr::check_device());
    vpr::GemmDesc g;
    g.dtype = dtype;
    g.A = A; g.a_mn = a_mn; g.lda = lda; g.a_bs = a_bs;
    g.B = B; g.b_mn = b_mn; g.ldb = ldb; g.b_bs = b_bs;
    g.M = M; g.N = N; g.K = K; g.batch = batch;
    g.out = out; g.out_kind = out_kind; g.ldo = ldo; g.out_bs = out_bs;
    g.bias = bias; g.relu = relu;
    g.mask = mask; g.ld_mask = ld_mask; g.mask_bs = mask_bs; g.mask_scale = mask_scale;
    g.alpha = alpha; g.k_splits = k_splits; g.bn = bn; g.colsum = colsum;
    g.dropout_p = dropout_p; g.dropout_seed = dropout_seed;
    return vpr::gemm_run(g, stream);
}

}  // extern "C"
