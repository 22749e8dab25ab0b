// One GEMM entry point for the whole library:
//   out[b][m][n] (op)= epilogue( alpha * sum_k A(b,m,k) * B(b,n,k) )
// dtype 1 (bf16 operands, fp32 accumulate) runs the persistent warp-specialised
// TMA + tcgen05 kernel; dtype 0 (fp32 operands) runs a plain fp32 FMA kernel that
// exists for the north-star's 1e-4 fp32 parity mode (tensor cores cannot give fp32
// products).  Both share the epilogue semantics below.
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>

namespace vpr {

enum { OUT_F32 = 0, OUT_BF16 = 1, OUT_ATOMIC_F32 = 2 };
enum { DT_F32 = 0, DT_BF16 = 1 };

struct GemmDesc {
    int dtype = DT_BF16;
    // A(m,k): a_mn == 0 -> stored [M rows][K inner] ("K-major"), lda = row stride
    //         a_mn == 1 -> stored [K rows][M inner] ("MN-major"), lda = row stride
    const void* A = nullptr; int a_mn = 0; int64_t lda = 0; int64_t a_bs = 0;
    // B(n,k): same convention with N in place of M.  b_bs == 0 -> shared by all batches.
    const void* B = nullptr; int b_mn = 0; int64_t ldb = 0; int64_t b_bs = 0;
    int M = 0, N = 0, K = 0, batch = 1;
    void* out = nullptr; int out_kind = OUT_F32; int64_t ldo = 0; int64_t out_bs = 0;
    const float* bias = nullptr;   // per-n, added before relu / mask
    int relu = 0;
    // v = mask(b,m,n) > 0 ? v * mask_scale : 0 ; mask has the operand dtype
    const void* mask = nullptr; int64_t ld_mask = 0; int64_t mask_bs = 0; float mask_scale = 1.f;
    float alpha = 1.f;
    // inverted dropout on the output (after relu): keep iff dropout_hash(seed, (b*M+m)*N+n) >= p*2^32
    float dropout_p = 0.f; uint64_t dropout_seed = 0;
    const uint64_t* dropout_seed_dev = nullptr;   // optional device counter added to dropout_seed (CUDA-graph replays)
    // colsum[n] += sum over (b, m) of the stored output (after mask / rounding); caller zeroes it
    float* colsum = nullptr;
    int k_splits = 1;              // > 1 requires OUT_ATOMIC_F32 (caller zeroes out)
    int bn = 0;                    // tcgen05 tile N (64/128/256); 0 = auto
    int cta_pair = 0;              // 0 = auto (256 x 256 CTA-pair tiles when N > 128), 2 = single-CTA tiles
    int reserve_sms = 0;           // leave this many SMs without a persistent CTA (room for a concurrent collective)
    int pdl = 0;                   // 1: the previous launch in the stream is one of ours: programmatic dependent launch;
                                   // 2: additionally none of this problem's operands is written by that launch
};

// How a group will be tiled: tile shape and the number of tiles that run concurrently (SMs, or SM
// pairs for 256 x 256 CTA-pair tiles).  Callers size split-K with it.
struct GemmTilePlan { int tile_m, tile_n, slots; };
GemmTilePlan gemm_tile_plan(const GemmDesc* gs, int count);

int gemm_run(const GemmDesc& g, cudaStream_t stream);

// SALAD backward, fused: dH = [dF W1c | dS W1s] * relu/dropout mask, and -- from the tiles that launch already has
// in shared memory -- the second-layer weight gradients dW1 = [dF^T H_c ; dS^T H_s] and the first-layer bias
// gradients (column sums of dH), all on the tensor core.  dtype bf16, l == 128, m == 64, hidden % 128 == 0.
struct DhFusedDesc {
    const void* dF; const void* dS;          // [T, 128], [T, 64] bf16
    const void* W1c; const void* W1s;        // [128, hidden], [64, hidden] bf16
    const void* H;                           // [T, 2*hidden] bf16 saved activation (mask and dW1 operand)
    void* dH;                                // [T, 2*hidden] bf16
    float* dW1c; float* dW1s;                // [128, hidden], [64, hidden] fp32, accumulated (caller zeroes)
    float* db0c; float* db0s;                // [hidden] fp32 each, accumulated (caller zeroes)
    float* db1c;                             // [128] fp32 column sums of dF (cluster_features.3.bias), or NULL
    long long T; int hidden; float mask_scale;
    int pdl = 0;
    int reserve_sms = 0;                     // leave this many SMs without a CTA (room for a concurrent collective)
};
// SALAD forward, second layer: F = H[:, :h] W1c^T + b1c (bf16), S = H[:, h:] W1s^T + b1s (fp32), weights resident
// in shared memory, H streamed once.  dtype bf16, l == 128, m == 64, hidden == 512.
struct FsFwdDesc {
    const void* H; const void* W1c; const void* W1s;     // [T, 2*hidden], [128, hidden], [64, hidden] bf16
    const float* bias_f; const float* bias_s;            // [128], [64]
    void* F; float* S;                                   // [T, 128] bf16, [T, 64] fp32
    long long T; int hidden;
    // token MLP second layer on two extra CTAs: TF[B, g] = HT[B, hidden] Wt1[g, hidden]^T + bias_t (g == 256)
    const void* HT; const void* Wt1; const float* bias_t; float* TF; int B; int g;
    int pdl = 0;
};
bool fs_fwd_supported(int dtype, int l, int m, int hidden);
int fs_fwd_run(const FsFwdDesc& d, cudaStream_t stream);

// SALAD forward, last step, one CTA per image: AGG = F^T P on the tensor core fused with the descriptor assembly
// (per-cluster L2 norm over l, token L2 norm, concat, global L2 norm).  dtype bf16, l == 128, m == 64.
struct AggFwdDesc {
    const void* F; const void* P;            // [B, n, 128], [B, n, 64] bf16
    const float* TF;                         // [B, g] fp32
    float* AGG;                              // [B, 128, 64] fp32 (kept for the backward)
    void* out;                               // [B, g + 8192] bf16
    int B, n, g;
    int pdl = 0;
};
bool agg_fwd_supported(int dtype, int l, int m, int g, int n);
int agg_fwd_run(const AggFwdDesc& d, cudaStream_t stream);

// SALAD forward, Sinkhorn + aggregation + descriptor assembly in one launch, one CTA per image (sink_agg_fwd.cu):
// n <= 256 tokens, l == 128, m == 64, bf16.  The plan goes from the solver's registers into the MMA operand in shared memory.
struct SinkAggFwdDesc {
    const float* S; const float* dust_bin;   // [B, n, 64] fp32 scores, device scalar
    const void* F;                           // [B, n, 128] bf16
    void* P; float* U; float* V;             // [B, n, 64] bf16 plan and [B, iters, 65] / [B, iters, n] potentials (kept for the backward)
    const float* TF;                         // [B, g] fp32
    float* AGG;                              // [B, 128, 64] fp32 (kept for the backward)
    void* out;                               // [B, g + 8192] bf16
    int B, n, g, iters;
    int pdl = 0;
};
bool sink_agg_fwd_supported(int dtype, int l, int m, int g, int n, int iters);
int sink_agg_fwd_run(const SinkAggFwdDesc& d, cudaStream_t stream);

// SALAD backward, first step, one CTA per image: cluster part of the descriptor-assembly backward fused with
// dF = P dAGG^T and dP = F dAGG (dAGG never leaves shared memory).  dtype bf16, l == 128, m == 64.
struct AggBwdDesc {
    const float* AGG; const float* TF; const void* d_out;    // [B,128,64] fp32, [B,g] fp32, [B, g+8192] bf16
    const void* P; const void* F;                            // [B, n, 64], [B, n, 128] bf16
    void* dF; float* dP;                                     // [B, n, 128] bf16, [B, n, 64] fp32
    void* dTF; float* g_bias_tok;                            // [B, g] bf16; [g] += column sums of dTF (or null)
    int B, n, g;
    int pdl = 0;
};
bool agg_bwd_supported(int dtype, int l, int m, int g);
int agg_bwd_run(const AggBwdDesc& d, cudaStream_t stream);

bool dh_fused_supported(int dtype, int l, int m, int hidden);
int dh_fused_run(const DhFusedDesc& d, cudaStream_t stream);
// Up to 4 independent problems with the same dtype and operand layout in ONE launch.
int gemm_run_group(const GemmDesc* gs, int count, cudaStream_t stream);

}  // namespace vpr
