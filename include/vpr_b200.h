/* C-ABI of libvpr_b200.so: the B200 (sm_100a) implementation of the one hot path of
 * anshium/visual-place-recognition-and-geopose-estimation -- SALAD aggregation head
 * (forward + backward), geopose heads + losses, descriptor kNN.
 *
 * Conventions
 *  - Every pointer is a DEVICE pointer owned by the caller (PyTorch); the library allocates no
 *    device memory.
 *  - State the library does keep (none of it changes results): per-process, the largest
 *    MaxDynamicSharedMemorySize set per kernel and the SM count of the device; per calling
 *    thread, a cache of 512 encoded TMA tensor maps keyed by (pointer, shape, strides) and one
 *    non-blocking side stream with five events on which vpr_salad_bwd forks its small
 *    token-MLP launches (created on first use, joined back into `stream` before the call
 *    returns; under stream capture they become a parallel branch of the caller's graph).
 *  - One process drives ONE device: the first call binds the library to the current device and
 *    every later call fails with a clear message if another device is current.
 *  - Every call is asynchronous on `stream`, re-entrant and does not synchronise.
 *  - Return value 0 = success; otherwise vpr_last_error() (thread-local) explains.
 *  - dtype: VPR_F32 = fp32 storage + fp32 FMA GEMMs (1e-4 parity mode),
 *           VPR_BF16 = bf16 storage, tcgen05 bf16 GEMMs with fp32 accumulation;
 *           Sinkhorn, norms and losses always compute in fp32.
 *
 * Reference interfaces replaced (paths under the reference repo):
 *  - vpr_salad_fwd            : the aggregator call inside `self.feature_extractor(x)`,
 *                               dinov2salad/dinov2salad_finetuning.py:36 and
 *                               dinov2salad/dinov2salad_validation.py:81, i.e. upstream
 *                               serizba/salad SALAD.forward reached through
 *                               torch.hub.load (dinov2salad_finetuning.py:19,92).
 *  - vpr_salad_bwd            : autograd through that call (the reference freezes it,
 *                               dinov2salad_finetuning.py:25-26,35; north-star adds backward).
 *  - vpr_sinkhorn_fwd / _bwd  : upstream get_matching_probs / log_otp_solver.
 *  - vpr_heads_loss_fwd_bwd   : `self.regressor(features)` + `loss_fn(preds, targets)` +
 *                               `loss.backward()`, dinov2salad_finetuning.py:28-32,37,120-124;
 *                               Huber variant dinov2salad_finetuning_2.py:118-123,154; sin/cos
 *                               angle head + losses angle_prediction/dinov2salad/dino_v2_gemini.py:
 *                               106-141,243-249 and angle_prediction/swin/
 *                               swin_angle_finetuning_sin_cos.py:56-69.
 *  - vpr_knn_topk             : no reference call site (SURVEY.md section 0 finding 4);
 *                               replaces torch.topk(Q @ G.T, k).
 *  Rows next to the path, each documented at its declaration below:
 *  - vpr_seam_ln_split_fwd/_bwd : the backbone's final LayerNorm + CLS / patch-token split ([UPSTREAM] DINOv2
 *                               forward_features, reached from dinov2salad_finetuning.py:36).
 *  - vpr_metrics_latlon / _angle, vpr_recall_at_k : dinov2salad_validation.py:84-101,
 *                               dinov2salad_finetuning_2.py:202-215, angle_prediction/efficient_net/
 *                               validation_script.py:180-202, dino_v2_gemini.py:117-141.
 *  - vpr_preprocess_image / _batch : transforms.Resize + ToTensor + Normalize, dinov2salad_finetuning.py:45-50.
 *  Environment knobs read once per process (A/B measurements): VPR_PDL, VPR_GEMM_2CTA, VPR_GEMM_MCB, VPR_FS_FUSED,
 *  VPR_FS_REVERSE, VPR_DH_FUSED, VPR_AGG_FUSED, VPR_AGG_BWD_FUSED, VPR_TOKEN_SIDE, VPR_KNN_L2_MB.  Read at every call:
 *  VPR_SINK_RT (0 = n <= 256 stays on the shared-memory-tile Sinkhorn kernels instead of the register-tile ones, 2 = their
 *  one-thread-per-row variant), VPR_SINKAGG_FUSED (1 = Sinkhorn + aggregation + descriptor assembly in one launch).
 */
#ifndef VPR_B200_H
#define VPR_B200_H

#include <stddef.h>
#include <stdint.h>
#include <cuda_runtime.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VPR_ABI_VERSION 5
#if defined(__GNUC__)
#define VPR_API __attribute__((visibility("default")))
#else
#define VPR_API
#endif
#define VPR_F32 0
#define VPR_BF16 1

typedef struct {
    int B;        /* images in the batch */
    int n;        /* patch tokens per image (h*w); n > m */
    int d;        /* num_channels (768 for ViT-B/14) */
    int hidden;   /* hidden width of the three MLPs (512 upstream) */
    int m;        /* num_clusters (64) */
    int l;        /* cluster_dim (128) */
    int g;        /* token_dim (256) */
    int iters;    /* Sinkhorn iterations (3 upstream) */
} VprSaladShape;

/* fp32 master parameters in upstream state_dict layout (1x1 conv weights are [out,in]). */
typedef struct {
    const float* cluster_w0; const float* cluster_b0;   /* [hidden,d], [hidden] */
    const float* cluster_w1; const float* cluster_b1;   /* [l,hidden], [l] */
    const float* score_w0;   const float* score_b0;     /* [hidden,d], [hidden] */
    const float* score_w1;   const float* score_b1;     /* [m,hidden], [m] */
    const float* token_w0;   const float* token_b0;     /* [hidden,d], [hidden] */
    const float* token_w1;   const float* token_b1;     /* [g,hidden], [g] */
    const float* dust_bin;                              /* scalar */
} VprSaladParams;

/* fp32 gradients, same layout; every buffer is OVERWRITTEN by vpr_salad_bwd. */
typedef struct {
    float* cluster_w0; float* cluster_b0;
    float* cluster_w1; float* cluster_b1;
    float* score_w0;   float* score_b0;
    float* score_w1;   float* score_b1;
    float* token_w0;   float* token_b0;
    float* token_w1;   float* token_b1;
    float* dust_bin;
} VprSaladGrads;

VPR_API const char* vpr_last_error(void);
VPR_API int vpr_abi_version(void);
/* Number of kernels this library has launched in this process (bench.py's gpu_launches). */
VPR_API unsigned long long vpr_launch_count(void);

/* Bytes of the activation buffer forward fills and backward reads / of the scratch
 * buffer either call may clobber. */
VPR_API size_t vpr_salad_save_bytes(const VprSaladShape* s, int dtype);
VPR_API size_t vpr_salad_workspace_bytes(const VprSaladShape* s, int dtype);

/* x: [B, n, d] token-major contiguous (the memory of upstream's permuted [B,d,h,w] view),
 * t: [B, d], out: [B, g + l*m] L2-normalised rows; all of `dtype`.
 * dropout_p > 0 applies the two Dropout(p) layers of cluster_features / score with a
 * counter-based mask keyed on `seed` + *seed_offset_dev (see vpr_dropout_mask). */
VPR_API int vpr_salad_fwd(const VprSaladParams* p, const void* x, const void* t, void* out, void* save,
                  void* workspace, size_t workspace_bytes, const VprSaladShape* s, int dtype,
                  float dropout_p, uint64_t seed, const uint64_t* seed_offset_dev, cudaStream_t stream);

/* *counter += 1 on the stream.  seed_offset_dev (a device counter, may be NULL) is added to `seed` by the
 * kernels themselves, so a forward captured in a CUDA graph draws a fresh dropout mask on every replay:
 * capture a vpr_counter_advance after the forward. */
VPR_API int vpr_counter_advance(uint64_t* counter, cudaStream_t stream);

/* d_out: [B, g + l*m]; dx: [B, n, d] or NULL; dt: [B, d] or NULL (all `dtype`). */
VPR_API int vpr_salad_bwd(const VprSaladParams* p, const void* x, const void* t, const void* d_out,
                  const void* save, void* workspace, size_t workspace_bytes, VprSaladGrads* grads,
                  void* dx, void* dt, const VprSaladShape* s, int dtype, float dropout_p,
                  int phase, int reserve_sms, cudaStream_t stream);
/* phase: VPR_BWD_ALL, or the same work in two calls -- VPR_BWD_PARAMS (every parameter gradient) then
 * VPR_BWD_INPUTS (dx, dt) -- so that a data-parallel caller can start the gradient all-reduce between them;
 * reserve_sms > 0 keeps that many SMs free of the two many-tile persistent kernels of the phase(s) this call runs
 * (fused dH pass in VPR_BWD_PARAMS, input-gradient GEMM in VPR_BWD_INPUTS) for a collective that runs beside them;
 * the first-layer weight-gradient GEMM is a single wave sized to the whole machine and never gives SMs up. */
#define VPR_BWD_ALL 0
#define VPR_BWD_PARAMS 1
#define VPR_BWD_INPUTS 2

/* The keep-mask ({0,1} bytes, [rows, cols]) the dropout epilogue uses for (seed). */
VPR_API int vpr_dropout_mask(uint64_t seed, float p, int64_t rows, int64_t cols, uint8_t* out, cudaStream_t stream);

/* Stand-alone Sinkhorn (upstream get_matching_probs + exp + drop dustbin row).
 * S: [B, n, m] fp32 token-major scores; P: [B, n, m] (`p_dtype`); U: [B, iters, m+1] fp32,
 * V: [B, iters, n] fp32 potentials kept for the backward.  dust_bin: device scalar. */
VPR_API int vpr_sinkhorn_fwd(const float* S, const float* dust_bin, void* P, int p_dtype, float* U, float* V,
                     int B, int n, int m, int iters, cudaStream_t stream);
/* dP: [B, n, m] fp32 -> dS: [B, n, m] (`ds_dtype`); d_bias (nullable): [m] fp32 += sum_{b,j} dS. */
VPR_API int vpr_sinkhorn_bwd(const float* S, const float* dust_bin, const float* U, const float* V,
                     const float* dP, void* dS, int ds_dtype, float* d_bias, int B, int n, int m,
                     int iters, cudaStream_t stream);

/* ---- geopose heads + losses, forward and backward in one call --------------------------- */
#define VPR_LATLON_MSE 0
#define VPR_LATLON_HUBER 1
#define VPR_LATLON_EXTERNAL 2    /* `latlon_target` holds dL/dpred [B,2] computed by the caller (autograd through
                                    `preds = model(inputs); loss = loss_fn(preds, targets); loss.backward()`,
                                    dinov2salad_finetuning.py:120-124); reported lat/lon loss is 0 */
#define VPR_ANGLE_NONE 0
#define VPR_ANGLE_ATAN2_MAAE 1   /* dino_v2_gemini.py:117-141 */
#define VPR_ANGLE_ACOS 2         /* swin_angle_finetuning_sin_cos.py:56-69 */
#define VPR_ANGLE_MSE 3          /* swin_angle_finetuning_gemini.py:183 on [sin,cos] */
#define VPR_ANGLE_SCALAR_MAAE 4  /* dino_angle_finetuning.py:37-52: head = Linear(D,1), pred = out % 360,
                                    loss = mean(min(d, 360-d)), d = |pred - target| % 360.  ang_w is [1, D], ang_b [1];
                                    sincos_pred[b] = (pred in [0,360), raw head output) */
#define VPR_HEADS_FORWARD_ONLY 1 /* flags: predictions and losses only; `grads` / `d_desc` may be NULL */
#define VPR_HEADS_PREPACKED 2    /* flags: vpr_heads_pack already filled this workspace from the current weights */
#define VPR_HEADS_SKIP_WGRAD 4   /* flags: everything except the two launches nothing downstream of d_desc waits for
                                    (the regressor.0.weight gradient GEMM and the angle head's weight gradient) ... */
#define VPR_HEADS_PREZEROED 16   /* flags: vpr_heads_zero already cleared this call's accumulators (workspace, losses,
                                    gradient buffers), e.g. on a side stream during the SALAD forward */
#define VPR_HEADS_WGRAD_ONLY 8   /* ... which a second call with the same arguments and workspace then issues, e.g. on
                                    another stream after an event, so that they run beside the SALAD backward */

typedef struct {
    int B;          /* rows in this call */
    int D;          /* descriptor width (8448) */
    int hidden;     /* regressor hidden width (512) */
    int latlon_loss;
    int angle_loss;
    float huber_delta;
    float loss_scale;   /* multiplies every gradient (not the reported losses) */
    int global_B;       /* denominator of the batch means (== B on one GPU; sum of B over ranks under DP) */
    float reg_dropout_p;   /* Dropout between ReLU and Linear(hidden,2): dinov2salad_finetuning_2.py:121 */
    uint64_t seed;
    float head_dropout_p;  /* Dropout on the angle head's input: dino_v2_gemini.py:104,111 (mask = vpr_dropout_mask
                              with seed + 0x9E3779B97F4A7C15 over [B, D]) */
    int flags;             /* VPR_HEADS_* */
    const uint64_t* seed_offset_dev;   /* optional device counter added to `seed` (CUDA-graph replays draw fresh
                                          masks; advance it with vpr_counter_advance) */
} VprHeadsShape;

typedef struct {
    const float* reg_w0; const float* reg_b0;   /* [hidden, D], [hidden]  regressor.0 */
    const float* reg_w1; const float* reg_b1;   /* [2, hidden], [2]       regressor.2 */
    const float* ang_w;  const float* ang_b;    /* [2, D], [2]            head        */
} VprHeadsParams;

typedef struct {
    float* reg_w0; float* reg_b0; float* reg_w1; float* reg_b1; float* ang_w; float* ang_b;
} VprHeadsGrads;

VPR_API size_t vpr_heads_workspace_bytes(const VprHeadsShape* s, int dtype);
/* Cast the big weight matrices to the run dtype ([regressor.0.weight ; angle head] -> the head of `workspace`).
 * vpr_heads_loss_fwd_bwd does this itself unless VPR_HEADS_PREPACKED is set; a caller that knows the weights have
 * not changed since the last call on the same workspace, or that wants the cast off the critical path (a side
 * stream next to the SALAD forward), calls it explicitly. */
VPR_API int vpr_heads_pack(const VprHeadsParams* p, void* workspace, size_t workspace_bytes, const VprHeadsShape* s,
                   int dtype, cudaStream_t stream);
/* The memsets vpr_heads_loss_fwd_bwd starts with (split-K accumulator in the workspace, the three losses, the gradient
 * buffers that are reduced into with atomics), for callers that take them off the critical path (VPR_HEADS_PREZEROED). */
VPR_API int vpr_heads_zero(float* losses, const VprHeadsGrads* grads, void* workspace, size_t workspace_bytes,
                   const VprHeadsShape* s, int dtype, cudaStream_t stream);
/* desc: [B, D] (`dtype`); latlon_target: [B,2] fp32 (z-scored); angle_target_deg: [B] fp32.
 * Outputs: latlon_pred [B,2] fp32, sincos_pred [B,2] fp32, losses [3] fp32 (lat/lon, angle, their sum),
 * grads (overwritten; NULL with VPR_HEADS_FORWARD_ONLY), d_desc [B, D] (`dtype`) or NULL. */
VPR_API int vpr_heads_loss_fwd_bwd(const VprHeadsParams* p, const void* desc, const float* latlon_target,
                           const float* angle_target_deg, float* latlon_pred, float* sincos_pred,
                           float* losses, VprHeadsGrads* grads, void* d_desc, void* workspace,
                           size_t workspace_bytes, const VprHeadsShape* s, int dtype, cudaStream_t stream);

/* ---- backbone -> head seam: final LayerNorm + CLS/patch split + cast ----------
 * Replaces, in one pass, [UPSTREAM] DINOv2 forward_features' `x_norm = self.norm(x)`,
 * `x_norm[:, 0]` / `x_norm[:, 1 + registers:]` and the reshape/permute the reference's hub model does
 * before its aggregator (reached from dinov2salad/dinov2salad_finetuning.py:36).
 * hidden: [B, 1 + skip + n, d] (`dtype_in`, pre-LayerNorm); x_tok: [B, n, d], t: [B, d] (`dtype_out`);
 * stats: [B * (1 + n), 2] fp32 (mean, rstd) or NULL when no backward will follow. */
VPR_API int vpr_seam_ln_split_fwd(const void* hidden, const float* ln_weight, const float* ln_bias, float eps, void* x_tok,
                          void* t, float* stats, int B, int n, int skip, int d, int dtype_in, int dtype_out,
                          cudaStream_t stream);
/* d_x_tok / d_t may be NULL (zero). d_hidden: [B, 1 + skip + n, d] (`dtype_in`), fully written.
 * d_ln_weight / d_ln_bias: [d] fp32, overwritten (or NULL). */
VPR_API int vpr_seam_ln_split_bwd(const void* hidden, const float* ln_weight, const float* stats, const void* d_x_tok,
                          const void* d_t, void* d_hidden, float* d_ln_weight, float* d_ln_bias, int B, int n, int skip,
                          int d, int dtype_in, int dtype_out, cudaStream_t stream);

/* ---- validation metrics on the device ------------------------------------------
 * vpr_metrics_latlon: pred/target [N,2] fp32; scaler_mean/scale [2] fp64 device pointers (StandardScaler mean_,
 * scale_; NULL = identity), predictions are always de-normalised, targets only when `targets_scaled`
 * (dinov2salad_validation.py:84-85 keeps them in original units, dinov2salad_finetuning_2.py:202-203 does not).
 * out_sums [4] fp64 = sum dlat^2, sum dlon^2, sum |dlat|, sum |dlon|  ->  final_loss = 0.5*(s0+s1)/N
 * (dinov2salad_validation.py:101), MSE = (s0+s1)/(2N), MAE = (s2+s3)/(2N) (dinov2salad_finetuning_2.py:211-215).
 * vpr_metrics_angle: sincos_pred [N,2] fp32, true_deg [N] fp32, pred_deg [N] fp32 or NULL, out_sum [1] fp64 =
 * sum of angular errors (MAAE = out_sum / N). cos_first = 0: [sin,cos], `% 360` on both
 * (angle_prediction/dinov2salad/dino_v2_gemini.py:117-141); 1: [cos,sin], `(deg+360) % 360`
 * (angle_prediction/efficient_net/validation_script.py:180-189).
 * vpr_recall_at_k: idx [Q,k] int64 (vpr_knn_topk output), labels int64; ks [nk] int32 device pointer, nk <= 16;
 * hits [nk] uint64 = queries with a same-label neighbour among their first ks[j]. No reference code. */
VPR_API int vpr_metrics_latlon(const float* pred, const float* target, const double* scaler_mean, const double* scaler_scale,
                       int targets_scaled, int64_t N, double* out_sums, cudaStream_t stream);
VPR_API int vpr_metrics_angle(const float* sincos_pred, const float* true_deg, int cos_first, int64_t N, float* pred_deg,
                      double* out_sum, cudaStream_t stream);
VPR_API int vpr_recall_at_k(const int64_t* idx, const int64_t* gallery_label, const int64_t* query_label, int64_t Q, int k,
                    int64_t G, const int* ks, int nk, unsigned long long* hits, cudaStream_t stream);

/* ---- image preprocessing in front of the backbone ------------------------------
 * `transforms.Resize((h, w))` on a PIL image + `ToTensor()` + `Normalize(mean, std)`,
 * dinov2salad/dinov2salad_finetuning.py:45-50.  The resize is Pillow's 8-bit bilinear resampler (antialiased
 * separable triangle filter, 22-bit fixed-point coefficients, uint8 between the two passes) restated bit-exactly.
 * rgb_hwc: [H, W, 3] uint8 (device); mean3 / std3: HOST pointers to 3 floats; out_chw: [3, out_h, out_w] (`dtype_out`);
 * out_u8_hwc: optional [out_h, out_w, 3] uint8 copy of the resized image (NULL to skip). */
VPR_API size_t vpr_preprocess_workspace_bytes(int H, int W, int out_h, int out_w);
/* B images of one size, stacked [B, H, W, 3] -> [B, 3, out_h, out_w] with the same four launches as one image. */
VPR_API size_t vpr_preprocess_batch_workspace_bytes(int B, int H, int W, int out_h, int out_w);
VPR_API int vpr_preprocess_batch(const uint8_t* rgb_bhwc, int B, int H, int W, int out_h, int out_w, const float* mean3,
                         const float* std3, void* out_bchw, int dtype_out, uint8_t* out_u8_bhwc, void* workspace,
                         size_t workspace_bytes, cudaStream_t stream);
VPR_API int vpr_preprocess_image(const uint8_t* rgb_hwc, int H, int W, int out_h, int out_w, const float* mean3,
                         const float* std3, void* out_chw, int dtype_out, uint8_t* out_u8_hwc, void* workspace,
                         size_t workspace_bytes, cudaStream_t stream);

/* ---- descriptor kNN ------------------------------------------------------------------- */
VPR_API size_t vpr_knn_workspace_bytes(int64_t Q, int64_t G, int D, int k);
/* q: [Q, D] bf16, g: [G, D] bf16 (this rank's gallery shard), scores = q . g (fp32 accumulate).
 * out_score [Q,k] fp32 descending, out_idx [Q,k] int64 = g_offset + local row; ties -> lower index. */
VPR_API int vpr_knn_topk(const void* q, const void* g, int64_t Q, int64_t G, int D, int k, int64_t g_offset,
                 float* out_score, int64_t* out_idx, void* workspace, size_t workspace_bytes,
                 cudaStream_t stream);
/* Merge R per-rank results ([R, Q, k] score/idx, e.g. after an allgather) into the global top-k. */
VPR_API int vpr_knn_merge(const float* scores, const int64_t* idx, int R, int64_t Q, int k, float* out_score,
                  int64_t* out_idx, cudaStream_t stream);

/* ---- gradient all-reduce over NVLink peer memory (fine-tuning) ---------------------
 * No reference code (the reference is single-GPU); replaces ncclAllReduce(sum, fp32) for the two
 * gradient messages of one fine-tuning step.  Every rank holds one symmetric allocation
 *   [ data (any size) | staging (vpr_comm_staging_elems floats) | flags (vpr_comm_flag_bytes, zeroed once) ]
 * and passes every rank's addresses of the three areas (peer mappings: CUDA IPC or torch symmetric memory);
 * mc_data = multicast (NVLS) address of the data area or NULL.  The call sums, in place and on every rank,
 * floats [offset, offset + count) of the data area; all ranks must issue the same calls in the same order on
 * the same communicator (= flags area).  Two communicators may be in flight at once.  Stream-ordered,
 * capturable in a CUDA graph, `ctas` (<= VPR_COMM_MAX_CTAS) small CTAs that can share SMs with a running GEMM. */
#define VPR_COMM_MAX_WORLD 8
#define VPR_COMM_MAX_CTAS 160
typedef struct {
    int rank, world;
    void* data[VPR_COMM_MAX_WORLD];
    void* staging[VPR_COMM_MAX_WORLD];
    void* flags[VPR_COMM_MAX_WORLD];
    void* mc_data;
    size_t staging_elems;       /* floats available in each staging area */
} VprCommPeers;
VPR_API size_t vpr_comm_flag_bytes(void);
VPR_API size_t vpr_comm_staging_elems(size_t max_count, int world);
VPR_API int vpr_allreduce_sum_f32(const VprCommPeers* peers, size_t offset_elems, size_t count, int ctas,
                          cudaStream_t stream);

/* ---- test hook: the GEMM every dense step above is built from --------------------------- */
VPR_API int vpr_gemm(int dtype, const void* A, int a_mn, int64_t lda, int64_t a_bs, const void* B, int b_mn,
             int64_t ldb, int64_t b_bs, int M, int N, int K, int batch, void* out, int out_kind,
             int64_t ldo, int64_t out_bs, const float* bias, int relu, const void* mask, int64_t ld_mask,
             int64_t mask_bs, float mask_scale, float alpha, int k_splits, int bn, float* colsum,
             float dropout_p, uint64_t dropout_seed, cudaStream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* VPR_B200_H */
