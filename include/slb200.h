/*
 * slb200.h — C-ABI of libslb200.so: the B200 (sm_100a) kernels behind the
 * SemanticLens concept-database build path.
 *
 * Every entry point is `extern "C"`, takes plain device pointers + sizes + a
 * CUDA stream handle (a `cudaStream_t` passed as `void*`), never allocates or
 * frees device memory, never synchronises the host, and returns 0 on success
 * or a negative SLB_E* code (message via slb_last_error(), thread-local).
 * All device buffers (inputs, outputs, state, workspaces) are owned by the
 * caller; pointers are borrowed for the duration of the stream-ordered work.
 *
 * Each function names the reference call-site it replaces (paths relative to
 * the reference repo root, jim-berend/semanticlens v0.2.1).
 */
#ifndef SLB200_H
#define SLB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SLB_VERSION 100 /* 0.1.0 */

/* status codes */
#define SLB_OK 0
#define SLB_EINVAL (-1)       /* bad argument (null pointer, negative size, bad enum, misaligned) */
#define SLB_EUNSUPPORTED (-2) /* valid but not implemented for this shape/dtype */
#define SLB_ECUDA (-3)        /* a CUDA runtime/driver call failed */
#define SLB_EWORKSPACE (-4)   /* caller-provided workspace too small */

/* element types of activation maps handed to the collect kernels */
#define SLB_DT_F32 0
#define SLB_DT_F16 1
#define SLB_DT_BF16 2

/* memory layout of a hooked activation map */
#define SLB_LAYOUT_NCHW 0 /* (B, C, inner) contiguous; reduce over the innermost `inner` = H*W */
#define SLB_LAYOUT_BTF 1  /* (B, inner, C) contiguous; reduce over the middle `inner` = tokens   */

/* aggregation operators == the reference's aggregators
 * (semanticlens/component_visualization/aggregators.py) */
#define SLB_AGG_MEAN 0    /* aggregate_conv_mean :38-61, aggregate_transformer_mean :90-114 */
#define SLB_AGG_MAX 1     /* aggregate_conv_max :64-87, aggregate_transformer_max :144-168   */
#define SLB_AGG_ABSMEAN 2 /* aggregate_transformer_absmean :117-141 */
#define SLB_AGG_ABSMAX 3  /* aggregate_transformer_absmax :171-195  */
#define SLB_AGG_TOKEN 4   /* get_aggregate_transformer_special_token :198-244 (BTF only) */

int slb_version(void);
const char* slb_last_error(void);

/* Number of kernels this library has launched in this process so far (bench.py reports the delta over the
 * timed region as "gpu_launches"). */
int64_t slb_launch_count(void);

/* Live per-kernel timing for bench.py's roofline line: between slb_profile_begin() and slb_profile_end() every
 * single-kernel entry point records two CUDA events around its launch, on the stream it launches on, plus its
 * algorithmic work. slb_profile_summary synchronises those events and returns one entry per kernel name. */
typedef struct {
    char name[48];
    int64_t launches;
    double ms;    /* sum of event-to-event times */
    double flops; /* algorithmic flops issued (tensor kernels) */
    double bytes; /* algorithmic bytes (streaming kernels) */
} SlbKernelTime;
int slb_profile_begin(void);
int slb_profile_end(void);
int slb_profile_summary(SlbKernelTime* out, int max_entries, int* n_entries);

/* Number of SMs / compute capability of the current device (host query, for sizing and for
 * failing loudly on a non-sm_100 device). Returns SLB_ECUDA if no device. */
int slb_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ------------------------------------------------------------------------------------------
 * collect: aggregate + streaming top-k
 * ---------------------------------------------------------------------------------------- */

/* K1. out[b*C + c] = agg_{i<inner} x[b, c, i]   (NCHW)   or   agg_{t<inner} x[b, t, c]  (BTF),
 * fp32 accumulation in a fixed, shape-only-dependent order (bit-reproducible for any batch size),
 * result rounded once to `dtype` (as torch does for half inputs) and stored as fp32.
 * Replaces `tensor.clone().flatten(2).mean(-1).detach().cpu()` and its six siblings
 * (aggregators.py:61,87,114,141,168,195,242): one streaming read, no clone, no D2H.
 * token_pos is used by SLB_AGG_TOKEN only (python-style negatives allowed). */
int slb_agg_reduce(const void* x, int dtype, int layout, int64_t B, int64_t C, int64_t inner, int agg_op,
                   int64_t token_pos, float* out, void* stream);

/* K2. Merge B new candidates per latent into the per-latent sorted top-k state.
 * Replaces ActMax.update (activation_caching.py:112-141): `acts.T.to(bf16)`, repeat ids, 2x cat,
 * CPU torch.topk, gather.
 *   cand        (B, C) row-major, SLB_DT_F32 (rounded RNE to bf16 in-kernel, NaN -> 0x7FC0) or SLB_DT_BF16
 *   ids         optional (B,) int64 sample ids; NULL => id = id_base + b (activation_caching.py:410-413)
 *   state_vals  (C, k) bf16 bit patterns, sorted; fresh state = -0.0 (0x8000)  (activation_caching.py:108)
 *   state_ids   (C, k) int64; fresh state = -1                                  (activation_caching.py:109)
 * Order written: (value desc with +0 == -0 and NaN greatest, real ids ascending, placeholders last).
 * ids must be unique, 0 <= id < 2^47-1. k + B <= 8192. */
int slb_topk_update(const void* cand, int cand_dtype, int64_t B, int64_t C, const int64_t* ids, int64_t id_base,
                    uint16_t* state_vals, int64_t* state_ids, int64_t k, void* stream);

/* K1+K2 in one call (what the forward hook does, activation_caching.py:403-416).
 * scratch: >= B*C*4 bytes of device memory, 16-byte aligned. */
int slb_agg_topk_update(const void* x, int dtype, int layout, int64_t B, int64_t C, int64_t inner, int agg_op,
                        int64_t token_pos, int64_t id_base, uint16_t* state_vals, int64_t* state_ids, int64_t k,
                        void* scratch, size_t scratch_bytes, void* stream);

/* K2, list mode: merge R per-rank sorted states (R, C, k) into one (C, k) state. The exchange step after the
 * image-sharded sweep (no reference counterpart: the reference is single-process). R*k <= 8192.
 * out_* may not alias the inputs. */
int slb_topk_merge_lists(const uint16_t* vals, const int64_t* ids, int64_t R, int64_t C, int64_t k,
                         uint16_t* out_vals, int64_t* out_ids, void* stream);

/* K5. out[r, :] = table[idx[r] (python-negative wraps), :]   — `embeds[sample_ids]`
 * (activation_based.py:387-390; id -1 aliases the last row). Rows are `D` fp32. */
int slb_gather_rows(const float* table, int64_t N, int64_t D, const int64_t* idx, int64_t n_idx, float* out,
                    void* stream);

/* ------------------------------------------------------------------------------------------
 * analyze: similarity / clarity / polysemanticity  (semanticlens/scores.py)
 * ---------------------------------------------------------------------------------------- */

/* F.normalize(x, dim=-1, eps) fused with the split-plane conversion: planes [2, rows, Kpad] of x / max(|x|, eps),
 * Kpad = D rounded up to a multiple of 64 (zero padded), and/or inv_norms [rows] = 1 / max(|x|, eps).
 * The planes hold plane_scale * x / max(|x|, eps). Either output may be NULL. D % 4 == 0, D <= 2048. (scores.py:120-121) */
int slb_normalize_split_rows(const float* x, int64_t rows, int64_t D, float eps, int plane_fmt, float plane_scale,
                             uint16_t* planes, float* inv_norms, void* stream);

/* K6. out[M,N] = normalize(x)[M,D] . normalize(y)[N,D]^T — scores.similarity_score's matmul branch
 * (scores.py:119-125, reached from Lens.text_probing / image_probing via lens.py:207-214). fp32 out, fp32-grade
 * accuracy (3-pass split-fp16 tcgen05 GEMM). N % 8 == 0. workspace: 256-byte aligned, >= the query below. */
size_t slb_cosine_gemm_workspace_bytes(int64_t M, int64_t N, int64_t D);
int slb_cosine_gemm(const float* x, int64_t M, const float* y, int64_t N, int64_t D, float* out, void* workspace,
                    size_t workspace_bytes, void* stream);

/* out[r] = F.cosine_similarity(x[r], y[r], eps) — similarity_score's equal-shape branch (scores.py:127). */
int slb_cosine_rows(const float* x, const float* y, int64_t rows, int64_t D, float eps, float* out, void* stream);

/* K7. out[c] = ((|mean_k normalize(V[c])|^2 - 1/k) / (k-1)) * k — scores.clarity_score (scores.py:19-47), one streaming
 * pass over V (C, k, D) fp32. D % 4 == 0, D <= 2048. k = 1 gives inf/nan like the reference. */
int slb_clarity(const float* V, int64_t C, int64_t k, int64_t D, float* out, void* stream);

/* K8. out[c] = polysemanticity of neuron c (float64) — scores.polysemanticity_score (scores.py:132-185): sklearn
 * KMeans(n_clusters=2, n_init, random_state) per neuron, 1 - cos(centre_1, centre_2); neurons whose smaller cluster has
 * < 2 members take the fallback of scores.py:173-184 when replace_empty_clusters != 0.
 * V (C, k, D) fp32 device, k <= 256. first_centers [n_init] / local_trial_uniforms [n_init][2] are HOST arrays: the
 * data-independent draws of sklearn's RandomState stream (per init `choice(k)` then `uniform(size=2)`, _kmeans.py:231,249).
 * workspace: 16-byte aligned device memory >= slb_polysem_workspace_bytes(C, k). */
size_t slb_polysem_workspace_bytes(int64_t C, int64_t k);
/* Diagnostic: SM clocks that thread 0 of every polysemanticity CTA spent per phase since the last reset, summed over CTAs
 * (synchronises the device): [0] Gram matrix, [1] row means, [2] k-means++ and first assignments, [3] first-iteration
 * sums, [4] Lloyd restarts, [5] score, [6] neurons processed. */
int slb_polysem_phase_clocks(uint64_t* out7, int reset);
int slb_polysem_2means(const float* V, int64_t C, int64_t k, int64_t D, const int64_t* first_centers,
                       const double* local_trial_uniforms, int n_init, int replace_empty_clusters, double* out,
                       void* workspace, size_t workspace_bytes, void* stream);

/* K8g. The general form of K8: any n_clusters in [2, 8] and any number of examples per neuron (scores.py:132-185 called with
 * n_clusters != 2, or a concept DB with more than 256 examples per neuron): sklearn KMeans(n_clusters, n_init,
 * random_state).fit per neuron in float64 in sample space, then 1 - clarity of the centres (or the fallback when a cluster
 * has fewer than 2 members). first_centers [n_init] and local_trial_uniforms [n_init][n_clusters - 1][n_local_trials] are
 * HOST arrays (n_local_trials = 2 + int(log n_clusters), sklearn _kmeans.py:226); they are copied on `stream`.
 * k >= n_clusters (sklearn raises otherwise). workspace >= slb_polysem_kmeans_workspace_bytes(...), 16-byte aligned. */
size_t slb_polysem_kmeans_workspace_bytes(int64_t C, int64_t k, int64_t D, int n_clusters, int n_init);
int slb_polysem_kmeans(const float* V, int64_t C, int64_t k, int64_t D, int n_clusters, const int64_t* first_centers,
                       const double* local_trial_uniforms, int n_local_trials, int n_init, int replace_empty_clusters,
                       double* out, void* workspace, size_t workspace_bytes, void* stream);

/* K9 (part). out[i] = max_j (S[i, j] - 2 * (j == row0 + i)) for a row block S (rows, cols) of the cosine matrix —
 * scores.redundancy_score's `(sims - 2 I).max(-1)` (scores.py:78-80) without materialising the identity. */
int slb_rowmax_offdiag(const float* S, int64_t rows, int64_t cols, int64_t row0, float* out, void* stream);

/* K9. redundancy_score (reference scores.py:51-81): out[0] = mean_i max_{j != i} cos(cones_i, cones_j) for cones (n, D)
 * fp32. The rows are L2-normalised into split planes once and ONE tensor-core GEMM of the planes against themselves keeps
 * only the running row maxima (diagonal lowered by 2, like `sims - 2 I`) in its epilogue: the n x n cosine matrix is
 * never written. NaN rows propagate like torch.max. D % 4 == 0, D <= 2048.
 * workspace: 256-byte aligned device memory >= slb_redundancy_workspace_bytes(n, D). */
size_t slb_redundancy_workspace_bytes(int64_t n, int64_t D);
int slb_redundancy(const float* cones, int64_t n, int64_t D, float* out, void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * embed: the CLIP / SigLIP ViT image tower  (foundation_models/clip.py:103-163 -> open_clip, not vendored)
 * ---------------------------------------------------------------------------------------- */

/* 16-bit "split plane" operands: s * x ~= hi + lo with hi = rn16(s * x), lo = rn16(s * x - hi), s a power of two chosen
 * per tensor so that typical magnitudes sit well inside the fp16 normal range (the lo plane then keeps its 11 bits; a lo
 * that underflows costs at most 2^-25 absolute). Because hi and lo share ONE scale, all three plane products of a GEMM
 * (hi.hi, hi.lo, lo.hi) accumulate in a single tensor-memory accumulator.
 * F16: 22 significant bits, |s * x| <= 65504 (saturating). BF16: 16 significant bits, fp32 range. */
#define SLB_PLANE_F16 0
#define SLB_PLANE_BF16 1
#define SLB_ACT_PLANE_SCALE 16.0f      /* every plane WRITTEN by a kernel (LayerNorm, attention, GEMM epilogue, patchify) */
#define SLB_WEIGHT_PLANE_SCALE 1024.0f /* what the towers expect of their weight planes (slb_split_planes(..., 1024, ...)) */

/* x fp32 [n] -> planes [2][n] (hi plane, then lo plane) of scale * x. n % 4 == 0. */
int slb_split_planes(const float* x, int64_t n, int plane_fmt, float scale, uint16_t* planes, void* stream);

/* epilogue selectors for slb_gemm_split */
#define SLB_EPI_NONE 0
#define SLB_EPI_GELU_ERF 1  /* nn.GELU()            (open_clip ViT-B-32 / ViT-L-14) */
#define SLB_EPI_QUICKGELU 2 /* x * sigmoid(1.702 x)  (open_clip *-quickgelu, OpenAI weights) */
#define SLB_EPI_GELU_TANH 3 /* nn.GELU("tanh")       (SigLIP / timm towers) */
#define SLB_EPI_RELU 4      /* max(z, 0)             (conv + BatchNorm + ReLU of the CLIP ModifiedResNet) */
#define SLB_EPI_ADD_RELU 5  /* max(z + residual, 0)  (the tail of a ResNet bottleneck: the ReLU follows the shortcut add) */
#define SLB_EPI_ADD_RELU_PLANES 6 /* the same, but `residual` points at the shortcut as split planes: const uint16_t [2][M][N] in the
                                   * GEMM's plane format at the activation scale (what out_planes of the previous block holds), cast to
                                   * const float*. A residual stream kept as planes never needs an fp32 copy of a block's output. */

#define SLB_PASSES_SPLIT_ACC 4 /* `passes` selector of slb_gemm_split, see below */

/* K4/K6. D[M,N] = act(alpha * (A[M,K] * W[N,K]^T) * row_scale[m] * col_scale[n] + bias[n]) + residual[M,N]
 * (SLB_EPI_ADD_RELU applies its max after the residual instead)
 * on the 5th-gen tensor cores (tcgen05.mma kind::f16, TMA-fed 128B-swizzled stages, fp32 accumulators in TMEM) with
 * fp32-grade accuracy: operands are split planes a_planes [2,M,K], w_planes [2,N,K] (row-major, K contiguous) and each
 * tile accumulates Ahi*Whi + Ahi*Wlo + Alo*Whi (passes = 3) or Ahi*Whi only (passes = 1) in ONE accumulator.
 * passes = SLB_PASSES_SPLIT_ACC computes the same three products but keeps the two cross terms in a second
 * accumulator that the epilogue adds in fp32: the tensor core truncates at every accumulate, so a third of the adds
 * into the large accumulator means a third of the (systematic) error — measured 2e-6 instead of 6e-6 of the largest
 * output at K = 1152. It runs on the one-CTA 128 x 128 tile only (no 256 x 256 pair tile: TMEM is full), so it is
 * for long chains of GEMMs without a normalisation in between (the ModifiedResNet), not for the ViT towers.
 * Replaces the fp32 GEMMs of open_clip's image tower (clip.py:118) and the cosine matmul of
 * scores.similarity_score (scores.py:120-125; row_scale/col_scale = inverse row norms).
 * Outputs: out_f32 [M,N] (nullable) and/or out_planes [2,M,N] (nullable) in the same plane format, ready to be the
 * A operand of the next GEMM. bias/residual/row_scale/col_scale nullable; residual may alias out_f32.
 * Requirements: K % 64 == 0, N % 8 == 0, all buffers 16-byte aligned. */
int slb_gemm_split(const uint16_t* a_planes, const uint16_t* w_planes, int plane_fmt, int64_t M, int64_t N, int64_t K,
                   float alpha /* 1 / (scale of A planes * scale of W planes) */, const float* bias, const float* residual,
                   const float* row_scale, const float* col_scale, int epilogue, int passes, float* out_f32,
                   uint16_t* out_planes /* written at SLB_ACT_PLANE_SCALE */, void* stream);

/* K3. out[b,c,y,x] = (u8[b,c,y,x]/255 - mean[c]) / std[c]  — `ToTensor` + `Normalize` of the open_clip eval transform
 * (clip.py:157-160) for already-sized planar u8 images, in torch's op order (two IEEE divisions).
 * mean3/std3 are HOST arrays of Cc floats. n_pix = H*W, a multiple of 16. */
int slb_u8_to_f32_norm(const uint8_t* img, int64_t B, int64_t Cc, int64_t n_pix, const float* mean3,
                       const float* std3, float* out, void* stream);

/* Bicubic resize + crop of one RGB image, bit-exact with Pillow's ImagingResample (what torchvision
 * Resize(S, BICUBIC) + CenterCrop(S) of the open_clip eval transform run per PIL image; clip.py:157-160):
 * src_hwc (h, w, 3) u8 is resampled to out_w x out_h (Keys a = -0.5 kernel, support scaled when shrinking, 22-bit
 * fixed-point coefficients, horizontal pass into an 8-bit intermediate, then vertical), of which only the window
 * [crop_top, +crop_h) x [crop_left, +crop_w) is produced, channels first: dst_chw (3, crop_h, crop_w) u8.
 * workspace: 256-byte aligned, >= slb_resize_workspace_bytes(...) (0 on bad sizes). */
size_t slb_resize_workspace_bytes(int64_t h, int64_t w, int64_t out_w, int64_t out_h, int64_t crop_left, int64_t crop_top,
                                  int64_t crop_w, int64_t crop_h);
int slb_resize_bicubic_u8(const uint8_t* src_hwc, int64_t h, int64_t w, int64_t out_w, int64_t out_h, int64_t crop_left,
                          int64_t crop_top, int64_t crop_w, int64_t crop_h, uint8_t* dst_chw, void* workspace,
                          size_t workspace_bytes, void* stream);

/* (B,3,S,S) fp32 -> im2col rows as split planes [2][B*(S/P)^2][slb_patch_k(P)] (row = (b, gy, gx), col = (c, py, px),
 * zero padded from 3*P*P up to a multiple of 64), so that the patch-embedding conv (kernel = stride = P, even P) is
 * one slb_gemm_split against the flattened, equally padded conv weight. */
int64_t slb_patch_k(int64_t P);
int slb_patchify(const float* img, int64_t B, int64_t S, int64_t P, int plane_fmt, uint16_t* out_planes, void* stream);

/* x[b,t,:] = (has_cls && t == 0 ? cls : patch[b, t - has_cls, :]) + pos[t,:]   (fp32; pos nullable). */
int slb_assemble_tokens(const float* patch, const float* cls, const float* pos, int64_t B, int64_t T, int64_t W,
                        int has_cls, float* out, void* stream);

/* Row LayerNorm over `cols` (fp32, two-pass, biased variance, eps inside the sqrt like torch); input rows are
 * `row_stride` floats apart (e.g. T*W to normalise only the class tokens). Writes fp32 [rows,cols] (nullable) and/or
 * split planes [2,rows,cols] (nullable). beta nullable. */
int slb_layernorm(const float* x, int64_t rows, int64_t cols, int64_t row_stride, const float* gamma, const float* beta,
                  float eps, int plane_fmt, float* out_f32, uint16_t* out_planes, void* stream);

/* softmax(scale * Q K^T) V per (batch, head) for short sequences (Tk <= 320, head_dim % 4 == 0, <= 128), fp32.
 * q: row (b, i) at q + b*q_batch_stride + i*q_row_stride, head h at + h*dh; k, v likewise with the kv strides
 * (packed nn.MultiheadAttention in_proj output: k = q + W, v = q + 2W, row stride 3W).
 * out [B,Tq,H*dh] as fp32 (nullable) and/or split planes [2, B*Tq, H*dh] (nullable). */
int slb_attention_small(const float* q, int64_t q_batch_stride, int64_t q_row_stride, const float* k, const float* v,
                        int64_t kv_batch_stride, int64_t kv_row_stride, int64_t B, int64_t Tq, int64_t Tk, int64_t H,
                        int64_t dh, float scale, int plane_fmt, float* out_f32, uint16_t* out_planes, void* stream);

/* The same attention for head_dim 64 reading q, k, v from the fp16 split planes [2][B*T][3*H*64] (at SLB_ACT_PLANE_SCALE)
 * that the in_proj GEMM emits
 * (packed nn.MultiheadAttention layout: q | k | v along the last axis): no fp32 round trip, no conversion in the
 * kernel (cp.async + ldmatrix + mma.sync). out as in slb_attention_small. */
int slb_attention_planes(const uint16_t* qkv_planes, int64_t B, int64_t T, int64_t H, int64_t dh, float scale,
                         int causal /* 1: key j visible to query i only if j <= i */, int plane_fmt, float* out_f32,
                         uint16_t* out_planes, void* stream);

/* Debug aid: with SLB_ATTN_TRACE=1 in the environment, CTA (0, 0) of the last tcgen05 attention launch records the SM-clock
 * offsets of its producer / MMA / softmax hand-offs into 256 host-mapped words (event type t, index i at [64 + 16 t + i]);
 * NULL otherwise. Read after synchronising the stream (scripts/trace_attention.py). */
const unsigned int* slb_attention_trace(void);

/* The whole CLIP / SigLIP ViT image tower in one call (open_clip VisionTransformer.forward, or its timm ViT with the
 * attention-pool head for SigLIP, behind clip.py:103-118).
 * Weight matrices are split planes prepared once by the caller (slb_split_planes); vectors are fp32. All pointers are
 * device pointers except `layer`, a HOST array of `layers` structs. */
#define SLB_POOL_CLS 0 /* features = ln_post(x)[:, 0] @ proj                      (CLIP VisionTransformer) */
#define SLB_POOL_MAP 1 /* features = attention-pool head over ln_post(x) [@ proj]  (SigLIP / timm AttentionPoolLatent) */

typedef struct {
    const float* ln1_g; const float* ln1_b;
    const uint16_t* w_qkv; const float* b_qkv;   /* attn.in_proj: planes [2, 3W, W], [3W] */
    const uint16_t* w_out; const float* b_out;   /* attn.out_proj: planes [2, W, W], [W] */
    const float* ln2_g; const float* ln2_b;
    const uint16_t* w_fc; const float* b_fc;     /* mlp.c_fc: planes [2, mlp, W], [mlp] */
    const uint16_t* w_proj; const float* b_proj; /* mlp.c_proj: planes [2, W, mlp], [W] */
} SlbVitLayer;

typedef struct {
    int32_t image_size, patch, width, layers, heads, mlp, embed_dim;
    int32_t act;       /* SLB_EPI_* of the MLP */
    int32_t plane_fmt; /* SLB_PLANE_* of every plane below */
    int32_t has_cls, pool;
    float ln_eps;
    const uint16_t* conv_w; /* planes [2, W, slb_patch_k(patch)] (conv1.weight flattened (c,py,px), zero padded) */
    const float* conv_b;    /* [W] or NULL (CLIP has no conv bias) */
    const float* cls;       /* [W] */
    const float* pos;       /* [T, W] */
    const float* ln_pre_g; const float* ln_pre_b;   /* NULL: no ln_pre */
    const float* ln_post_g; const float* ln_post_b;
    const uint16_t* proj;   /* planes [2, embed_dim, W] (= visual.proj transposed) or NULL */
    const SlbVitLayer* layer;
    /* SLB_POOL_MAP only: a learned latent query attends over the final tokens, then x + mlp(norm(x)). */
    const float* map_q;                                /* [W] q-projection of the latent (image independent, incl. bias) */
    const uint16_t* map_w_kv; const float* map_b_kv;   /* planes [2, 2W, W], [2W]: k | v projections of the tokens */
    const uint16_t* map_w_out; const float* map_b_out; /* planes [2, W, W], [W] */
    const float* map_ln_g; const float* map_ln_b;      /* [W] */
    const uint16_t* map_w_fc; const float* map_b_fc;   /* planes [2, mlp, W], [mlp] */
    const uint16_t* map_w_proj; const float* map_b_proj; /* planes [2, W, mlp], [W] */
} SlbVitWeights;

/* Bytes of device workspace slb_vit_forward needs for a batch of B images (0 on bad arguments). */
size_t slb_vit_workspace_bytes(const SlbVitWeights* w, int64_t B);

/* img (B,3,S,S) fp32 preprocessed -> out (B, embed_dim) fp32, un-normalised like open_clip's encode_image.
 * workspace: 256-byte aligned device memory of at least slb_vit_workspace_bytes(w, B). */
int slb_vit_forward(const SlbVitWeights* w, const float* img, int64_t B, float* out, void* workspace,
                    size_t workspace_bytes, void* stream);

/* The trunk of a ViT in slices (the accelerated probed forward: torchvision's VisionTransformer under forward hooks on its
 * encoder blocks, run by the reference as self.model(x), activation_based.py:341-358). layer_begin == 0 embeds the images
 * first; then blocks [layer_begin, layer_end) run. The residual stream (B*T, W) fp32 — the output of block layer_end - 1 —
 * is the first region of the workspace (slb_vit_workspace_bytes). The pooling fields of `w` are not used. */
int slb_vit_trunk(const SlbVitWeights* w, const float* img, int64_t B, int32_t layer_begin, int32_t layer_end, void* workspace,
                  size_t workspace_bytes, void* stream);

/* The CLIP / SigLIP text tower in one call (open_clip CLIP.encode_text / CustomTextCLIP.encode_text behind
 * clip.py:120-135; reached from Lens.text_probing, lens.py:166-203): token + positional embedding, `layers` pre-LN blocks
 * (causal mask for CLIP, none for SigLIP), ln_final, the pooled token's feature (CLIP: end-of-text; SigLIP: last position)
 * times text_projection (+ bias for SigLIP); un-normalised. head_dim must be 64 and plane_fmt fp16 (tensor-core attention). */
typedef struct {
    int32_t context, vocab, width, layers, heads, mlp, embed_dim;
    int32_t act, plane_fmt;
    float ln_eps;
    const float* tok_emb;  /* [vocab, W] */
    const float* pos;      /* [context, W] */
    const float* ln_final_g; const float* ln_final_b;
    const uint16_t* proj;  /* planes [2, embed_dim, W] (= text_projection transposed) */
    const SlbVitLayer* layer; /* HOST array */
    /* SigLIP text towers (open_clip TextTransformer with no_causal_mask, pool_type "last", proj_bias; reached through
     * clip.py:190-215 SigLipV2.encode_text): bidirectional attention, a biased projection; the caller passes the LAST
     * position of every sequence as eot_rows. CLIP: non_causal = 0, proj_b = NULL. */
    int32_t non_causal;
    const float* proj_b;   /* [embed_dim] or NULL */
} SlbTextWeights;

size_t slb_text_workspace_bytes(const SlbTextWeights* w, int64_t B);

/* tokens (B, context) int64 device; eot_rows (B,) int64 device = b * context + argmax_t tokens[b, t] (computed by the
 * caller); out (B, embed_dim) fp32. workspace: 256-byte aligned, >= slb_text_workspace_bytes(w, B). */
int slb_text_forward(const SlbTextWeights* w, const int64_t* tokens, const int64_t* eot_rows, int64_t B, float* out,
                     void* workspace, size_t workspace_bytes, void* stream);

/* ---- CLIP ModifiedResNet image tower (open_clip "RN50" / "RN101" behind clip.py:103-118; BASELINE configs[0]) ----
 * Activations are channels-last split planes [2, B*H*W, C] at SLB_ACT_PLANE_SCALE, so every 1x1 convolution IS a
 * slb_gemm_split over them and a 3x3 convolution is the same GEMM over im2col planes; eval-mode BatchNorm rides in the
 * GEMM epilogue as col_scale = gamma / sqrt(var + eps), bias = beta - mean * col_scale, followed by the ReLU. */

/* Columns of the im2col matrix of a k x k convolution over cin channels: cin*k*k rounded up to a multiple of 64. */
int64_t slb_conv_k(int64_t cin, int64_t ksize);

/* im2col of the stem's first convolution (3x3, stride 2, pad 1) straight from the preprocessed images:
 * img (B,3,S,S) fp32 NCHW -> planes [2, B*(S/2)^2, 64], column (ky*3 + kx)*3 + c, zero past 27. S must be even. */
int slb_im2col_stem(const float* img, int64_t B, int64_t S, int plane_fmt, uint16_t* out_planes, void* stream);

/* The stem's first convolution computed directly (27 products per output do not pay for an im2col matrix and a tensor-core
 * pass): img (B,3,S,S) fp32 NCHW, w_planes [2, cout, 64] (columns (ky*3 + kx)*3 + c, at SLB_WEIGHT_PLANE_SCALE), 3x3 / stride 2 /
 * pad 1, then max(z * scale[n] + shift[n], 0) (eval BatchNorm + ReLU) -> out_planes [2, B*(S/2)^2, cout] at
 * SLB_ACT_PLANE_SCALE. fp32 accumulation in (ky, kx, c) order. cout = 32 or 64, S even. (open_clip ModifiedResNet.conv1/bn1/act1) */
int slb_stem_conv3x3s2(const float* img, int64_t B, int64_t S, const uint16_t* w_planes, int64_t cout, int plane_fmt,
                       const float* scale, const float* shift, uint16_t* out_planes, void* stream);

/* im2col of a 3x3, stride 1, pad 1 convolution over channels-last planes: in [2, B*H*W, C] ->
 * out [2, B*H*W, slb_conv_k(C, 3)], column (ky*3 + kx)*C + c (weights are laid out (cout, ky, kx, cin) to match),
 * zero outside the image and past 9C. C must be a multiple of 8. Plane bits are moved untouched. */
int slb_im2col3x3(const uint16_t* in_planes, int64_t B, int64_t H, int64_t W, int64_t C, uint16_t* out_planes, void* stream);

/* nn.AvgPool2d(2) over channels-last planes: in [2, B*H*W, C] -> out [2, B*(H/2)*(W/2), C]; H, W even, C % 8 == 0.
 * The four values are summed in fp32 from hi + lo and re-split. */
int slb_avgpool2_planes(const uint16_t* in_planes, int64_t B, int64_t H, int64_t W, int64_t C, int plane_fmt,
                        uint16_t* out_planes, void* stream);

/* ---- accelerated probed-model forward (opt-in; semanticlens_b200/probed.py) -------------------------------------------
 * The reference runs the user's probed model as `self.model(x)` under forward hooks (activation_based.py:341-358). For
 * torchvision-style ResNets (Conv2d / BatchNorm2d / ReLU / MaxPool2d: the probed models of BASELINE configs[0..3]) the
 * same maps can be produced by slb_gemm_split over channels-last planes; these entries are what that needs beyond the
 * CLIP tower's. Hooked maps leave as channels-last fp32 (B, H*W, C), which slb_agg_reduce reads as SLB_LAYOUT_BTF. */

/* Implicit-GEMM convolution: out[(b, yo, xo), n] = epi(alpha * sum_{ky,kx,c} x[b, yo*stride + ky - pad, xo*stride + kx - pad, c]
 * * w[n, (ky*ksize + kx)*C + c] * col_scale[n] + bias[n]) (+ residual), the arguments of slb_gemm_split otherwise. x_planes
 * [2, B*H*W, C] channels-last; w_planes [2, N, slb_conv_k(C, ksize)]. The im2col matrix is never written: the GEMM's producer
 * warp fetches (filter tap, 64-channel chunk) k-blocks with TMA im2col-mode loads (zero fill outside the image, stride in the
 * tensor map's traversal strides). C % 64 == 0, or C == 32 (a k-block is then one whole tap: 64-byte rows under the 64-byte
 * swizzle, two MMA steps — the 3x3 convolutions of the CLIP ModifiedResNet stem); N % 8 == 0, stride 1 or 2, ksize <= 7. */
int slb_conv_gemm(const uint16_t* x_planes, int64_t B, int64_t H, int64_t W, int64_t C, int ksize, int stride, int pad,
                  const uint16_t* w_planes, int64_t N, int plane_fmt, float alpha, const float* bias, const float* residual,
                  const float* col_scale, int epilogue, int passes, float* out_f32, uint16_t* out_planes, void* stream);

/* slb_gemm_split / slb_conv_gemm that ALSO leave raw_f32 [M, N] = alpha * (A W^T) * row_scale — the value before column
 * scale, bias, activation and residual: the raw output of a convolution whose BatchNorm / ReLU / shortcut ride in the same
 * epilogue, for a forward hook registered on the nn.Conv2d itself (BASELINE configs[3]: all 53 convolutions hooked). */
int slb_gemm_split_raw(const uint16_t* a_planes, const uint16_t* w_planes, int plane_fmt, int64_t M, int64_t N, int64_t K, float alpha,
                       const float* bias, const float* residual, const float* row_scale, const float* col_scale, int epilogue,
                       int passes, float* out_f32, uint16_t* out_planes, float* raw_f32, void* stream);
int slb_conv_gemm_raw(const uint16_t* x_planes, int64_t B, int64_t H, int64_t W, int64_t C, int ksize, int stride, int pad,
                      const uint16_t* w_planes, int64_t N, int plane_fmt, float alpha, const float* bias, const float* residual,
                      const float* col_scale, int epilogue, int passes, float* out_f32, uint16_t* out_planes, float* raw_f32,
                      void* stream);

/* im2col of a ksize x ksize / stride / pad convolution straight from NCHW fp32 images (the 7x7 / 2 / 3 stem):
 * img (B,C,H,W) -> planes [2, B*Ho*Wo, slb_conv_k(C, ksize)], column (ky*ksize + kx)*C + c, zero outside / past C k k. */
int slb_im2col_nchw(const float* img, int64_t B, int64_t C, int64_t H, int64_t W, int ksize, int stride, int pad, int plane_fmt,
                    uint16_t* out_planes, void* stream);

/* slb_im2col3x3 with a stride of 1 or 2 (pad 1): out [2, B*Ho*Wo, slb_conv_k(C, 3)], Ho = (H - 1) / stride + 1. */
int slb_im2col3x3_strided(const uint16_t* in_planes, int64_t B, int64_t H, int64_t W, int64_t C, int stride, uint16_t* out_planes,
                          void* stream);

/* Every second pixel of every second row (the operand of a 1x1 / stride 2 convolution): [2, B*H*W, C] ->
 * [2, B*ceil(H/2)*ceil(W/2), C]. */
int slb_subsample2_planes(const uint16_t* in_planes, int64_t B, int64_t H, int64_t W, int64_t C, uint16_t* out_planes, void* stream);

/* y = raw * scale[c] + shift[c] (+ residual), optional ReLU, over a channels-last fp32 map [M, C]: the BatchNorm / shortcut
 * / ReLU that follows a convolution whose RAW output a forward hook has to see. out_f32 [M, C] and / or out_planes
 * [2, M, C] (nullable); residual may alias out_f32. C % 8 == 0. */
int slb_affine_act(const float* raw, int64_t M, int64_t C, const float* scale, const float* shift, const float* residual, int relu,
                   int plane_fmt, float* out_f32, uint16_t* out_planes, void* stream);

/* BatchNorm (scale / shift) + ReLU + MaxPool2d(3, stride 2, pad 1) over a channels-last fp32 map (B, H, W, C) ->
 * (B, Ho, Wo, C), Ho = (H - 1) / 2 + 1, as fp32 and / or planes (nullable). C % 8 == 0. */
int slb_bn_relu_maxpool(const float* raw, int64_t B, int64_t H, int64_t W, int64_t C, const float* scale, const float* shift,
                        int plane_fmt, float* out_f32, uint16_t* out_planes, void* stream);

/* Tokens of open_clip's AttentionPool2d: x (B, HW, C) fp32 channels-last feature map, pos (HW+1, C) ->
 * tok_planes [2, B*(HW+1), C] = [mean over positions; positions] + pos, and query_planes [2, B, C] = token 0 of
 * every image (the only query whose output is kept). */
int slb_pool_tokens(const float* x, const float* pos, int64_t B, int64_t HW, int64_t C, int plane_fmt, uint16_t* tok_planes,
                    uint16_t* query_planes, void* stream);

typedef struct {
    const uint16_t* w;  /* planes [2, cout, slb_conv_k(cin, ksize)]: conv weight permuted to (cout, ky, kx, cin), zero padded */
    const float* scale; /* [cout] gamma / sqrt(running_var + eps) */
    const float* shift; /* [cout] beta - running_mean * scale */
    int32_t cin, cout, ksize, reserved;
} SlbConvBn;

typedef struct {
    int32_t image_size, width, heads, out_dim;
    int32_t blocks[4]; /* bottlenecks per stage, e.g. {3, 4, 6, 3} */
    int32_t plane_fmt, n_convs;
    /* HOST array in execution order: stem conv1, conv2, conv3; then per bottleneck conv1, conv2, conv3 and, for the first
     * bottleneck of every stage, its downsample convolution. n_convs = 3 + sum(3 * blocks[i] + 1). */
    const SlbConvBn* convs;
    const float* pos;                          /* [(S/32)^2 + 1, 32 * width] attnpool.positional_embedding */
    const uint16_t* w_q; const float* b_q;     /* planes [2, E, E], [E]       (E = 32 * width) */
    const uint16_t* w_kv; const float* b_kv;   /* planes [2, 2E, E], [2E]: k_proj | v_proj */
    const uint16_t* w_c; const float* b_c;     /* planes [2, out_dim, E], [out_dim] */
} SlbRnWeights;

size_t slb_rn_workspace_bytes(const SlbRnWeights* w, int64_t B);

/* img (B,3,S,S) fp32 preprocessed -> out (B, out_dim) fp32, un-normalised like open_clip's encode_image.
 * width must be a multiple of 64 and S of 32. workspace: 256-byte aligned, >= slb_rn_workspace_bytes(w, B). */
int slb_rn_forward(const SlbRnWeights* w, const float* img, int64_t B, float* out, void* workspace, size_t workspace_bytes,
                   void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SLB200_H */
