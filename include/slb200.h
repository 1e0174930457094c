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

#ifdef __cplusplus
}
#endif
#endif /* SLB200_H */
