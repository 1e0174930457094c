// The CLIP ViT image tower as ONE C-ABI call: (B,3,S,S) preprocessed fp32 images -> (B, embed_dim) features.
//
// Replaces `self.model.encode_image(img)` of the reference (foundation_models/clip.py:103-118), i.e. open_clip's
// VisionTransformer.forward: conv1 patch embedding (as a GEMM over im2col planes), class token + positional
// embedding, ln_pre, `layers` pre-LN residual blocks (nn.MultiheadAttention + MLP), ln_post on the class token, proj.
// Every dense contraction is slb_gemm_split (tcgen05, split planes, fp32-grade); LayerNorm and attention emit the split
// planes the next GEMM consumes, bias / activation / residual are GEMM epilogues. The residual stream stays fp32.
// No allocation: the caller supplies the workspace (slb_vit_workspace_bytes). Nothing synchronises.
#include "tc_common.cuh"

#include <algorithm>

namespace {

// every GEMM of the towers multiplies activation planes (SLB_ACT_PLANE_SCALE) with weight planes (SLB_WEIGHT_PLANE_SCALE)
constexpr float kAlpha = 1.0f / (SLB_ACT_PLANE_SCALE * SLB_WEIGHT_PLANE_SCALE);

struct WsLayout {
    size_t x, qkv, planes_a, planes_b, patch_f32, head_f32, total;
};

size_t align_up(size_t v) { return (v + 255) & ~(size_t)255; }

bool layout_for(const SlbVitWeights* w, int64_t B, WsLayout* L) {
    if (!w || w->patch <= 0 || w->image_size % w->patch) return false;
    const int64_t g = w->image_size / w->patch;
    const int64_t T = g * g + (w->has_cls ? 1 : 0);
    const int64_t W = w->width, rows = B * T;
    const int64_t Kc = slb_patch_k(w->patch);
    size_t o = 0;
    L->x = o;          o += align_up((size_t)rows * W * 4);
    L->qkv = o;        o += align_up((size_t)rows * 3 * W * 4);
    // planes_a: LN output / attention output planes [2, rows, W]  — also holds the im2col planes [2, B*g*g, Kc]
    size_t pa = std::max<size_t>((size_t)2 * rows * W * 2, (size_t)2 * B * g * g * Kc * 2);
    L->planes_a = o;   o += align_up(pa);
    // planes_b: MLP hidden planes [2, rows, mlp] — also the attention output planes
    size_t pb = std::max<size_t>((size_t)2 * rows * w->mlp * 2, (size_t)2 * rows * W * 2);
    L->planes_b = o;   o += align_up(pb);
    L->patch_f32 = o;  o += align_up((size_t)B * g * g * W * 4);
    L->head_f32 = o;   o += align_up((size_t)B * W * 4);  // attention-pool head: pooled vector before its MLP
    L->total = o;
    return true;
}

// `layers` pre-LN residual blocks on x (rows = B*T, fp32 residual stream): x += attn(ln_1(x)); x += mlp(ln_2(x)).
// Shared by the image towers (causal = 0) and the CLIP text tower (causal = 1).
int run_blocks(const SlbVitLayer* layer, int n_layers, float* x, float* qkv, uint16_t* pa, uint16_t* pb, int64_t B, int64_t T,
               int64_t W, int heads, int64_t mlp, int act, int fmt, float eps, int causal, void* stream) {
    const int64_t rows = B * T, dh = W / heads;
    int rc;
#define SLB_TRY(call)            \
    do {                         \
        rc = (call);             \
        if (rc != SLB_OK) return rc; \
    } while (0)
    for (int l = 0; l < n_layers; ++l) {
        const SlbVitLayer& ly = layer[l];
        SLB_TRY(slb_layernorm(x, rows, W, W, ly.ln1_g, ly.ln1_b, eps, fmt, nullptr, pa, stream));
        if (fmt == SLB_PLANE_F16 && dh == 64) {
            // in_proj writes q | k | v as split planes (same bytes as fp32) and attention consumes them directly
            uint16_t* qkv_planes = reinterpret_cast<uint16_t*>(qkv);
            SLB_TRY(slb_gemm_split(pa, ly.w_qkv, fmt, rows, 3 * W, W, kAlpha, ly.b_qkv, nullptr, nullptr, nullptr, SLB_EPI_NONE, 3,
                                   nullptr, qkv_planes, stream));
            SLB_TRY(slb_attention_planes(qkv_planes, B, T, heads, dh, 1.0f / sqrtf((float)dh), causal, fmt, nullptr, pb, stream));
        } else {
            SLB_REQUIRE(!causal, SLB_EUNSUPPORTED, "causal attention needs head_dim 64 and fp16 planes");
            SLB_TRY(slb_gemm_split(pa, ly.w_qkv, fmt, rows, 3 * W, W, kAlpha, ly.b_qkv, nullptr, nullptr, nullptr, SLB_EPI_NONE, 3,
                                   qkv, nullptr, stream));
            SLB_TRY(slb_attention_small(qkv, T * 3 * W, 3 * W, qkv + W, qkv + 2 * W, T * 3 * W, 3 * W, B, T, T, heads, dh,
                                        1.0f / sqrtf((float)dh), fmt, nullptr, pb, stream));
        }
        SLB_TRY(slb_gemm_split(pb, ly.w_out, fmt, rows, W, W, kAlpha, ly.b_out, x, nullptr, nullptr, SLB_EPI_NONE, 3, x, nullptr,
                               stream));
        SLB_TRY(slb_layernorm(x, rows, W, W, ly.ln2_g, ly.ln2_b, eps, fmt, nullptr, pa, stream));
        SLB_TRY(slb_gemm_split(pa, ly.w_fc, fmt, rows, mlp, W, kAlpha, ly.b_fc, nullptr, nullptr, nullptr, act, 3, nullptr, pb, stream));
        SLB_TRY(slb_gemm_split(pb, ly.w_proj, fmt, rows, W, mlp, kAlpha, ly.b_proj, x, nullptr, nullptr, SLB_EPI_NONE, 3, x, nullptr,
                               stream));
    }
#undef SLB_TRY
    return SLB_OK;
}

}  // namespace

extern "C" size_t slb_vit_workspace_bytes(const SlbVitWeights* w, int64_t B) {
    WsLayout L;
    if (B < 0 || !layout_for(w, B, &L)) return 0;
    return L.total;
}

extern "C" int slb_vit_forward(const SlbVitWeights* w, const float* img, int64_t B, float* out, void* workspace,
                               size_t workspace_bytes, void* stream) {
    SLB_REQUIRE(w != nullptr && B >= 0, SLB_EINVAL, "slb_vit_forward: bad arguments");
    if (B == 0) return SLB_OK;
    SLB_REQUIRE(img && out && workspace, SLB_EINVAL, "slb_vit_forward: null pointer");
    SLB_REQUIRE(w->layers >= 0 && (w->layer || w->layers == 0) && w->conv_w && w->pos && w->ln_post_g, SLB_EINVAL,
                "slb_vit_forward: incomplete weights");
    SLB_REQUIRE(w->width % 64 == 0 && w->mlp % 64 == 0 && w->heads > 0 && w->width % w->heads == 0, SLB_EUNSUPPORTED,
                "slb_vit_forward: width and mlp must be multiples of 64");
    SLB_REQUIRE(w->patch % 2 == 0, SLB_EUNSUPPORTED, "slb_vit_forward: patch size must be even");
    SLB_REQUIRE((w->pool == SLB_POOL_CLS && w->has_cls) || w->pool == SLB_POOL_MAP, SLB_EUNSUPPORTED,
                "slb_vit_forward: pooling must be SLB_POOL_CLS (with a class token) or SLB_POOL_MAP");
    SLB_REQUIRE(!w->has_cls || w->cls, SLB_EINVAL, "slb_vit_forward: class token missing");
    if (w->pool == SLB_POOL_MAP)
        SLB_REQUIRE(w->map_q && w->map_w_kv && w->map_w_out && w->map_ln_g && w->map_w_fc && w->map_w_proj, SLB_EINVAL,
                    "slb_vit_forward: incomplete attention-pool head");
    WsLayout L;
    SLB_REQUIRE(layout_for(w, B, &L), SLB_EINVAL, "slb_vit_forward: image_size must be a multiple of patch");
    SLB_REQUIRE(((uintptr_t)workspace % 256) == 0, SLB_EINVAL, "slb_vit_forward: workspace must be 256-byte aligned");
    SLB_REQUIRE(workspace_bytes >= L.total, SLB_EWORKSPACE, "slb_vit_forward: workspace needs %zu bytes, got %zu", L.total,
                workspace_bytes);

    unsigned char* ws = static_cast<unsigned char*>(workspace);
    float* x = reinterpret_cast<float*>(ws + L.x);
    float* qkv = reinterpret_cast<float*>(ws + L.qkv);
    uint16_t* pa = reinterpret_cast<uint16_t*>(ws + L.planes_a);
    uint16_t* pb = reinterpret_cast<uint16_t*>(ws + L.planes_b);
    float* patch_f32 = reinterpret_cast<float*>(ws + L.patch_f32);
    float* head_f32 = reinterpret_cast<float*>(ws + L.head_f32);

    const int fmt = w->plane_fmt;
    const int64_t g = w->image_size / w->patch;
    const int64_t T = g * g + (w->has_cls ? 1 : 0), W = w->width, rows = B * T, dh = W / w->heads;
    const int64_t Kc = slb_patch_k(w->patch);  // conv_w planes are [2, width, Kc], zero padded past 3*P*P
    int rc;
#define SLB_TRY(call)            \
    do {                         \
        rc = (call);             \
        if (rc != SLB_OK) return rc; \
    } while (0)

    // patch embedding: im2col planes -> GEMM (+ conv bias if any) -> tokens
    SLB_TRY(slb_patchify(img, B, w->image_size, w->patch, fmt, pa, stream));
    SLB_TRY(slb_gemm_split(pa, w->conv_w, fmt, B * g * g, W, Kc, kAlpha, w->conv_b, nullptr, nullptr, nullptr, SLB_EPI_NONE, 3,
                           patch_f32, nullptr, stream));
    SLB_TRY(slb_assemble_tokens(patch_f32, w->cls, w->pos, B, T, W, w->has_cls ? 1 : 0, x, stream));
    if (w->ln_pre_g) SLB_TRY(slb_layernorm(x, rows, W, W, w->ln_pre_g, w->ln_pre_b, w->ln_eps, fmt, x, nullptr, stream));

    SLB_TRY(run_blocks(w->layer, w->layers, x, qkv, pa, pb, B, T, W, w->heads, w->mlp, w->act, fmt, w->ln_eps, 0, stream));

    if (w->pool == SLB_POOL_MAP) {
        // final LayerNorm over ALL tokens, then the attention-pool head (timm AttentionPoolLatent / HF Siglip "MAP" head)
        SLB_TRY(slb_layernorm(x, rows, W, W, w->ln_post_g, w->ln_post_b, w->ln_eps, fmt, nullptr, pa, stream));
        float* kv = qkv;  // [rows, 2W] fp32
        SLB_TRY(slb_gemm_split(pa, w->map_w_kv, fmt, rows, 2 * W, W, kAlpha, w->map_b_kv, nullptr, nullptr, nullptr, SLB_EPI_NONE, 3, kv,
                               nullptr, stream));
        // one query (the projected latent, shared by every image: batch stride 0) over the T tokens of each image
        SLB_TRY(slb_attention_small(w->map_q, 0, W, kv, kv + W, T * 2 * W, 2 * W, B, 1, T, w->heads, dh,
                                    1.0f / sqrtf((float)dh), fmt, nullptr, pb, stream));
        SLB_TRY(slb_gemm_split(pb, w->map_w_out, fmt, B, W, W, kAlpha, w->map_b_out, nullptr, nullptr, nullptr, SLB_EPI_NONE, 3,
                               head_f32, nullptr, stream));
        SLB_TRY(slb_layernorm(head_f32, B, W, W, w->map_ln_g, w->map_ln_b, w->ln_eps, fmt, nullptr, pa, stream));
        SLB_TRY(slb_gemm_split(pa, w->map_w_fc, fmt, B, w->mlp, W, kAlpha, w->map_b_fc, nullptr, nullptr, nullptr, w->act, 3, nullptr, pb,
                               stream));
        if (w->proj) {
            SLB_TRY(slb_gemm_split(pb, w->map_w_proj, fmt, B, W, w->mlp, kAlpha, w->map_b_proj, head_f32, nullptr, nullptr, SLB_EPI_NONE,
                                   3, nullptr, pa, stream));
            SLB_TRY(slb_gemm_split(pa, w->proj, fmt, B, w->embed_dim, W, kAlpha, nullptr, nullptr, nullptr, nullptr, SLB_EPI_NONE, 3, out,
                                   nullptr, stream));
        } else {
            SLB_REQUIRE(w->embed_dim == w->width, SLB_EINVAL, "slb_vit_forward: embed_dim must equal width without a projection");
            SLB_TRY(slb_gemm_split(pb, w->map_w_proj, fmt, B, W, w->mlp, kAlpha, w->map_b_proj, head_f32, nullptr, nullptr, SLB_EPI_NONE,
                                   3, out, nullptr, stream));
        }
    } else if (w->proj) {
        // ln_post on the class tokens (rows T*W apart), then the projection
        SLB_TRY(slb_layernorm(x, B, W, T * W, w->ln_post_g, w->ln_post_b, w->ln_eps, fmt, nullptr, pa, stream));
        SLB_TRY(slb_gemm_split(pa, w->proj, fmt, B, w->embed_dim, W, kAlpha, nullptr, nullptr, nullptr, nullptr, SLB_EPI_NONE, 3, out,
                               nullptr, stream));
    } else {
        SLB_TRY(slb_layernorm(x, B, W, T * W, w->ln_post_g, w->ln_post_b, w->ln_eps, fmt, out, nullptr, stream));
    }
#undef SLB_TRY
    return SLB_OK;
}

// The trunk of a ViT in slices, for a caller that wants the residual stream between blocks (the accelerated probed
// forward, semanticlens_b200/probed.py: torchvision's VisionTransformer under forward hooks on its encoder blocks —
// BASELINE configs[2]). layer_begin == 0 first embeds the images (patch GEMM, class token, positions, ln_pre); then blocks
// [layer_begin, layer_end) run. The residual stream [B*T, W] fp32 is the FIRST region of the workspace, valid after the call.
extern "C" int slb_vit_trunk(const SlbVitWeights* w, const float* img, int64_t B, int32_t layer_begin, int32_t layer_end,
                             void* workspace, size_t workspace_bytes, void* stream) {
    SLB_REQUIRE(w != nullptr && B >= 0, SLB_EINVAL, "slb_vit_trunk: bad arguments");
    if (B == 0) return SLB_OK;
    SLB_REQUIRE(workspace && (img || layer_begin > 0), SLB_EINVAL, "slb_vit_trunk: null pointer");
    SLB_REQUIRE(0 <= layer_begin && layer_begin <= layer_end && layer_end <= w->layers && (w->layer || w->layers == 0), SLB_EINVAL,
                "slb_vit_trunk: bad layer range [%d, %d) of %d", layer_begin, layer_end, w->layers);
    SLB_REQUIRE(w->conv_w && w->pos && (!w->has_cls || w->cls), SLB_EINVAL, "slb_vit_trunk: incomplete weights");
    SLB_REQUIRE(w->width % 64 == 0 && w->mlp % 64 == 0 && w->heads > 0 && w->width % w->heads == 0 && w->patch % 2 == 0,
                SLB_EUNSUPPORTED, "slb_vit_trunk: width and mlp must be multiples of 64, the patch size even");
    WsLayout L;
    SLB_REQUIRE(layout_for(w, B, &L), SLB_EINVAL, "slb_vit_trunk: image_size must be a multiple of patch");
    SLB_REQUIRE(L.x == 0, SLB_ECUDA, "slb_vit_trunk: the residual stream must lead the workspace");
    SLB_REQUIRE(((uintptr_t)workspace % 256) == 0, SLB_EINVAL, "slb_vit_trunk: workspace must be 256-byte aligned");
    SLB_REQUIRE(workspace_bytes >= L.total, SLB_EWORKSPACE, "slb_vit_trunk: workspace needs %zu bytes, got %zu", L.total, workspace_bytes);
    unsigned char* ws = static_cast<unsigned char*>(workspace);
    float* x = reinterpret_cast<float*>(ws + L.x);
    float* qkv = reinterpret_cast<float*>(ws + L.qkv);
    uint16_t* pa = reinterpret_cast<uint16_t*>(ws + L.planes_a);
    uint16_t* pb = reinterpret_cast<uint16_t*>(ws + L.planes_b);
    float* patch_f32 = reinterpret_cast<float*>(ws + L.patch_f32);
    const int fmt = w->plane_fmt;
    const int64_t g = w->image_size / w->patch;
    const int64_t T = g * g + (w->has_cls ? 1 : 0), W = w->width, rows = B * T;
    int rc;
    if (layer_begin == 0) {
        rc = slb_patchify(img, B, w->image_size, w->patch, fmt, pa, stream);
        if (rc != SLB_OK) return rc;
        rc = slb_gemm_split(pa, w->conv_w, fmt, B * g * g, W, slb_patch_k(w->patch), kAlpha, w->conv_b, nullptr, nullptr, nullptr,
                            SLB_EPI_NONE, 3, patch_f32, nullptr, stream);
        if (rc != SLB_OK) return rc;
        rc = slb_assemble_tokens(patch_f32, w->cls, w->pos, B, T, W, w->has_cls ? 1 : 0, x, stream);
        if (rc != SLB_OK) return rc;
        if (w->ln_pre_g) {
            rc = slb_layernorm(x, rows, W, W, w->ln_pre_g, w->ln_pre_b, w->ln_eps, fmt, x, nullptr, stream);
            if (rc != SLB_OK) return rc;
        }
    }
    return run_blocks(w->layer + layer_begin, layer_end - layer_begin, x, qkv, pa, pb, B, T, W, w->heads, w->mlp, w->act, fmt, w->ln_eps,
                      0, stream);
}

// ---------------------------------------------------------------------------------------------
// CLIP text tower
// ---------------------------------------------------------------------------------------------
namespace {
struct TextLayout {
    size_t x, qkv, planes_a, planes_b, emb, head, total;
};
bool text_layout(const SlbTextWeights* w, int64_t B, TextLayout* L) {
    if (!w || w->context <= 0 || w->width <= 0) return false;
    const int64_t rows = B * w->context, W = w->width;
    size_t o = 0;
    L->x = o;         o += align_up((size_t)rows * W * 4);
    L->qkv = o;       o += align_up((size_t)rows * 3 * W * 4);
    L->planes_a = o;  o += align_up((size_t)2 * rows * W * 2);
    L->planes_b = o;  o += align_up(std::max<size_t>((size_t)2 * rows * w->mlp * 2, (size_t)2 * rows * W * 2));
    L->emb = o;       o += align_up((size_t)rows * W * 4);
    L->head = o;      o += align_up((size_t)B * W * 4);
    L->total = o;
    return true;
}
}  // namespace

extern "C" size_t slb_text_workspace_bytes(const SlbTextWeights* w, int64_t B) {
    TextLayout L;
    if (B < 0 || !text_layout(w, B, &L)) return 0;
    return L.total;
}

extern "C" int slb_text_forward(const SlbTextWeights* w, const int64_t* tokens, const int64_t* eot_rows, int64_t B, float* out,
                                void* workspace, size_t workspace_bytes, void* stream) {
    SLB_REQUIRE(w != nullptr && B >= 0, SLB_EINVAL, "slb_text_forward: bad arguments");
    if (B == 0) return SLB_OK;
    SLB_REQUIRE(tokens && eot_rows && out && workspace, SLB_EINVAL, "slb_text_forward: null pointer");
    SLB_REQUIRE(w->tok_emb && w->pos && w->ln_final_g && w->proj && (w->layer || w->layers == 0), SLB_EINVAL,
                "slb_text_forward: incomplete weights");
    SLB_REQUIRE(w->width % 64 == 0 && w->mlp % 64 == 0 && w->heads > 0 && w->width == 64 * w->heads, SLB_EUNSUPPORTED,
                "slb_text_forward: width must be 64 * heads");
    SLB_REQUIRE(w->plane_fmt == SLB_PLANE_F16, SLB_EUNSUPPORTED, "slb_text_forward: fp16 planes only");
    TextLayout L;
    SLB_REQUIRE(text_layout(w, B, &L), SLB_EINVAL, "slb_text_forward: bad configuration");
    SLB_REQUIRE(((uintptr_t)workspace % 256) == 0, SLB_EINVAL, "slb_text_forward: workspace must be 256-byte aligned");
    SLB_REQUIRE(workspace_bytes >= L.total, SLB_EWORKSPACE, "slb_text_forward: workspace needs %zu bytes, got %zu", L.total,
                workspace_bytes);
    unsigned char* ws = static_cast<unsigned char*>(workspace);
    float* x = reinterpret_cast<float*>(ws + L.x);
    float* qkv = reinterpret_cast<float*>(ws + L.qkv);
    uint16_t* pa = reinterpret_cast<uint16_t*>(ws + L.planes_a);
    uint16_t* pb = reinterpret_cast<uint16_t*>(ws + L.planes_b);
    float* emb = reinterpret_cast<float*>(ws + L.emb);
    float* head = reinterpret_cast<float*>(ws + L.head);
    const int64_t T = w->context, W = w->width;
    int rc;
    // x = token_embedding[tokens] + positional_embedding
    rc = slb_gather_rows(w->tok_emb, w->vocab, W, tokens, B * T, emb, stream);
    if (rc != SLB_OK) return rc;
    rc = slb_assemble_tokens(emb, nullptr, w->pos, B, T, W, 0, x, stream);
    if (rc != SLB_OK) return rc;
    rc = run_blocks(w->layer, w->layers, x, qkv, pa, pb, B, T, W, w->heads, w->mlp, w->act, w->plane_fmt, w->ln_eps,
                    w->non_causal ? 0 : 1, stream);
    if (rc != SLB_OK) return rc;
    // ln_final acts per token, so it commutes with picking the end-of-text rows: gather first, normalise B rows
    rc = slb_gather_rows(x, B * T, W, eot_rows, B, head, stream);
    if (rc != SLB_OK) return rc;
    rc = slb_layernorm(head, B, W, W, w->ln_final_g, w->ln_final_b, w->ln_eps, w->plane_fmt, nullptr, pa, stream);
    if (rc != SLB_OK) return rc;
    return slb_gemm_split(pa, w->proj, w->plane_fmt, B, w->embed_dim, W, kAlpha, w->proj_b, nullptr, nullptr, nullptr, SLB_EPI_NONE, 3,
                          out, nullptr, stream);
}
