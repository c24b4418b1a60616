// C ABI (include/opadpo_b200.h) — operator-level entry points.
#include <cstdio>
#include <cstring>

#include "../../include/opadpo_b200.h"
#include "common.h"
#include "attention.h"
#include "collate.h"
#include "decode_gemm.h"
#include "gemm.h"
#include "elementwise.h"
#include "engine.h"

namespace opadpo {
thread_local char g_last_error[1024] = {0};
std::atomic<int64_t> g_launches{0};
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
  va_end(ap);
}
}  // namespace opadpo

using namespace opadpo;

extern "C" {

const char* opadpo_last_error(void) { return g_last_error; }
int opadpo_abi_version(void) { return 7; }
int64_t opadpo_launch_count(void) { return g_launches.load(); }

int opadpo_gemm_n_tiles(int32_t N) { return gemm_num_n_tiles(N); }

int opadpo_op_gemm(const opadpo_gemm_desc* d, void* stream) {
  if (!d) {
    set_error("opadpo_op_gemm: null descriptor");
    return OPADPO_EINVAL;
  }
  GemmArgs a;
  a.M = d->M;
  a.N = d->N;
  a.K1 = d->K1;
  a.K2 = d->K2;
  a.a_mn_major = d->a_mn_major != 0;
  a.b_mn_major = d->b_mn_major != 0;
  a.A1 = {static_cast<const __nv_bfloat16*>(d->A1), d->lda1};
  a.B1 = {static_cast<const __nv_bfloat16*>(d->B1), d->ldb1};
  a.A2 = {static_cast<const __nv_bfloat16*>(d->A2), d->lda2};
  a.B2 = {static_cast<const __nv_bfloat16*>(d->B2), d->ldb2};
  a.a2_group_n = d->a2_group_n;
  a.a2_group_stride = d->a2_group_stride;
  a.splits = d->splits;
  a.epilogue = d->epilogue;
  a.alpha = d->alpha;
  a.C = d->C;
  a.ldc = d->ldc;
  a.bias = static_cast<const __nv_bfloat16*>(d->bias);
  a.residual = static_cast<const __nv_bfloat16*>(d->residual);
  a.ldr = d->ldr;
  a.act = d->act;
  a.labels = d->labels;
  a.lse_partials = static_cast<float4*>(d->lse_partials);
  a.label_logit = d->label_logit;
  a.lse = d->lse;
  a.dlp = d->dlp;
  a.dent = d->dent;
  a.ent = d->ent;
  a.rope_pos = d->rope_pos;
  a.rope_cs = static_cast<const float2*>(d->rope_cs);
  a.rope_cols = d->rope_cols;
  a.rope_S = d->rope_S;
  a.max_ctas = d->max_ctas;
  a.cluster = d->cluster_mode - 1;  /* 0 (zero-initialised desc) = auto, 1 off, 2 B-multicast cluster, 3 CTA pair */
  int rc = gemm_launch(a, static_cast<cudaStream_t>(stream));
  if (rc != 0) {
    set_error("%s", gemm_last_error());
    return rc <= -10 ? OPADPO_ECUDA : OPADPO_EINVAL;
  }
  return OPADPO_OK;
}


int opadpo_op_decode_gemm(const void* x, int64_t ldx, const void* W, int64_t ldw, int32_t Mb, int32_t N, int32_t K,
                          int32_t splits, int32_t max_splits, float alpha, float* part, int32_t* splits_out, void* stream) {
  if (!x || !W || !part) {
    set_error("opadpo_op_decode_gemm: null pointer");
    return OPADPO_EINVAL;
  }
  DecodeGemmArgs a;
  a.Mb = Mb;
  a.N = N;
  a.K = K;
  a.x = static_cast<const __nv_bfloat16*>(x);
  a.ldx = ldx;
  a.W = static_cast<const __nv_bfloat16*>(W);
  a.ldw = ldw;
  a.part = part;
  a.ldc = N;
  a.alpha = alpha;
  a.splits = splits > 0 ? splits : decode_gemm_pick_splits(Mb, N, K, max_splits);
  if (splits_out) *splits_out = a.splits;
  int rc = decode_gemm_launch(a, static_cast<cudaStream_t>(stream));
  if (rc != 0) {
    set_error("%s", decode_gemm_last_error());
    return rc <= -10 ? OPADPO_ECUDA : OPADPO_EINVAL;
  }
  return OPADPO_OK;
}

static AttnArgs to_attn(const opadpo_attn_desc* d) {
  AttnArgs a;
  a.N = d->N; a.S = d->S; a.H = d->H; a.HD = d->HD;
  a.causal = d->causal != 0;
  a.scale = d->scale;
  a.q = static_cast<const __nv_bfloat16*>(d->q);
  a.k = static_cast<const __nv_bfloat16*>(d->k);
  a.v = static_cast<const __nv_bfloat16*>(d->v);
  a.ld_qkv = d->ld_qkv;
  a.o = static_cast<__nv_bfloat16*>(d->o);
  a.ld_o = d->ld_o;
  a.lse = d->lse;
  a.kbits = d->kbits;
  a.prefix_len = d->prefix_len;
  a.segment_len = d->segment_len;
  a.dO = static_cast<const __nv_bfloat16*>(d->dO);
  a.ld_do = d->ld_do;
  a.delta = d->delta;
  a.dq = static_cast<__nv_bfloat16*>(d->dq);
  a.dk = static_cast<__nv_bfloat16*>(d->dk);
  a.dv = static_cast<__nv_bfloat16*>(d->dv);
  a.ld_dqkv = d->ld_dqkv;
  return a;
}

int opadpo_op_attn_fwd(const opadpo_attn_desc* d, void* stream) {
  if (!d) { set_error("opadpo_op_attn_fwd: null descriptor"); return OPADPO_EINVAL; }
  int rc = attn_forward(to_attn(d), static_cast<cudaStream_t>(stream));
  if (rc != 0) { set_error("%s", attn_last_error()); return rc <= -10 ? OPADPO_ECUDA : OPADPO_EINVAL; }
  return OPADPO_OK;
}

int opadpo_op_attn_bwd(const opadpo_attn_desc* d, void* stream) {
  if (!d) { set_error("opadpo_op_attn_bwd: null descriptor"); return OPADPO_EINVAL; }
  int rc = attn_backward(to_attn(d), static_cast<cudaStream_t>(stream));
  if (rc != 0) { set_error("%s", attn_last_error()); return rc <= -10 ? OPADPO_ECUDA : OPADPO_EINVAL; }
  return OPADPO_OK;
}

int opadpo_op_pack_mask(const uint8_t* mask, uint64_t* bits, int32_t N, int32_t S, void* stream) {
  int rc = attn_pack_mask(mask, bits, N, S, static_cast<cudaStream_t>(stream));
  if (rc != 0) { set_error("pack_mask launch failed"); return OPADPO_ECUDA; }
  return OPADPO_OK;
}

int opadpo_prof_begin(void) {
  gemm_prof_begin();
  return OPADPO_OK;
}
int opadpo_prof_end(double* gemm_ms, double* gemm_flops, int64_t* gemm_launches) {
  if (!gemm_ms || !gemm_flops || !gemm_launches) return OPADPO_EINVAL;
  return gemm_prof_end(gemm_ms, gemm_flops, gemm_launches);
}

int opadpo_op_dpo_pair(const opadpo_dpo_pair_desc* d, void* stream) {
  g_last_error[0] = 0;
  if (!d || !d->pi_c || !d->pi_r || !d->ref_c || !d->ref_r || !d->out || d->n <= 0) {
    set_error("opadpo_op_dpo_pair: null pointer or n <= 0");
    return OPADPO_EINVAL;
  }
  DpoPairArgs a;
  a.pi_c = d->pi_c; a.pi_r = d->pi_r; a.ref_c = d->ref_c; a.ref_r = d->ref_r;
  a.s_c = d->s_c; a.s_r = d->s_r; a.d_c = d->d_c; a.d_r = d->d_r; a.out = d->out; a.n = d->n;
  a.beta = d->beta; a.label_smoothing = d->label_smoothing; a.reference_free = d->reference_free;
  a.divergence = d->divergence; a.alpha_div = d->alpha_div; a.loss_coef = d->loss_coef;
  a.anchor_mode = d->anchor_mode; a.anchor_sign = d->anchor_sign; a.anchor_value = d->anchor_value;
  a.anchor_coef = d->anchor_coef; a.grad_scale = d->grad_scale;
  int rc = dpo_pair_loss(a, static_cast<cudaStream_t>(stream));
  if (rc != 0) { set_error("dpo_pair_loss launch failed"); return OPADPO_ECUDA; }
  return OPADPO_OK;
}

// ---------------------------------------------------------------------------------------------- collator stage
int opadpo_op_collate_responses(const opadpo_segments* s, int32_t R, int32_t pad_id, int32_t eos_id, int32_t mode,
                                int64_t* ids, uint8_t* attn_mask, float* scores, float* relations, int32_t* overflow,
                                void* stream) {
  g_last_error[0] = 0;
  if (!s || !ids || s->b < 0 || R <= 0 || (s->b > 0 && (!s->seg_offsets || !s->row_offsets))) {  // seg_* may be NULL when no row has an entry
    set_error("opadpo_op_collate_responses: null pointer or bad sizes");
    return OPADPO_EINVAL;
  }
  collate::Segments g{s->tokens, s->seg_offsets, s->seg_score, s->seg_relation, s->seg_flags, s->row_offsets};
  return collate_responses(g, s->b, R, pad_id, eos_id, mode, ids, attn_mask, scores, relations, overflow,
                           static_cast<cudaStream_t>(stream));
}
int opadpo_op_collate_queries(const int32_t* tokens, const int64_t* offsets, int32_t b, int32_t Q, int32_t pad_id,
                              int32_t placeholder_id, int32_t image_token, int64_t* ids, uint8_t* attn_mask, void* stream) {
  g_last_error[0] = 0;
  if (!offsets || !ids || b < 0 || Q <= 0) {
    set_error("opadpo_op_collate_queries: null pointer or bad sizes");
    return OPADPO_EINVAL;
  }
  return collate_queries(tokens, offsets, b, Q, pad_id, placeholder_id, image_token, ids, attn_mask,
                         static_cast<cudaStream_t>(stream));
}
int opadpo_op_mask_images(const float* images, float* out, int32_t n, int32_t C, int32_t H, int32_t W, const int64_t* idx,
                          int32_t m, int32_t block, int32_t bf16_pixels, float* means_out, void* stream) {
  g_last_error[0] = 0;
  if (!images || !out || n < 0 || C <= 0 || H <= 0 || W <= 0 || m < 0 || (m > 0 && !idx) || block < 1 || H % block || W % block) {
    set_error("opadpo_op_mask_images: null pointer or bad sizes (block must divide H and W)");
    return OPADPO_EINVAL;
  }
  return mask_images(images, out, n, C, H, W, idx, m, block, bf16_pixels, means_out, static_cast<cudaStream_t>(stream));
}
int opadpo_op_truncate_after_eos(const int64_t* ids, int64_t* out, int32_t b, int32_t R, int64_t eos_id, int64_t pad_id,
                                 const int64_t* extra, int32_t n_extra, void* stream) {
  g_last_error[0] = 0;
  if (!ids || !out || b < 0 || R <= 0 || n_extra < 0 || (n_extra > 0 && !extra)) {
    set_error("opadpo_op_truncate_after_eos: null pointer or bad sizes");
    return OPADPO_EINVAL;
  }
  return truncate_after_eos(ids, out, b, R, eos_id, pad_id, extra, n_extra, static_cast<cudaStream_t>(stream));
}
int opadpo_op_attention_holes(uint8_t* out, const int64_t* idx, int32_t n, int32_t P, int32_t m, void* stream) {
  g_last_error[0] = 0;
  if (!out || n < 0 || P <= 0 || m < 0 || (m > 0 && !idx)) {
    set_error("opadpo_op_attention_holes: null pointer or bad sizes");
    return OPADPO_EINVAL;
  }
  return attention_holes(out, idx, n, P, m, static_cast<cudaStream_t>(stream));
}
int opadpo_op_preprocess_images(const uint8_t* pixels, const int64_t* h_offsets, const int32_t* h_heights,
                                const int32_t* h_widths, int32_t n, int32_t S, const float* mean3, const float* std3,
                                const uint8_t* bg3, int32_t pad_to_square, float* out, void* stream) {
  g_last_error[0] = 0;
  if (!pixels || !h_offsets || !h_heights || !h_widths || !mean3 || !std3 || !out || n < 0 || S <= 0) {
    set_error("opadpo_op_preprocess_images: null pointer or bad sizes");
    return OPADPO_EINVAL;
  }
  return preprocess_images(pixels, h_offsets, h_heights, h_widths, n, S, mean3, std3, bg3, pad_to_square, out,
                           static_cast<cudaStream_t>(stream));
}

// ---------------------------------------------------------------------------------------------- engine level
struct opadpo_engine {
  Engine* impl;
};

#define ENGINE_GUARD(e)                                   \
  g_last_error[0] = 0;                                    \
  if (!(e) || !(e)->impl) {                               \
    set_error("%s: null engine", __func__);               \
    return OPADPO_EINVAL;                                 \
  }

int opadpo_engine_create(const opadpo_config* cfg, opadpo_engine** out) {
  g_last_error[0] = 0;
  if (!cfg || !out) { set_error("opadpo_engine_create: null argument"); return OPADPO_EINVAL; }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    set_error("opadpo_engine_create: no CUDA device (the engine has no CPU path)");
    return OPADPO_ECUDA;
  }
  Engine* impl = new Engine(*cfg);
  int rc = impl->init();
  if (rc != 0) { delete impl; return rc; }
  *out = new opadpo_engine{impl};
  return OPADPO_OK;
}
int opadpo_engine_destroy(opadpo_engine* e) {
  if (!e) return OPADPO_OK;
  delete e->impl;
  delete e;
  return OPADPO_OK;
}
int opadpo_engine_load_tensor(opadpo_engine* e, const char* name, const void* dev_ptr, int32_t dtype, int64_t rows,
                              int64_t cols, void* stream) {
  ENGINE_GUARD(e);
  return e->impl->load_tensor(name, dev_ptr, dtype, rows, cols, static_cast<cudaStream_t>(stream));
}
int opadpo_engine_load_lora(opadpo_engine* e, int32_t adapter, const char* module, const void* A, const void* B,
                            int32_t dtype, void* stream) {
  ENGINE_GUARD(e);
  return e->impl->load_lora(adapter, module, A, B, dtype, static_cast<cudaStream_t>(stream));
}
int opadpo_engine_export_lora(opadpo_engine* e, int32_t adapter, const char* module, float* A_out, float* B_out,
                              void* stream) {
  ENGINE_GUARD(e);
  return e->impl->export_lora(adapter, module, A_out, B_out, static_cast<cudaStream_t>(stream));
}
int opadpo_engine_lora_shape(opadpo_engine* e, const char* module, int64_t* in_features, int64_t* out_features) {
  ENGINE_GUARD(e);
  return e->impl->lora_shape(module, in_features, out_features);
}
int opadpo_engine_set_trainable(opadpo_engine* e, int32_t adapter) {
  ENGINE_GUARD(e);
  return e->impl->set_trainable(adapter);
}
int opadpo_engine_buffers(opadpo_engine* e, float** master, float** grad, float** exp_avg, float** exp_avg_sq,
                          int64_t* n_total, int64_t* n_train) {
  ENGINE_GUARD(e);
  return e->impl->buffers(master, grad, exp_avg, exp_avg_sq, n_total, n_train);
}
int opadpo_engine_lora_offsets(opadpo_engine* e, const char* module, int64_t* offA, int64_t* offB) {
  ENGINE_GUARD(e);
  return e->impl->lora_offsets(module, offA, offB);
}
int opadpo_engine_set_option(opadpo_engine* e, const char* name, int64_t value) {
  ENGINE_GUARD(e);
  if (!name) { set_error("opadpo_engine_set_option: null name"); return OPADPO_EINVAL; }
  return e->impl->set_option(name, value);
}
int opadpo_engine_set_backward_callback(opadpo_engine* e, opadpo_layer_cb cb, void* user) {
  ENGINE_GUARD(e);
  return e->impl->set_backward_callback(cb, user);
}
int opadpo_engine_adopt_grad_buffer(opadpo_engine* e, float* buffer, int64_t count) {
  ENGINE_GUARD(e);
  return e->impl->adopt_grad_buffer(buffer, count);
}
int opadpo_engine_layer_grad_range(opadpo_engine* e, int32_t layer, int64_t* offset, int64_t* count) {
  ENGINE_GUARD(e);
  if (!offset || !count) { set_error("opadpo_engine_layer_grad_range: null pointer"); return OPADPO_EINVAL; }
  return e->impl->layer_grad_range(layer, offset, count);
}
int opadpo_engine_sync_shadow(opadpo_engine* e, void* stream) {
  ENGINE_GUARD(e);
  return e->impl->sync_shadow(static_cast<cudaStream_t>(stream));
}
int opadpo_forward_logprobs(opadpo_engine* e, int32_t adapter, const float* images, int32_t b, const int64_t* queries,
                            const uint8_t* qmask, const uint8_t* img_mask, int32_t Q, int32_t K,
                            const int64_t* const* responses, int32_t R, float temperature, float* out_logprobs,
                            float* out_entropies, int32_t save_slot, void* stream) {
  ENGINE_GUARD(e);
  if (!images || !queries || !qmask || !responses || !out_logprobs) {
    set_error("opadpo_forward_logprobs: null pointer");
    return OPADPO_EINVAL;
  }
  return e->impl->forward_logprobs(adapter, images, b, queries, qmask, img_mask, Q, K, responses, R, temperature,
                                   out_logprobs, out_entropies, save_slot, static_cast<cudaStream_t>(stream));
}
int opadpo_backward(opadpo_engine* e, int32_t save_slot, const float* dlogprobs, void* stream) {
  ENGINE_GUARD(e);
  if (!dlogprobs) { set_error("opadpo_backward: null dlogprobs"); return OPADPO_EINVAL; }
  return e->impl->backward(save_slot, dlogprobs, nullptr, static_cast<cudaStream_t>(stream));
}
int opadpo_backward_ent(opadpo_engine* e, int32_t save_slot, const float* dlogprobs, const float* dentropies, void* stream) {
  ENGINE_GUARD(e);
  if (!dlogprobs && !dentropies) { set_error("opadpo_backward_ent: both gradients are null"); return OPADPO_EINVAL; }
  return e->impl->backward(save_slot, dlogprobs, dentropies, static_cast<cudaStream_t>(stream));
}
int opadpo_zero_grad(opadpo_engine* e, void* stream) {
  ENGINE_GUARD(e);
  return e->impl->zero_grad(static_cast<cudaStream_t>(stream));
}
int opadpo_optimizer_step(opadpo_engine* e, float lr, float beta1, float beta2, float eps, float weight_decay,
                          float max_grad_norm, float grad_scale, float* grad_norm_out, void* stream) {
  ENGINE_GUARD(e);
  return e->impl->optimizer_step(lr, beta1, beta2, eps, weight_decay, max_grad_norm, grad_scale, grad_norm_out,
                                 static_cast<cudaStream_t>(stream));
}
int opadpo_generate(opadpo_engine* e, int32_t adapter, const float* images, int32_t b, const int64_t* queries,
                    const uint8_t* qmask, int32_t Q, int32_t max_new, const opadpo_sampler* sampler, int64_t* out_ids,
                    void* stream) {
  ENGINE_GUARD(e);
  if (!images || !queries || !qmask || !out_ids) { set_error("opadpo_generate: null pointer"); return OPADPO_EINVAL; }
  return e->impl->generate(adapter, images, b, queries, qmask, Q, max_new, sampler, out_ids,
                           static_cast<cudaStream_t>(stream));
}

}  // extern "C"
