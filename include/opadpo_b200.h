/*
 * opadpo_b200 — C ABI of the Blackwell-native OPA-DPO hot path (libopadpo_b200.so).
 *
 * Drop-in boundary for zhyang2226/OPA-DPO's per-token log-prob forward, the DPO loss, the LoRA backward /
 * optimiser step and the greedy rollout. Plain pointers and sizes only: device pointers are raw CUDA device
 * addresses (caller-owned), `stream` is a cudaStream_t passed as void*. Every function returns 0 on success
 * and a negative OPADPO_E* code on failure; the message is available from opadpo_last_error().
 *
 * Two levels are exported:
 *   (1) operator level  (opadpo_op_*)  — one call per fused kernel; what the kernel parity tests bind;
 *   (2) engine level    (opadpo_engine_*, opadpo_forward_logprobs, ...) — what the reference-side plugin binds
 *       (see INTEGRATION.md for the reference-side ctypes stub).
 */
#ifndef OPADPO_B200_H
#define OPADPO_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OPADPO_OK 0
#define OPADPO_EINVAL (-1)
#define OPADPO_ECUDA (-2)
#define OPADPO_ENOMEM (-3)
#define OPADPO_ESTATE (-4)
#define OPADPO_ENOENT (-5) /* a tensor / module name this engine does not hold (distinct from a shape error on one it does) */

/* Message of the last failing call on this thread. */
const char* opadpo_last_error(void);
/* Library / ABI version (bumped on any signature change). */
int opadpo_abi_version(void);

/* ------------------------------------------------------------------------------------------------------------
 * Operator level: the tcgen05 GEMM with LoRA fused as a second K segment.
 *   C[M,N] = epilogue(alpha * (A1[M,K1] * B1[N,K1]^T + A2[M,K2] * B2[N,K2]^T))
 * Replaces (reference, all [UPSTREAM] through peft 0.5.0 lora.Linear wrapped nn.Linear; call sites
 * opadpo/dpo_models/qlora_model.py:133-143,211-216 and opadpo/dpo_models/rl_models.py:120).
 * ---------------------------------------------------------------------------------------------------------- */
enum { OPADPO_EPI_BF16 = 0, OPADPO_EPI_F32_ATOMIC = 1, OPADPO_EPI_LSE = 2, OPADPO_EPI_CEGRAD = 3, OPADPO_EPI_F32_PARTIAL = 4,
       OPADPO_EPI_SWIGLU_BF16 = 6, /* B rows interleaved [128 gate | 128 up] per 256 columns; C bf16 [M, N/2] = silu(gate) * up with
                                      the unfused path's bf16 roundings (bit-identical to projection + SwiGLU) */
       OPADPO_EPI_SWIGLU_BWD = 7, /* down-projection dgrad with the SwiGLU backward in its epilogue: acc = d(act) [M, N = F]; residual =
                                     the saved gate|up activations [M, 2F] (ldr); C = d(gate|up) [M, 2F] (ldc): C[:, j] = d gate_j,
                                     C[:, F + j] = d up_j; same bits as the unfused pair (d(act) rounded to bf16 in between) */
       OPADPO_EPI_BF16_ROPE = 5 /* bf16 output with HF LLaMA's rotary embedding [UPSTREAM apply_rotary_pos_emb] applied to the
                                    first rope_cols columns (heads of 128 columns, column j paired with j + 64) */ };
enum { OPADPO_ACT_NONE = 0, OPADPO_ACT_GELU = 1, OPADPO_ACT_QUICK_GELU = 2 };

typedef struct opadpo_gemm_desc {
  int32_t M, N, K1, K2;
  int32_t a_mn_major, b_mn_major; /* 0: contraction index contiguous; 1: output index contiguous */
  const void* A1; int64_t lda1;   /* bf16 */
  const void* B1; int64_t ldb1;
  const void* A2; int64_t lda2;   /* may be NULL when K2 == 0 */
  const void* B2; int64_t ldb2;
  int32_t a2_group_n, a2_group_stride;
  int32_t splits;
  int32_t epilogue;               /* OPADPO_EPI_* */
  float alpha;
  void* C; int64_t ldc;           /* bf16 (EPI_BF16/CEGRAD) or f32 (EPI_F32_ATOMIC) */
  const void* bias;               /* bf16 [N] or NULL */
  const void* residual; int64_t ldr; /* bf16 [M, ldr] or NULL */
  int32_t act;                    /* OPADPO_ACT_* */
  const int32_t* labels;          /* EPI_LSE / EPI_CEGRAD: [M], negative = ignore */
  void* lse_partials;             /* EPI_LSE: float4 [M, opadpo_gemm_n_tiles(N)] */
  float* label_logit;             /* EPI_LSE: [M] */
  const float* lse;               /* EPI_CEGRAD: [M] */
  const float* dlp;               /* EPI_CEGRAD: [M] */
  int32_t max_ctas;               /* 0 = one CTA per SM */
  int32_t cluster_mode;           /* 0 auto, 1 off, 2: 2-CTA cluster that multicasts the shared B tile, 3: CTA pair (tcgen05 cta_group::2, 256x256 MMA) */
  const float* dent;              /* EPI_CEGRAD: d(loss)/d(entropy) [M] or NULL (opa_models/opa_trainer.py:63,82-85) */
  const float* ent;               /* EPI_CEGRAD: per-row entropy [M], required with dent */
  const int32_t* rope_pos;        /* EPI_BF16_ROPE: position of every output row [M], or NULL = row % rope_S */
  const void* rope_cs;            /* EPI_BF16_ROPE: float2 (cos, sin) table [positions, 64] */
  int32_t rope_cols, rope_S;
} opadpo_gemm_desc;

int opadpo_op_gemm(const opadpo_gemm_desc* d, void* stream);
int opadpo_gemm_n_tiles(int32_t N);
/* Weight-streaming GEMM of the rollout's decode step (one new token per sequence: model.generate's cached steps,
 * online_generator.py:292-309): part[split, m, n] = alpha * sum_{k in split} x[m, k] * W[n, k] for Mb <= 128 activation
 * rows x [Mb, K] bf16 against a Linear weight W [N, K] bf16; fp32 slices part [splits, Mb, N] that the consumer sums in
 * slice order (deterministic). splits = 0 picks the count (balance of the persistent grid vs slice traffic), max_splits
 * caps it (0 = 16); the count used is returned through splits_out. */
int opadpo_op_decode_gemm(const void* x, int64_t ldx, const void* W, int64_t ldw, int32_t Mb, int32_t N, int32_t K,
                          int32_t splits, int32_t max_splits, float alpha, float* part, int32_t* splits_out, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Operator level: tcgen05 flash attention. q/k/v/o/dO/dq/dk/dv are bf16 row-major [N*S, H*HD] views with the
 * given leading dimensions (head h at columns [h*HD,(h+1)*HD)); lse/delta are f32 [N*S, H]; kbits is the packed
 * key-validity mask u64 [N, ceil(S/64)] (NULL = all valid) made by opadpo_op_pack_mask.
 * Replaces HF LlamaFlashAttention2 / flash_attn_varlen_func and CLIPAttention [UPSTREAM]; call site
 * opadpo/dpo_models/qlora_model.py:133-143 (use_flash_attention_2) and rl_models.py:120.
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct opadpo_attn_desc {
  int32_t N, S, H, HD;
  int32_t causal;
  float scale;
  const void *q, *k, *v; int64_t ld_qkv;
  void* o; int64_t ld_o;
  float* lse;
  const uint64_t* kbits;
  const void* dO; int64_t ld_do;   /* backward only from here */
  float* delta;                    /* f32 [N*S, H] scratch: rowsum(dO * O), written by the backward's first kernel, read by its second */
  void *dq, *dk, *dv; int64_t ld_dqkv;
  /* shared-prefix packing (causal): rows are [prefix | resp_0 | resp_1 ...] with segments of segment_len tokens after
   * prefix_len positions; a response query sees the prefix and its own segment only. segment_len = 0: off. */
  int32_t prefix_len, segment_len;
} opadpo_attn_desc;

int opadpo_op_attn_fwd(const opadpo_attn_desc* d, void* stream);
int opadpo_op_attn_bwd(const opadpo_attn_desc* d, void* stream);
int opadpo_op_pack_mask(const uint8_t* mask, uint64_t* bits, int32_t N, int32_t S, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Operator level: one DPO preference pair — per-token sigmoid log-ratio loss, its gradient w.r.t. the policy
 * log-probs, the AncPO anchor term and the masked-mean reward statistics, in one single-CTA warp-reduce kernel.
 * Replaces DPOTrainer.dpo_loss (opadpo/dpo_models/dpo_trainer.py:429-473), the `.mean()` reductions of
 * compute_policy_loss (:631,693), the anchor (:702-712) and the reward stats (:785-796).
 * All pointers are device f32 [n] (n = b*R); s_c/s_r/d_c/d_r may be NULL. d_c/d_r are accumulated (+=).
 * out[8] = {mean loss, mean anchor, masked-mean chosen reward, masked-mean rejected reward, #chosen mask,
 *          #rejected mask, 0, 0} where reward = beta*(pi - ref) and mask = (ref != 0).
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct opadpo_dpo_pair_desc {
  const float *pi_c, *pi_r, *ref_c, *ref_r;
  const float *s_c, *s_r;
  float *d_c, *d_r;
  float* out;
  int64_t n;
  float beta, label_smoothing;
  int32_t reference_free;
  int32_t divergence;      /* 0 reverse_kl, 1 js_divergence, 2 alpha_divergence */
  float alpha_div;
  float loss_coef;         /* weight of mean(loss) in the total loss */
  int32_t anchor_mode;     /* 0 none, 1 mDPO: -logsigmoid(sign*(cr - a)), 2: (cr - a)^2 */
  float anchor_sign, anchor_value, anchor_coef;
  float grad_scale;        /* multiplies every gradient (1/gradient_accumulation_steps) */
} opadpo_dpo_pair_desc;
int opadpo_op_dpo_pair(const opadpo_dpo_pair_desc* d, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Operator level: the collator stage — what the reference does on the CPU between the tokenizer /
 * image decoder and the policy forward. Tokenisation, JSON parsing and JPEG decoding stay with the caller; everything
 * after them runs here. Integer / byte work: results are bit-identical to the reference's.
 * ---------------------------------------------------------------------------------------------------------- */
/* Raw tokenizer output (input_ids INCLUDING the leading BOS) of every `copied content` / `rewritten content` entry of a
 * batch of GPT-4V reports, flattened (DataCollatorForCausalLM.__call__, utils/data_utils_dpo.py:139-199). All device. */
typedef struct opadpo_segments {
  const int32_t* tokens;      /* [seg_offsets[n_segments]] */
  const int64_t* seg_offsets; /* [n_segments + 1] */
  const float* seg_score;     /* [n_segments] score_mapping weight (data_utils_dpo.py:32-37), or NULL */
  const float* seg_relation;  /* [n_segments] error_type_mapping weight (:38-42), or NULL */
  const uint8_t* seg_flags;   /* [n_segments] OPADPO_SEG_* */
  const int64_t* row_offsets; /* [b + 1]: segments of row i are [row_offsets[i], row_offsets[i+1]) in report-key order */
  int32_t b;
} opadpo_segments;
enum { OPADPO_SEG_KEEP_BOS = 1,     /* entry of the FIRST report key: its BOS stays (`count == 0`, :170,183) */
       OPADPO_SEG_STRIP_SPACE = 2,  /* copied entry: a leading 29871 piece is dropped (:186-187) */
       OPADPO_SEG_VERBATIM = 4      /* plain tokenisation (:120-137): the tokens as they are — a BOS-only text still puts its
                                       BOS in the row, whereas a BOS-only report entry is absent (:168,180) */ };
enum { OPADPO_COLLATE_ADD_EOS = 1,  /* add_eos (:73-87): the first pad position becomes EOS */
       OPADPO_COLLATE_PAD_EOS = 2,  /* pad_eos (:89-99): the EOS position inherits the previous position's weight (1 if 0) */
       OPADPO_COLLATE_TRUNCATE = 4, /* tokenizer truncation (plain branch :120-137); without it a row longer than R sets
                                       overflow[row], which is where the reference raises and falls back (:259-278) */
       OPADPO_COLLATE_TOKEN_MASK = 8 /* attn_mask = the tokenizer's own mask as the plain branch hands it out (:122-137): 1
                                        where a token was placed; with ADD_EOS its first 0 holds eos_id, because add_eos
                                        walks every tensor of the encoding (:76-79). Default: ids != pad (:253-255) */ };
/* -> ids int64 [b,R] right-padded; attn_mask u8 [b,R] (ids != pad), scores / relations f32 [b,R] (0 at pads), overflow
 * i32 [b]: each nullable. Replaces pad_and_stack / add_eos / pad_eos and the torch.cat chain of :200-235. */
int opadpo_op_collate_responses(const opadpo_segments* s, int32_t R, int32_t pad_id, int32_t eos_id, int32_t mode,
                                int64_t* ids, uint8_t* attn_mask, float* scores, float* relations, int32_t* overflow,
                                void* stream);
/* Ragged query tokens (device: tokens i32, offsets i64 [b+1]) -> left-padded, right-truncated ids int64 [b,Q] with
 * placeholder_id (30861) replaced by image_token (-200), and the attention mask u8 [b,Q] (data_utils_dpo.py:118-121). */
int opadpo_op_collate_queries(const int32_t* tokens, const int64_t* offsets, int32_t b, int32_t Q, int32_t pad_id,
                              int32_t placeholder_id, int32_t image_token, int64_t* ids, uint8_t* attn_mask, void* stream);
/* Rollout post-processing (truncate_after_eos_with_padding, opadpo/generator_models/generator.py:244-273, called at
 * online_generator.py:310-323): per row of ids int64 [b,R] the cut is the first EOS, then overridden — in list order, even
 * when it lies behind the EOS — by the first occurrence of every additional stop token that is present; everything after
 * the cut becomes pad_id; a row without any stop token is left alone. extra: HOST int64 [n_extra], n_extra <= 8. out may
 * alias in. Replaces a .tolist() round trip through the host. */
int opadpo_op_truncate_after_eos(const int64_t* ids, int64_t* out, int32_t b, int32_t R, int64_t eos_id, int64_t pad_id,
                                 const int64_t* extra, int32_t n_extra, void* stream);
/* CoPO image masks (mask_single_image, opadpo/dpo_models/dpo_trainer.py:83-109): out [n,C,H,W] f32 = images with, per
 * image, the pixels (block = 1, 'random') or block x block tiles (block = 14, 'blockwise') listed in idx int64 [n,m]
 * (the caller's torch.randperm prefix, so the random stream stays the reference's) set to that image's mean in every
 * channel. bf16_pixels != 0: pixels and mean are rounded to bf16 first (the trainer masks bf16 images, :283-286).
 * means_out f32 [n] nullable. out may alias images. */
int opadpo_op_mask_images(const float* images, float* out, int32_t n, int32_t C, int32_t H, int32_t W, const int64_t* idx,
                          int32_t m, int32_t block, int32_t bf16_pixels, float* means_out, void* stream);
/* CoPO 'attention' (mask_percentage_per_row, dpo_trainer.py:119-125): out u8 [n,P] = 1, 0 at idx int64 [n,m]. */
int opadpo_op_attention_holes(uint8_t* out, const int64_t* idx, int32_t n, int32_t P, int32_t m, void* stream);
/* DPO_Dataset.__getitem__ image path (utils/data_utils_dpo.py:311-335): decoded uint8 RGB images [H_i, W_i, 3]
 * concatenated in DEVICE memory at byte offsets h_offsets (HOST int64 [n]; heights / widths HOST int32 [n]) ->
 * expand2square on the background colour bg3 (HOST u8[3]; NULL = int(mean*255)) when pad_to_square, Pillow bicubic
 * resize of the shortest edge to S (two 8-bit passes with 22-bit fixed-point coefficients), centre crop SxS, rescale
 * 1/255, (x - mean)/std -> out f32 [n,3,S,S]. mean3 / std3: HOST float[3]. */
int opadpo_op_preprocess_images(const uint8_t* pixels, const int64_t* h_offsets, const int32_t* h_heights,
                                const int32_t* h_widths, int32_t n, int32_t S, const float* mean3, const float* std3,
                                const uint8_t* bg3, int32_t pad_to_square, float* out, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Engine level. One engine per process per GPU (current CUDA device at create time). The engine owns the base
 * weights (bf16), two LoRA adapter sets (0 = lora_policy, 1 = lora_ref_policy; qlora_model.py:21,66-93), the
 * trainable adapter's fp32 masters / gradients / Adam moments as ONE flat bucket, and all activations.
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct opadpo_engine opadpo_engine;

typedef struct opadpo_config {
  /* LLaMA decoder [UPSTREAM LlamaConfig] */
  int32_t d_model, n_layers, n_heads, d_ffn, vocab;
  float rms_eps, rope_theta;
  int32_t max_positions;
  /* CLIP vision tower [UPSTREAM CLIPVisionConfig]; v_layers = number of encoder layers EXECUTED
   * (mm_vision_select_layer = -2 on a 24-layer tower -> 23; llava_modifications.patch:15185-15217) */
  int32_t v_d, v_layers, v_heads, v_ffn, v_image, v_patch;
  float v_eps;
  /* LoRA (opa_train.py:466-484): r, alpha; targets = every Linear but lm_head */
  int32_t lora_r;
  float lora_alpha;
  int32_t pad_id, eos_id, image_token;   /* 0, 2, -200 (utils/constants.py:20-33) */
} opadpo_config;

enum { OPADPO_DT_BF16 = 0, OPADPO_DT_F32 = 1 };

int opadpo_engine_create(const opadpo_config* cfg, opadpo_engine** out);
int opadpo_engine_destroy(opadpo_engine* e);
/* Copy one base tensor (HF state-dict name, e.g. "model.layers.3.self_attn.q_proj.weight") from device memory. */
int opadpo_engine_load_tensor(opadpo_engine* e, const char* name, const void* dev_ptr, int32_t dtype, int64_t rows,
                              int64_t cols, void* stream);
/* Copy one LoRA pair for module path `module` (e.g. "model.layers.3.self_attn.q_proj"): A [r, in], B [out, r]. */
int opadpo_engine_load_lora(opadpo_engine* e, int32_t adapter, const char* module, const void* A, const void* B,
                            int32_t dtype, void* stream);
int opadpo_engine_export_lora(opadpo_engine* e, int32_t adapter, const char* module, float* A_out, float* B_out,
                              void* stream);
/* Shape of a LoRA target: in/out features; returns OPADPO_EINVAL if `module` is not a LoRA target. */
int opadpo_engine_lora_shape(opadpo_engine* e, const char* module, int64_t* in_features, int64_t* out_features);
/* Make `adapter` trainable: allocates fp32 masters (from the current values), gradient bucket and Adam moments. */
int opadpo_engine_set_trainable(opadpo_engine* e, int32_t adapter);
/* Flat buffers of the trainable adapter (device pointers; n_train leading elements receive gradients). */
int opadpo_engine_buffers(opadpo_engine* e, float** master, float** grad, float** exp_avg, float** exp_avg_sq,
                          int64_t* n_total, int64_t* n_train);
/* Where module's A [groups*r... see DESIGN.md] and B live inside the flat bucket (element offsets). */
int opadpo_engine_lora_offsets(opadpo_engine* e, const char* module, int64_t* offA, int64_t* offB);
/* Engine options. "share_prefix" (default 1): compute the image+query prefix once per sample and let its K responses
 * attend to it through the segment mask (identical results to the reference's K stacked rows, fewer FLOPs).
 * "merge_frozen" (default 0): inference forwards (save_slot = -1) and rollouts of a NON-trainable adapter run on decoder
 * weights merged once, W' = W + (alpha/r) B A (peft merge_adapter semantics; bf16 W'), so the reference pass of the DPO
 * step carries no LoRA GEMMs; costs one extra copy of the decoder weights, rebuilt when that adapter is reloaded.
 * "fuse_rope" (default 1): the q|k|v projection applies RoPE in its GEMM epilogue (one bf16 rounding instead of two), and
 * the attention backward applies the inverse rotation to dq / dk in its epilogues (no standalone RoPE kernel in either pass).
 * "fuse_swiglu" (default 1; acts with merge_frozen): the merged frozen adapter's gate|up projection
 * applies SwiGLU in its epilogue over an interleaved copy of the merged weight (bit-identical results).
 * "fuse_swiglu_bwd" (default 0): the down projection's dgrad GEMM turns its d(act) tile into d(gate), d(up) in the epilogue
 * (OPADPO_EPI_SWIGLU_BWD; bit-identical to the separate SwiGLU-backward pass, d(act) never goes to memory; measured no
 * faster at 7B — the GEMM becomes epilogue-bound — hence opt-in).
 * "wgrad_stream" (default 1): the LoRA weight-gradient GEMMs of opadpo_backward run on an engine-owned side stream and are
 * joined before the call's last kernel, the layer callback and the projector (same results up to fp32-atomics order).
 * "merge_rollout" (default 1): rollouts merge the adapter per call. "adam_step": optimizer-state restore. */
int opadpo_engine_set_option(opadpo_engine* e, const char* name, int64_t value);
/* Communication overlap: `cb(layer, user)` is called on the host thread inside opadpo_backward right after the backward
 * kernels of decoder layer `layer` (n_layers-1 .. 0), then of the projector (layer == n_layers), have been enqueued; the
 * slice [offset, offset+count) of the gradient bucket reported by opadpo_engine_layer_grad_range is then final in stream
 * order, so the caller can start its all-reduce while earlier layers are still being differentiated. cb = NULL disables. */
typedef void (*opadpo_layer_cb)(int32_t layer, void* user);
int opadpo_engine_set_backward_callback(opadpo_engine* e, opadpo_layer_cb cb, void* user);
int opadpo_engine_layer_grad_range(opadpo_engine* e, int32_t layer, int64_t* offset, int64_t* count);
/* Make the engine accumulate the trainable adapter's gradients in a caller-owned device buffer of `count` >= n_train floats
 * (e.g. a symmetric-memory allocation every data-parallel rank maps, so the all-reduce can run in place). The current
 * gradients are copied over, the engine's own bucket is freed; the buffer must outlive the engine. */
int opadpo_engine_adopt_grad_buffer(opadpo_engine* e, float* buffer, int64_t count);
/* After the caller modified the fp32 masters directly (e.g. optimizer state restore): refresh the bf16 shadows. */
int opadpo_engine_sync_shadow(opadpo_engine* e, void* stream);

/* AutoregressivePolicy.forward (opadpo/dpo_models/rl_models.py:75-144): K responses per query, stacked [K*b] rows.
 *   images [b,3,v_image,v_image] f32; queries int64 [b,Q] (left-padded, one image_token); qmask u8 [b,Q];
 *   img_mask u8 [b,P] or NULL (CoPO 'attention' holes, rl_models.py:103-105); responses: K device pointers to
 *   int64 [b,R] (right-padded). out_logprobs / out_entropies (nullable): f32 [K,b,R].
 *   save_slot: -1 = inference; 0/1 = keep activations for opadpo_backward(slot). */
int opadpo_forward_logprobs(opadpo_engine* e, int32_t adapter, const float* images, int32_t b, const int64_t* queries,
                            const uint8_t* qmask, const uint8_t* img_mask, int32_t Q, int32_t K,
                            const int64_t* const* responses, int32_t R, float temperature, float* out_logprobs,
                            float* out_entropies, int32_t save_slot, void* stream);
/* d(loss)/d(logprobs) [K,b,R] f32 -> accumulates LoRA gradients into the flat bucket (autograd of rl_models.py:120
 * through the trainable adapter; rl_trainer.py:162). Releases the slot. */
int opadpo_backward(opadpo_engine* e, int32_t save_slot, const float* dlogprobs, void* stream);
/* Same with a second incoming gradient d(loss)/d(entropies) [K,b,R] f32 (either pointer may be NULL, not both): the
 * entropy output of opadpo_forward_logprobs is differentiable too, which is what OPA's entropy regulariser needs
 * (opadpo/opa_models/opa_trainer.py:63,82-85,113-116; autograd through `entropies`). */
int opadpo_backward_ent(opadpo_engine* e, int32_t save_slot, const float* dlogprobs, const float* dentropies, void* stream);
int opadpo_zero_grad(opadpo_engine* e, void* stream);
/* clip_grad_norm_ (rl_trainer.py:165-168) + AdamW (utils/trainer_utils.py:9-38) over the flat bucket.
 * grad_scale multiplies the bucket first (1/world_size after a sum all-reduce). grad_norm_out: device f32 or NULL. */
int opadpo_optimizer_step(opadpo_engine* e, float lr, float beta1, float beta2, float eps, float weight_decay,
                          float max_grad_norm, float grad_scale, float* grad_norm_out, void* stream);
/* Rollout (model.generate of online_generator.py:292-309 / rl_models.py:166-188): writes out_ids int64
 * [b, Q + max_new] = queries | generated (pad after EOS). adapter = -1 runs the bare base model.
 * sampler = NULL or do_sample = 0: greedy (ids reproduce the reference exactly). do_sample = 1: HF warper order
 * temperature -> top_k -> top_p -> multinomial with a counter-based RNG keyed by (seed, sequence, step); any top_k >= 0
 * (0 = off; ties with the k-th largest survive, TopKLogitsWarper) and any top_p in (0, 1] (TopPLogitsWarper).
 * suppress_tokens[0..n_suppress): HF suppress_tokens (logit := -inf before argmax / the warpers), what
 * online_generator.py:298-302 and rl_models.py:173-177 pass as [eos] under suppress_eos_at_generation; greedy too. */
#define OPADPO_MAX_SUPPRESS 8
typedef struct opadpo_sampler {
  int32_t do_sample;
  int32_t top_k;
  float top_p;
  float temperature;
  uint64_t seed;
  int32_t n_suppress;
  int32_t suppress_tokens[OPADPO_MAX_SUPPRESS];
} opadpo_sampler;
int opadpo_generate(opadpo_engine* e, int32_t adapter, const float* images, int32_t b, const int64_t* queries,
                    const uint8_t* qmask, int32_t Q, int32_t max_new, const opadpo_sampler* sampler, int64_t* out_ids,
                    void* stream);
/* ------------------------------------------------------------------------------------------------------------
 * Data-parallel gradient all-reduce as one kernel over NVLink / NVSwitch peer memory (csrc/collective.cu): the local fp32
 * slice is packed to bf16 into a SYMMETRIC staging buffer (same allocation mapped by every rank: multicast address for NVLS,
 * or per-peer addresses), summed over ranks with fp32 accumulation (multimem.ld_reduce in the switch; each rank reduces its
 * 1/world share and multimem.st's the result to all ranks), and unpacked back into the fp32 slice. In place; the result is
 * bit-identical on every rank. Replaces the DDP reduction of the LoRA gradients in the reference's accelerate loop
 * (opadpo/dpo_models/rl_trainer.py:155-175; torch DDP over NCCL [UPSTREAM]). Every rank must issue the same sequence of
 * calls (same count / blocks / epoch) on one stream per rank. Mapping the symmetric buffers is the caller's job (the Python
 * host uses torch.distributed._symmetric_memory; any CUDA VMM / multicast-object exporter works).
 *   signal_peer[r]: address, in THIS process, of rank r's signal area: uint32 [OPADPO_AR_MAX_BLOCKS][OPADPO_MAX_RANKS],
 *                   zero before the first call; epoch: first call 1, then +2 per call (entry barrier = epoch, exit = epoch+1);
 *   err_flag: host-visible (pinned) uint32 the kernel sets if a cross-rank wait exceeded timeout_ms (0 = 60 s) — the
 *             kernel then finishes without hanging the GPU and the caller must treat the step as failed.
 * ---------------------------------------------------------------------------------------------------------- */
#define OPADPO_MAX_RANKS 8
#define OPADPO_AR_MAX_BLOCKS 256
typedef struct opadpo_allreduce_desc {
  int32_t rank, world;
  float* grad;                             /* local fp32 slice, 32 B aligned */
  void* stage_local;                       /* local address of the slice's bf16 staging, 16 B aligned */
  uint64_t stage_mc;                       /* multicast address of the same staging bytes, or 0 (then stage_peer is used) */
  uint64_t stage_peer[OPADPO_MAX_RANKS];   /* per-rank addresses of the same staging bytes (rank's own = stage_local) */
  uint64_t signal_peer[OPADPO_MAX_RANKS];
  int64_t count;                           /* elements, multiple of 8 */
  uint32_t epoch;
  int32_t blocks;                          /* CTAs (no shared memory: they co-reside with the backward's GEMM CTAs) */
  uint32_t* err_flag;
  int32_t timeout_ms;
  int32_t fp32_in_place;                   /* 1: no staging — `grad` is itself a symmetric allocation and stage_local / stage_mc /
                                              stage_peer are the mappings of THIS fp32 slice; exact fp32 sums, half the HBM traffic
                                              of the bf16-wire form (which needs no symmetric gradient bucket) */
} opadpo_allreduce_desc;
int opadpo_op_allreduce_bf16wire(const opadpo_allreduce_desc* d, void* stream);

/* Measurement hooks (bench.py): between begin and end every tcgen05 GEMM launch is bracketed by CUDA events on its own
 * stream; end synchronises and returns the summed kernel time, the executed FLOPs (2*M*N*K with K padded to the 64-wide
 * k-block) and the launch count. */
int opadpo_prof_begin(void);
int opadpo_prof_end(double* gemm_ms, double* gemm_flops, int64_t* gemm_launches);
/* Number of kernels this library has launched so far in this process. */
int64_t opadpo_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* OPADPO_B200_H */
