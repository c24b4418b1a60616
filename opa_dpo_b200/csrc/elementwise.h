// HBM-bound kernels around the GEMMs: norms, RoPE, SwiGLU/GELU, embedding splice, CLIP embedding assembly,
// log-prob finalisation, the fused DPO loss (+grad +stats), grad-norm and AdamW over the flat LoRA bucket,
// decode-time attention and the greedy sampler. All are single-pass, 16-byte vectorised, fp32 math on bf16 storage.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdint>

#ifndef OPADPO_MAX_SUPPRESS
#define OPADPO_MAX_SUPPRESS 8  // include/opadpo_b200.h
#endif

namespace opadpo {

using bf16 = __nv_bfloat16;

// y[r] = w * bf16(x[x_idx[r]] * rsqrt(mean(x^2) + eps))   [UPSTREAM HF LlamaRMSNorm]. x_idx (nullable = identity) lets the
// final norm run only on the R response-predicting rows of each sequence (rl_models.py:120-121 slices logits[:, -R-1:-1]).
int rmsnorm_fwd(const bf16* x, const bf16* w, bf16* y, int64_t rows, int d, float eps, const int32_t* x_idx, cudaStream_t s);
// dx[dx_idx[r]] = (dres ? dres[dx_idx[r]] : 0) + rstd*(w*dy[r]) - x*rstd^3*mean(w*dy*x) with x = x[x_idx[r]]
// (norm weight is frozen: no dw). rstd is recomputed from x. dx may alias dres.
int rmsnorm_bwd(const bf16* dy, const bf16* x, const bf16* w, const bf16* dres, bf16* dx, int64_t rows, int d, float eps,
                const int32_t* x_idx, const int32_t* dx_idx, cudaStream_t s);
// y = (x - mean) * rsqrt(var + eps) * w + b   [UPSTREAM CLIP LayerNorm]
int layernorm_fwd(const bf16* x, const bf16* w, const bf16* b, bf16* y, int64_t rows, int d, float eps, cudaStream_t s);

// In-place rotary embedding (rotate_half layout) on `nheads` heads of width hd starting at x (row stride ld).
// Position of row t is (pos ? pos[t] : (t % S)) + pos_offset + (pos_scalar ? *pos_scalar : 0)  (device scalar: decode). cs: float2 [max_pos, hd/2] (cos, sin). inverse=1 applies the
// transpose rotation (backward).
int rope_inplace(bf16* x, int64_t ld, int64_t T, int S, const int32_t* pos, int pos_offset, const int32_t* pos_scalar,
                 int nheads, int hd, const float2* cs, int inverse, cudaStream_t s);
int rope_table(float2* cs, int max_pos, int hd, float theta, cudaStream_t s);

// act = silu(gu[:, :F]) * gu[:, F:]  ;  backward: dgu from dact and gu
int swiglu_fwd(const bf16* gu, bf16* act, int64_t T, int F, cudaStream_t s);
// gate|up weight [2F, d] -> rows interleaved in blocks of 128 (gate block, up block, ...) for the fused SwiGLU epilogue
int interleave_gate_up(const bf16* W, bf16* out, int F, int d, cudaStream_t s);
int swiglu_bwd(const bf16* dact, const bf16* gu, bf16* dgu, int64_t T, int F, cudaStream_t s);
// exact (erf) GELU, forward and backward (projector, llava_modifications.patch:15354-15367)
int gelu_fwd(const bf16* x, bf16* y, int64_t n, cudaStream_t s);
int gelu_bwd(const bf16* dy, const bf16* x, bf16* dx, int64_t n, cudaStream_t s);

// ids [N, L] (int32; one IMAGE_TOKEN_INDEX=-200 per row) -> embeds [N, S=L-1+P, d]; image rows come from
// feats [n_img, P, d] with image index img_of_row[n]. Also expands the token mask to [N, S] (image rows take
// img_mask [n_img, P] if given, else 1) and records the image slot start img_pos[n]
// (llava_modifications.patch:14618-14706).
int embed_splice(const int32_t* ids, const uint8_t* tok_mask, const uint8_t* img_mask, const int32_t* img_of_row,
                 const bf16* table, const bf16* feats, bf16* out, uint8_t* out_mask, int32_t* img_pos, int N, int L,
                 int P, int d, int vocab, cudaStream_t s);
// dfeats[img, p, :] = sum over rows n with img_of_row[n]==img of dX[n, img_pos[n]+p, :]
int embed_splice_bwd(const bf16* dX, const int32_t* img_of_row, const int32_t* img_pos, bf16* dfeats, int N, int S,
                     int n_img, int P, int d, cudaStream_t s);
// out[r, :] = table[ids[r], :]   (decode-time embedding)
int embed_rows(const int32_t* ids, const bf16* table, bf16* out, int rows, int d, int vocab, cudaStream_t s);

// CLIP: images [n, 3, R, R] (f32) -> patch matrix [n*P, Kp] bf16 (Kp = padded 3*ps*ps), then token assembly.
int clip_im2col(const float* images, bf16* patches, int n, int R, int ps, int Kp, cudaStream_t s);
int clip_assemble(const bf16* patch_emb, const bf16* cls, const bf16* pos, bf16* out, int n, int P, int d, cudaStream_t s);
// drop CLS: out[n, p, :] = x[n, 1+p, :]
int clip_drop_cls(const bf16* x, bf16* out, int n, int P, int d, cudaStream_t s);

// Combine lm_head LSE partials (base-2 domain, see gemm.cu EPI_LSE): logprob = (label_logit - lse) * (label>=0),
// entropy = (lse - E[z]) * (label>=0), lse and the unmasked entropy (ent_raw) saved for backward (all nullable except
// logprob).
int logprob_finalize(const float4* partials, int ntiles, const float* label_logit, const int32_t* labels, float* logprob,
                     float* entropy, float* lse, float* ent_raw, int64_t M, cudaStream_t s);

// One DPO pair (dpo_trainer.py:429-473 + anchor :702-712): per-token loss, gradient and reductions in one pass.
struct DpoPairArgs {
  const float *pi_c, *pi_r, *ref_c, *ref_r;  // [n]
  const float *s_c, *s_r;                    // nullable scores
  float *d_c, *d_r;                          // nullable; accumulated (+=): dL/dpi_c, dL/dpi_r
  float* out;  // [8]: mean loss, mean anchor, masked-mean cr, masked-mean rr, sum mask_c, sum mask_r, -, -
  int64_t n;
  float beta, label_smoothing;
  int reference_free;
  int divergence;  // 0 reverse-KL, 1 JS, 2 alpha
  float alpha_div;
  float loss_coef;    // weight of mean(loss) in the total (gradient uses loss_coef / n)
  int anchor_mode;    // 0 none, 1 mDPO -logsigmoid(sign*(cr - a)), 2 (cr - a)^2
  float anchor_sign, anchor_value, anchor_coef;
  float grad_scale;   // extra factor on every gradient (1/grad_accumulation)
};
int dpo_pair_loss(const DpoPairArgs& a, cudaStream_t s);

// out[0] += sum(g^2)
int sumsq_accum(const float* g, int64_t n, float* out, cudaStream_t s);
// AdamW (utils/trainer_utils.py:9-38; beta (0.9,0.999), eps 1e-8) on fp32 masters; writes the bf16 shadow the GEMMs read.
// grad is scaled by gscale * min(1, max_norm / (sqrt(*sumsq)*gscale + 1e-6)) when sumsq != null and max_norm > 0.
int adamw_step(float* master, const float* grad, float* m, float* v, bf16* shadow, int64_t n, float lr, float b1, float b2,
               float eps, float wd, int step, float gscale, const float* sumsq, float max_norm, cudaStream_t s);
// *out = sqrt(*sumsq) * gscale
int norm_from_sumsq(const float* sumsq, float gscale, float* out, cudaStream_t s);
// out[i] = i * stride + offset
int fill_affine(int32_t* out, int n, int stride, int offset, cudaStream_t s);
// out[i] = i % mod
int fill_mod(int32_t* out, int n, int mod, cudaStream_t s);
// rollout prologue: int64 queries -> int32 ids + mask copy, and out[n, :Q] = queries[n]
int queries_to_i32(const int64_t* queries, const uint8_t* qmask, int32_t* ids, uint8_t* mask, int64_t* out, int64_t out_ld,
                   int N, int Q, cudaStream_t s);
int f32_to_bf16(const float* src, bf16* dst, int64_t n, cudaStream_t s);
int bf16_to_f32(const bf16* src, float* dst, int64_t n, cudaStream_t s);

// Stack the K responses behind the query (rl_models.py:95-112). packed = 0: K*b rows [query | resp_k]; packed = 1: b rows
// [query | resp_0 | ... | resp_{K-1}] (shared prefix). ids/mask [N, L]; labels [K*b*R] in stacked order (-1 at pad).
int stack_inputs(const int64_t* queries, const uint8_t* qmask, const int64_t* const* responses, int K, int b, int Q, int R,
                 int pad_id, int packed, int32_t* ids, uint8_t* mask, int32_t* labels, cudaStream_t s);
// RoPE position of every padded-row slot; gather/scatter row indices of the response-predicting hidden rows;
// fold of the shared last-prefix-row gradients (see elementwise.cu).
int build_positions(int32_t* pos, int N, int S, int S0, int R, int packed, cudaStream_t s);
int build_row_index(int32_t* src, int32_t* dst, int K, int b, int R, int S, int S0, int packed, int T, cudaStream_t s);
int fold_rows(bf16* dx, int K, int b, int S, int S0, int d, int T, cudaStream_t s);

// ---- decode (rollout) -------------------------------------------------------------------------------------
// Copy the prefill keys/values of one layer from the packed qkv activation [N*S, 3d] into the static cache
// kc/vc [N, S_max, d].
int kv_store_prefill(const bf16* qkv, bf16* kc, bf16* vc, int N, int S, int S_max, int d, cudaStream_t s);
// One new token per sequence: qkv [N, 3d] (already rotated); appends k,v at position len = len_base + *cur (device
// counter, so the step is CUDA-graph replayable) and attends over keys 0..len with validity mask kmask [N, S_max] (u8).
int decode_attention(const bf16* qkv, bf16* kc, bf16* vc, const uint8_t* kmask, bf16* o, int N, int H, int hd, int len_base,
                     const int32_t* cur, int S_max, float scale, cudaStream_t s);
// Split-K partial logits [splits, N, V] -> argmax id (lowest index on ties, as torch.argmax). Finished rows
// (done[n] != 0) emit pad_id; done is set when eos is produced; writes next_ids [N] and out[n*out_ld + *cur].
// `suppress` (host array, n_suppress <= OPADPO_MAX_SUPPRESS): HF suppress_tokens, logit := -inf.
int greedy_pick(const float* partials, int splits, int N, int V, int eos_id, int pad_id, const int32_t* suppress,
                int n_suppress, uint8_t* done, int32_t* next_ids, int64_t* out, int64_t out_ld, const int32_t* cur,
                cudaStream_t s);
// Sampling pick: suppress_tokens -> temperature -> top-k -> top-p -> multinomial (HF processor / warper order), counter-based
// RNG keyed by (seed, sequence, step). Any top_k >= 0 (0 = off; ties with the k-th largest survive, as in TopKLogitsWarper)
// and any top_p in (0, 1]; vocab <= 40960.
int sample_pick(const float* partials, int splits, int N, int V, int top_k, float top_p, float temperature, uint64_t seed,
                int eos_id, int pad_id, const int32_t* suppress, int n_suppress, uint8_t* done, int32_t* next_ids,
                int64_t* out, int64_t out_ld, const int32_t* cur, cudaStream_t s);
int advance_counter(int32_t* cur, cudaStream_t s);
// Deterministic reductions of gemm EPI_F32_PARTIAL slices [splits, M, N] fused with the op that follows the linear.
int partial_to_bf16(const float* part, int splits, int64_t n, bf16* out, cudaStream_t s);
// q|k|v slices [splits, M, 3d] -> bf16 [M, 3d] with RoPE (position pos_base + *cur) applied to the q and k heads on the fp32 sums
int partial_qkv_rope(const float* part, int splits, bf16* qkv, int64_t M, int d, int hd, const float2* cs, int pos_base,
                     const int32_t* cur, cudaStream_t s);
int partial_residual_rmsnorm(const float* part, int splits, const bf16* res, bf16* x_out, const bf16* w, bf16* h, int64_t M,
                             int d, float eps, cudaStream_t s);
int partial_swiglu(const float* part, int splits, bf16* act, int64_t M, int F, cudaStream_t s);

}  // namespace opadpo
