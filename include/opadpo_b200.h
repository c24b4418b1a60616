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
enum { OPADPO_EPI_BF16 = 0, OPADPO_EPI_F32_ATOMIC = 1, OPADPO_EPI_LSE = 2, OPADPO_EPI_CEGRAD = 3 };
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
} opadpo_gemm_desc;

int opadpo_op_gemm(const opadpo_gemm_desc* d, void* stream);
int opadpo_gemm_n_tiles(int32_t N);

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
  float* delta;
  void *dq, *dk, *dv; int64_t ld_dqkv;
} opadpo_attn_desc;

int opadpo_op_attn_fwd(const opadpo_attn_desc* d, void* stream);
int opadpo_op_attn_bwd(const opadpo_attn_desc* d, void* stream);
int opadpo_op_pack_mask(const uint8_t* mask, uint64_t* bits, int32_t N, int32_t S, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* OPADPO_B200_H */
