// tcgen05 GEMM for the OPA-DPO hot path (all linears of LLaMA / CLIP / projector / lm_head, LoRA included).
//
//   C[M,N] = epilogue( alpha * ( A1 (M x K1) * B1 (N x K1)^T  +  A2 (M x K2) * B2 (N x K2)^T ) )
//
// The second K-segment is how LoRA is fused into the base-weight GEMM: the LoRA-A product t = s*x*A^T is a
// skinny pre-GEMM and [x | t] * [W^T ; B^T] accumulates into the SAME TMEM accumulator (peft lora.Linear:
// W x + s*B(A x); reference call site opadpo/dpo_models/qlora_model.py:211-216).
//
// Operand "major": K-major means the contraction index is contiguous in memory (row-major [rows, K]);
// MN-major means the output index is contiguous (row-major [K, rows]) — used for dgrad (B = W as stored)
// and for LoRA wgrad (both operands contract over tokens), so no transposed copies are ever made.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdint>

namespace opadpo {

enum GemmEpilogue : int {
  EPI_BF16 = 0,      // C bf16 = act(alpha*acc + bias) + residual
  EPI_F32_ATOMIC = 1,  // C f32 += alpha*acc  (red.global.add; split-K capable)  -- LoRA wgrad into the grad bucket
  EPI_LSE = 2,       // fused lm_head log-sum-exp partials + label-logit gather (never materialises logits)
  EPI_CEGRAD = 3,    // C bf16 = dlp[m]*alpha*(onehot(label[m]) - p) - dent[m]*alpha*p*(ln p + ent[m]),  p = exp(alpha*acc - lse[m])
  EPI_F32_PARTIAL = 4,  // C f32 [splits, M, ldc]: plain per-split stores (decode: weight streaming fills all SMs, fixed-order reduce)
  EPI_SWIGLU_BF16 = 6,  // C bf16 [M, N/2]: B rows interleaved [gate 128 | up 128] per 256-wide tile; C[:, t*128 + j] = silu(gate_j) * up_j
  EPI_SWIGLU_BWD = 7,   // down-projection dgrad with the SwiGLU backward in its epilogue: acc = d(act) tile [M, N = F]; `residual` points at
                        // the saved gate|up activations [M, 2F] (ldr), C is d(gate|up) [M, 2F] (ldc): C[:, j] = dgate_j, C[:, F + j] = dup_j
  EPI_BF16_ROPE = 5,    // C bf16 = rotary(alpha*acc) on the first rope_cols columns (heads of 128, HF rotate_half pairing), plain beyond
};

enum GemmAct : int { ACT_NONE = 0, ACT_GELU = 1, ACT_QUICK_GELU = 2 };

struct GemmOperand {
  const __nv_bfloat16* ptr = nullptr;
  int64_t ld = 0;  // leading dimension in elements
};

struct GemmArgs {
  int M = 0, N = 0;
  int K1 = 0, K2 = 0;
  bool a_mn_major = false, b_mn_major = false;
  GemmOperand A1, B1, A2, B2;
  // Segment-2 A column offset for fused groups (q|k|v, gate|up): off = (n0 / a2_group_n) * a2_group_stride.
  int a2_group_n = 0, a2_group_stride = 0;
  int splits = 1;  // split-K over segment 1 (EPI_F32_ATOMIC only)
  int epilogue = EPI_BF16;
  float alpha = 1.f;
  void* C = nullptr;
  int64_t ldc = 0;
  const __nv_bfloat16* bias = nullptr;      // [N]
  const __nv_bfloat16* residual = nullptr;  // [M, ldr]
  int64_t ldr = 0;
  int act = ACT_NONE;
  // EPI_LSE / EPI_CEGRAD
  const int32_t* labels = nullptr;  // [M]; <0 = ignore
  float4* lse_partials = nullptr;   // [M, ceil(N/256)] (max, sumexp, sum z*exp, -)
  float* label_logit = nullptr;     // [M]
  const float* lse = nullptr;       // [M]
  const float* dlp = nullptr;       // [M] (may be null when only dent is given)
  const float* dent = nullptr;      // EPI_CEGRAD: d(loss)/d(entropy) [M] or null; adds -dent*alpha*p_v*(ln p_v + ent)
  const float* ent = nullptr;       // EPI_CEGRAD: entropy of each row [M] (required with dent)
  // EPI_BF16_ROPE: position of every output row (null: row % rope_S), table [positions][64] of (cos, sin), rotated columns
  const int32_t* rope_pos = nullptr;
  const float2* rope_cs = nullptr;
  int rope_cols = 0, rope_S = 0;
  int max_ctas = 0;                 // 0 = all SMs
  int cluster = -1;                 // -1 auto, 0 off, 1 B-multicast 2-CTA cluster, 2 CTA pair (cta_group::2 MMA, K-major A)
};

// Returns 0 on success, negative on error (message via gemm_last_error()).
int gemm_launch(const GemmArgs& args, cudaStream_t stream);
const char* gemm_last_error();
int gemm_num_n_tiles(int N);
// number of K splits gemm_launch will really use for (K1, K2, requested splits): sizes the EPI_F32_PARTIAL buffer
int gemm_effective_splits(int K1, int K2, int splits);
// bench instrumentation: time every GEMM launch with CUDA events on its stream
void gemm_prof_begin();
int gemm_prof_end(double* total_ms, double* total_flops, int64_t* launches);

}  // namespace opadpo
