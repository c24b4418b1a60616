// tcgen05 flash attention (forward + backward) for the LLaMA decoder (hd 128, causal, key-padding mask) and the
// CLIP tower (hd 64, non-causal, forward only). Replaces HF LlamaFlashAttention2 -> flash_attn_varlen_func
// [UPSTREAM transformers 4.34.1 / flash-attn 2.5.3], enabled at opadpo/dpo_models/qlora_model.py:133-143.
//
// Layout: q, k, v, o, dO, dq, dk, dv are row-major [T = N*S, H*HD] views (leading dimension given), head h at
// columns [h*HD, (h+1)*HD). lse and delta are fp32 [T, H]. Key validity is a bit mask [N, ceil(S/64)] u64
// (bit c of word j set = key j*64+c may be attended), which also covers CoPO's image-token holes
// (opadpo/dpo_models/rl_models.py:103-105).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdint>

namespace opadpo {

struct AttnArgs {
  int N = 0, S = 0, H = 0, HD = 0;
  bool causal = true;
  float scale = 1.f;
  const __nv_bfloat16 *q = nullptr, *k = nullptr, *v = nullptr;
  int64_t ld_qkv = 0;
  __nv_bfloat16* o = nullptr;  // forward output; backward input
  int64_t ld_o = 0;
  float* lse = nullptr;           // [T, H] natural-log LSE of scaled scores (forward out, backward in)
  const uint64_t* kbits = nullptr;  // may be null = all keys valid
  // Shared-prefix packing (causal only): a row is [prefix (prefix_len) | resp_0 | resp_1 | ...] with segments of
  // segment_len; response queries see the prefix and their own segment. segment_len == 0 disables it.
  int prefix_len = 0, segment_len = 0;
  // backward only
  const __nv_bfloat16* dO = nullptr;
  int64_t ld_do = 0;
  float* delta = nullptr;  // [T, H] scratch (rowsum(dO*O))
  __nv_bfloat16 *dq = nullptr, *dk = nullptr, *dv = nullptr;
  int64_t ld_dqkv = 0;
  // backward only, optional: q and k went through the rotary embedding before the forward (HF LLaMA: column j of a head
  // paired with j + HD/2, angle table rope_cs [position][HD/2] = (cos, sin)); with rope_cs set, the backward returns the
  // gradients w.r.t. the UN-rotated q and k (the inverse rotation is applied to dq / dk in the kernels' epilogues, on the fp32
  // accumulators). rope_pos [T] = position of each token row, or null for row index modulo S.
  const float2* rope_cs = nullptr;
  const int32_t* rope_pos = nullptr;
};

int attn_forward(const AttnArgs& a, cudaStream_t stream);
int attn_backward(const AttnArgs& a, cudaStream_t stream);
// mask bytes [N, S] (non-zero = valid) -> bit words [N, ceil(S/64)]
int attn_pack_mask(const uint8_t* mask, uint64_t* bits, int N, int S, cudaStream_t stream);
const char* attn_last_error();

}  // namespace opadpo
