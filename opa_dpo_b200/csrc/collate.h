// Collator-stage kernels: what the reference does on the CPU between the tokenizer / image decoder
// and the policy forward, as HBM-bound integer / byte kernels. Launchers only; per-element logic lives in collate_core.h.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "collate_core.h"

namespace opadpo {

// segments -> ids int64 [b,R] (+ optional attention mask u8, scores / relations f32, overflow flags i32 [b]).
int collate_responses(const collate::Segments& s, int b, int R, int pad_id, int eos_id, int mode, int64_t* ids, uint8_t* mask,
                      float* scores, float* relations, int32_t* overflow, cudaStream_t st);
// ragged query tokens -> left-padded ids int64 [b,Q] + mask u8 [b,Q]
int collate_queries(const int32_t* tokens, const int64_t* offsets, int b, int Q, int pad_id, int placeholder_id, int image_token,
                    int64_t* ids, uint8_t* mask, cudaStream_t st);
// CoPO image masks: out = images with the pixels (block = 1) or block x block tiles (block = 14) listed in idx [n,m] set
// to the image's mean in every channel. bf16_pixels: pixels and mean rounded to bf16 first. means_out: f32 [n] or null.
int mask_images(const float* images, float* out, int n, int C, int H, int W, const int64_t* idx, int m, int block, int bf16_pixels,
                float* means_out, cudaStream_t st);
// rollout post-processing: out = in with everything behind the deciding stop token set to pad (h_extra: HOST array)
int truncate_after_eos(const int64_t* in, int64_t* out, int b, int R, int64_t eos_id, int64_t pad_id, const int64_t* h_extra,
                       int n_extra, cudaStream_t st);
// CoPO 'attention': out u8 [n,P] = 1 everywhere, 0 at idx [n,m]
int attention_holes(uint8_t* out, const int64_t* idx, int n, int P, int m, cudaStream_t st);
// uint8 RGB images (device, concatenated; offsets/heights/widths are HOST arrays) -> f32 [n,3,S,S]
int preprocess_images(const uint8_t* pixels, const int64_t* h_offsets, const int32_t* h_heights, const int32_t* h_widths, int n,
                      int S, const float* mean3, const float* std3, const uint8_t* bg3, int pad_to_square, float* out,
                      cudaStream_t st);

}  // namespace opadpo
