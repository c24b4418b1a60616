// Per-element logic of the collator-stage kernels (collate.cu), written once as host/device functions so that the very
// same arithmetic the GPU threads run can also be compiled with g++ into the CPU test harness (tests/host_harness) and
// compared with the oracle bit-for-bit. Nothing here launches anything; the product only ever calls these from kernels.
//
// Reference behaviour being reproduced (paths under /root/reference):
//   utils/data_utils_dpo.py:44-99, 118-121, 139-235   (segments -> padded ids / score / image-relation rows, EOS rules)
//   utils/data_utils_dpo.py:311-333                   (expand2square + CLIP preprocess; Pillow ImagingResample [UPSTREAM])
//   opadpo/dpo_models/dpo_trainer.py:83-125           (CoPO masks)
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define OPADPO_HD __host__ __device__ __forceinline__
#else
#define OPADPO_HD inline
#endif

namespace opadpo {
namespace collate {

constexpr int32_t kSpaceToken = 29871;  // data_utils_dpo.py:186
enum : int32_t { SEG_KEEP_BOS = 1, SEG_STRIP_SPACE = 2, SEG_VERBATIM = 4 };
enum : int32_t { MODE_ADD_EOS = 1, MODE_PAD_EOS = 2, MODE_TRUNCATE = 4, MODE_TOKEN_MASK = 8 };
constexpr int kPrecisionBits = 32 - 8 - 2;  // Pillow's 8bpc fixed-point coefficient precision

// Raw tokenizer output of every report entry, flattened; see include/opadpo_b200.h (opadpo_segments).
struct Segments {
  const int32_t* tokens;
  const int64_t* seg_offsets;
  const float* seg_score;
  const float* seg_relation;
  const uint8_t* seg_flags;
  const int64_t* row_offsets;
};

// The slice of segment j that reaches the row (data_utils_dpo.py:166-186): nothing when the entry tokenised to BOS only;
// BOS dropped unless the entry belongs to the first report key; a copied entry then loses a leading 29871 piece.
OPADPO_HD void segment_range(const Segments& s, int64_t j, int64_t* begin, int64_t* end) {
  int64_t b = s.seg_offsets[j];
  const int64_t e = s.seg_offsets[j + 1];
  const int fl = s.seg_flags[j];
  if (fl & SEG_VERBATIM) {  // plain tokenisation (data_utils_dpo.py:120-137): the tokenizer output as is, even when it is BOS only
    *begin = b;
    *end = e;
    return;
  }
  if (e - b <= 1) {
    *begin = *end = e;
    return;
  }
  if (!(fl & SEG_KEEP_BOS)) b += 1;
  if ((fl & SEG_STRIP_SPACE) && s.tokens[b] == kSpaceToken) b += 1;
  *begin = b;
  *end = e;
}

OPADPO_HD int64_t row_length(const Segments& s, int row) {
  int64_t n = 0;
  for (int64_t j = s.row_offsets[row]; j < s.row_offsets[row + 1]; ++j) {
    int64_t b, e;
    segment_range(s, j, &b, &e);
    n += e - b;
  }
  return n;
}

// Token and per-token weights at position p of the concatenated row; false when p lies beyond its tokens.
OPADPO_HD bool token_at(const Segments& s, int row, int64_t p, int32_t* tok, float* score, float* rel) {
  int64_t acc = 0;
  for (int64_t j = s.row_offsets[row]; j < s.row_offsets[row + 1]; ++j) {
    int64_t b, e;
    segment_range(s, j, &b, &e);
    const int64_t n = e - b;
    if (p < acc + n) {
      *tok = s.tokens[b + (p - acc)];
      *score = s.seg_score ? s.seg_score[j] : 0.f;
      *rel = s.seg_relation ? s.seg_relation[j] : 0.f;
      return true;
    }
    acc += n;
  }
  return false;
}

// add_eos (data_utils_dpo.py:73-87) and pad_eos (:89-99) on a finished row. first_pad / first_eos: smallest position
// holding pad_id / eos_id before the insertion (R when none).
OPADPO_HD void finish_row(int64_t* ids, float* score, float* rel, int R, int first_pad, int first_eos, int mode, int eos_id) {
  if ((mode & MODE_ADD_EOS) && first_pad < R) {
    ids[first_pad] = eos_id;
    if (first_pad < first_eos) first_eos = first_pad;
  }
  if ((mode & MODE_PAD_EOS) && first_eos < R) {
    const int prev = first_eos == 0 ? R - 1 : first_eos - 1;  // index -1 wraps to the last position, as in the reference
    if (score) {
      const float v = score[prev];
      score[first_eos] = v != 0.f ? v : 1.f;
    }
    if (rel) {
      const float v = rel[prev];
      rel[first_eos] = v != 0.f ? v : 1.f;
    }
  }
}

// Left-padded query row (data_utils_dpo.py:118-121): right-truncated to Q, placeholder id -> image token.
OPADPO_HD void query_at(const int32_t* tokens, const int64_t* offsets, int row, int p, int Q, int pad_id, int placeholder_id,
                        int image_token, int64_t* id, uint8_t* mask) {
  const int64_t b = offsets[row];
  int64_t len = offsets[row + 1] - b;
  if (len > Q) len = Q;
  const int64_t start = Q - len;
  if (p >= start) {
    const int32_t t = tokens[b + (p - start)];
    *id = t == placeholder_id ? image_token : t;
    *mask = 1;
  } else {
    *id = pad_id;
    *mask = 0;
  }
}

// truncate_after_eos_with_padding (opadpo/generator_models/generator.py:244-273): the position after which a rollout row
// becomes pad, or -1 to leave the row alone. first[0] = first index of EOS, first[1 + k] = first index of the k-th extra
// stop token (R when absent). A stop token that is present OVERRIDES whatever came before it in the list, even when it
// lies behind the EOS — the reference's loop assigns, it does not take the minimum.
constexpr int kMaxStopTokens = 8;
OPADPO_HD int truncate_end_from_firsts(const int* first, int n_extra, int R) {
  int end = first[0] < R ? first[0] : -1;
  for (int k = 0; k < n_extra; ++k)
    if (first[1 + k] < R) end = first[1 + k];
  return end;
}

// ------------------------------------------------------------------------------------------------ images
// One image's geometry and where its resampling tables live. The source is the uint8 RGB image [H, W, 3] seen through
// expand2square (a Ws x Hs canvas of the background colour with the image at (ox, oy)); the output is the SxS window at
// (left, top) of the canvas resized to new_w x new_h. Tables (int32, Pillow's normalised 22-bit coefficients):
// kkx [S, ksx], bx [S, 2] = (first source column, taps), kky [S, ksy], by [S, 2].
struct ImageDesc {
  int64_t pix_off;   // byte offset of the image in the pixel buffer
  int64_t tab_off;   // int32 offset of its tables: kkx | bx | kky | by
  int64_t tmp_off;   // byte offset of its horizontal-pass rows [Hs, S, 3] in the scratch buffer
  int32_t H, W, Hs, Ws, ox, oy;
  int32_t ksx, ksy;
};

OPADPO_HD uint8_t clip8(int32_t v) {
  v >>= kPrecisionBits;  // arithmetic shift, as Pillow's clip8
  return static_cast<uint8_t>(v < 0 ? 0 : (v > 255 ? 255 : v));
}

OPADPO_HD int32_t canvas_pixel(const uint8_t* img, const ImageDesc& d, int x, int y, int c, const uint8_t* bg) {
  const int sx = x - d.ox, sy = y - d.oy;
  if (sx < 0 || sy < 0 || sx >= d.W || sy >= d.H) return bg[c];
  return img[(static_cast<int64_t>(sy) * d.W + sx) * 3 + c];
}

// Horizontal pass of ImagingResample for canvas row y, output column xx (window-relative), channel c.
OPADPO_HD uint8_t resample_h(const uint8_t* img, const ImageDesc& d, const int32_t* tab, int S, int y, int xx, int c,
                             const uint8_t* bg) {
  const int32_t* kk = tab + static_cast<int64_t>(xx) * d.ksx;
  const int32_t* bx = tab + static_cast<int64_t>(S) * d.ksx + 2 * xx;
  const int x0 = bx[0], n = bx[1];
  int32_t ss = 1 << (kPrecisionBits - 1);
  for (int k = 0; k < n; ++k) ss += canvas_pixel(img, d, x0 + k, y, c, bg) * kk[k];
  return clip8(ss);
}

// rescale + normalize of the pinned image processor: uint8 * (1/255) in float64 -> float32, then (x - mean) / std in
// float32 [UPSTREAM transformers 4.34.1 image_transforms.rescale / normalize].
OPADPO_HD float normalise_pixel(uint8_t v, float mean, float std) {
  const float x = static_cast<float>(static_cast<double>(v) * (1.0 / 255.0));
  return (x - mean) / std;
}

// Canvas rows above / below the pasted image hold only the background colour, so their horizontal pass is the same for
// all of them: it is computed once, for the first such row, and every other padding row reads that one.
OPADPO_HD bool is_padding_row(const ImageDesc& d, int y) { return y < d.oy || y >= d.oy + d.H; }
OPADPO_HD int tmp_row(const ImageDesc& d, int y) {
  if (!is_padding_row(d, y)) return y;
  return d.oy > 0 ? 0 : d.oy + d.H;
}
OPADPO_HD bool h_row_needed(const ImageDesc& d, int y) { return tmp_row(d, y) == y; }

// Vertical pass over the horizontal-pass rows `tmp` [Hs, S, 3] for output (yy, xx, c), then normalisation.
OPADPO_HD float resample_v(const uint8_t* tmp, const ImageDesc& d, const int32_t* tab, int S, int yy, int xx, int c, float mean,
                           float std) {
  const int32_t* kky = tab + static_cast<int64_t>(S) * d.ksx + 2 * S;
  const int32_t* kk = kky + static_cast<int64_t>(yy) * d.ksy;
  const int32_t* by = kky + static_cast<int64_t>(S) * d.ksy + 2 * yy;
  const int y0 = by[0], n = by[1];
  int32_t ss = 1 << (kPrecisionBits - 1);
  for (int k = 0; k < n; ++k)
    ss += static_cast<int32_t>(tmp[(static_cast<int64_t>(tmp_row(d, y0 + k)) * S + xx) * 3 + c]) * kk[k];
  return normalise_pixel(clip8(ss), mean, std);
}

// ---- the forms the fast kernels use (same arithmetic, different data movement)
// Byte i (= x * 3 + c) of canvas row y: what the horizontal kernel stages into shared memory.
OPADPO_HD uint8_t canvas_byte(const uint8_t* img, const ImageDesc& d, int y, int i, const uint8_t* bg) {
  return static_cast<uint8_t>(canvas_pixel(img, d, i / 3, y, i % 3, bg));
}
// The same byte with the row already resolved: src = first byte of image row (y - oy), or null for a background-only row;
// [lo, hi) = canvas bytes covered by the image. One range test per byte, the modulo only outside the image.
OPADPO_HD uint8_t canvas_byte_row(const uint8_t* src, int lo, int hi, int i, const uint8_t* bg) {
  return (src && i >= lo && i < hi) ? src[i - lo] : bg[i % 3];
}
// Horizontal pass for output column xx, all three channels, from a staged canvas row: the coefficient is loaded once per
// tap and applied to the three interleaved bytes.
OPADPO_HD void resample_h_staged3(const uint8_t* row, const ImageDesc& d, const int32_t* tab, int S, int xx, uint8_t res[3]) {
  const int32_t* kk = tab + static_cast<int64_t>(xx) * d.ksx;
  const int32_t* bx = tab + static_cast<int64_t>(S) * d.ksx + 2 * xx;
  const int x0 = bx[0], n = bx[1];
  const uint8_t* px = row + x0 * 3;
  int32_t s0 = 1 << (kPrecisionBits - 1), s1 = s0, s2 = s0;
  for (int k = 0; k < n; ++k) {
    const int32_t c = kk[k];
    s0 += static_cast<int32_t>(px[3 * k]) * c;
    s1 += static_cast<int32_t>(px[3 * k + 1]) * c;
    s2 += static_cast<int32_t>(px[3 * k + 2]) * c;
  }
  res[0] = clip8(s0);
  res[1] = clip8(s1);
  res[2] = clip8(s2);
}
// Horizontal pass for output byte e (= xx * 3 + c) from a staged canvas row.
OPADPO_HD uint8_t resample_h_staged(const uint8_t* row, const ImageDesc& d, const int32_t* tab, int S, int e) {
  const int xx = e / 3, c = e - xx * 3;
  const int32_t* kk = tab + static_cast<int64_t>(xx) * d.ksx;
  const int32_t* bx = tab + static_cast<int64_t>(S) * d.ksx + 2 * xx;
  const int x0 = bx[0], n = bx[1];
  int32_t ss = 1 << (kPrecisionBits - 1);
  for (int k = 0; k < n; ++k) ss += static_cast<int32_t>(row[(x0 + k) * 3 + c]) * kk[k];
  return clip8(ss);
}
// Vertical pass for the four output bytes 4q .. 4q+3 of output row yy (needs S * 3 % 4 == 0 and 4-byte aligned rows):
// one 32-bit load per tap row instead of four byte loads. res[j] is the resampled byte, not yet normalised.
OPADPO_HD void resample_v_quad(const uint8_t* tmp, const ImageDesc& d, const int32_t* tab, int S, int yy, int q, uint8_t res[4]) {
  const int32_t* kky = tab + static_cast<int64_t>(S) * d.ksx + 2 * S;
  const int32_t* kk = kky + static_cast<int64_t>(yy) * d.ksy;
  const int32_t* by = kky + static_cast<int64_t>(S) * d.ksy + 2 * yy;
  const int y0 = by[0], n = by[1];
  int32_t s0 = 1 << (kPrecisionBits - 1), s1 = s0, s2 = s0, s3 = s0;
  for (int k = 0; k < n; ++k) {
    const uint32_t w = *reinterpret_cast<const uint32_t*>(tmp + static_cast<int64_t>(tmp_row(d, y0 + k)) * S * 3 + 4 * q);
    const int32_t c = kk[k];
    s0 += static_cast<int32_t>(w & 0xffu) * c;
    s1 += static_cast<int32_t>((w >> 8) & 0xffu) * c;
    s2 += static_cast<int32_t>((w >> 16) & 0xffu) * c;
    s3 += static_cast<int32_t>(w >> 24) * c;
  }
  res[0] = clip8(s0);
  res[1] = clip8(s1);
  res[2] = clip8(s2);
  res[3] = clip8(s3);
}

// float32 -> bf16 (round to nearest even) -> float32: the trainer masks images it has already cast to bf16.
OPADPO_HD float round_bf16(float x) {
  union {
    float f;
    uint32_t u;
  } v;
  v.f = x;
  v.u = (v.u + 0x7FFFu + ((v.u >> 16) & 1u)) & 0xFFFF0000u;
  return v.f;
}

}  // namespace collate
}  // namespace opadpo
