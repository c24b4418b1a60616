// CPU TEST HARNESS (test infrastructure, never shipped, never loaded by opa_dpo_b200): runs the collator kernels'
// shared per-element core (opa_dpo_b200/csrc/collate_core.h, collate_tables.h) in plain loops that mirror the CUDA thread
// mapping of collate.cu, so the exact arithmetic the GPU executes can be checked against the oracle on a machine
// without a GPU. Built by tests/test_collate_cpu.py with g++.
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../opa_dpo_b200/csrc/collate_core.h"
#include "../../opa_dpo_b200/csrc/collate_tables.h"

using namespace opadpo::collate;

extern "C" {

void host_collate_responses(const int32_t* tokens, const int64_t* seg_offsets, const float* seg_score, const float* seg_relation,
                            const uint8_t* seg_flags, const int64_t* row_offsets, int b, int R, int pad_id, int eos_id, int mode,
                            int64_t* ids, uint8_t* mask, float* scores, float* relations, int32_t* overflow) {
  Segments s{tokens, seg_offsets, seg_score, seg_relation, seg_flags, row_offsets};
  for (int row = 0; row < b; ++row) {
    int first_pad = R, first_eos = R;
    int64_t* ids_row = ids + static_cast<int64_t>(row) * R;
    float* sc_row = scores ? scores + static_cast<int64_t>(row) * R : nullptr;
    float* rel_row = relations ? relations + static_cast<int64_t>(row) * R : nullptr;
    for (int p = 0; p < R; ++p) {
      int32_t tok = 0;
      float sc = 0.f, rel = 0.f;
      const bool ok = token_at(s, row, p, &tok, &sc, &rel);
      const int64_t id = ok ? tok : pad_id;
      ids_row[p] = id;
      if (sc_row) sc_row[p] = ok ? sc : 0.f;
      if (rel_row) rel_row[p] = ok ? rel : 0.f;
      if (mask && (mode & MODE_TOKEN_MASK)) mask[static_cast<int64_t>(row) * R + p] = ok ? 1 : 0;
      if (id == pad_id && p < first_pad) first_pad = p;
      if (id == eos_id && p < first_eos) first_eos = p;
    }
    finish_row(ids_row, sc_row, rel_row, R, first_pad, first_eos, mode, eos_id);
    const int64_t len = row_length(s, row);
    if (overflow) overflow[row] = (!(mode & MODE_TRUNCATE) && len > R) ? 1 : 0;
    if (mask && (mode & MODE_TOKEN_MASK) && (mode & MODE_ADD_EOS) && len < R) mask[static_cast<int64_t>(row) * R + len] = static_cast<uint8_t>(eos_id);
    if (mask && !(mode & MODE_TOKEN_MASK))
      for (int p = 0; p < R; ++p) mask[static_cast<int64_t>(row) * R + p] = ids_row[p] != pad_id ? 1 : 0;
  }
}

void host_collate_queries(const int32_t* tokens, const int64_t* offsets, int b, int Q, int pad_id, int placeholder_id,
                          int image_token, int64_t* ids, uint8_t* mask) {
  for (int row = 0; row < b; ++row)
    for (int p = 0; p < Q; ++p)
      query_at(tokens, offsets, row, p, Q, pad_id, placeholder_id, image_token, &ids[static_cast<int64_t>(row) * Q + p],
               &mask[static_cast<int64_t>(row) * Q + p]);
}

// fast = 0: the generic per-tap bounds-checked forms; fast = 1: the forms the staged / quad kernels use (canvas row staged
// in a buffer, four bytes per 32-bit load). Rows the kernels skip (all-background rows but the first) are poisoned.
void host_preprocess_images(const uint8_t* pixels, const int64_t* offsets, const int32_t* heights, const int32_t* widths, int n,
                            int S, const float* mean3, const float* std3, const uint8_t* bg3, int pad_to_square, int fast,
                            float* out) {
  for (int i = 0; i < n; ++i) {
    ImageDesc d{};
    std::vector<int32_t> tables;
    int64_t tmp_bytes = 0;
    plan_image(heights[i], widths[i], S, pad_to_square != 0, &d, &tables, &tmp_bytes);
    d.pix_off = offsets[i];
    d.tmp_off = 0;
    std::vector<uint32_t> tmp_words((static_cast<size_t>(tmp_bytes) + 3) / 4, 0xAAAAAAAAu);   // 4-byte aligned
    uint8_t* tmp = reinterpret_cast<uint8_t*>(tmp_words.data());
    const uint8_t* img = pixels + d.pix_off;
    const int32_t* tab = tables.data() + d.tab_off;
    std::vector<uint8_t> row(static_cast<size_t>(d.Ws) * 3);
    for (int y = 0; y < d.Hs; ++y) {
      if (!h_row_needed(d, y)) continue;
      if (fast == 1) {
        const uint8_t* src = is_padding_row(d, y) ? nullptr : img + static_cast<int64_t>(y - d.oy) * d.W * 3;
        for (int b = 0; b < d.Ws * 3; ++b) row[b] = canvas_byte_row(src, d.ox * 3, (d.ox + d.W) * 3, b, bg3);
        for (int xx = 0; xx < S; ++xx) resample_h_staged3(row.data(), d, tab, S, xx, tmp + (static_cast<int64_t>(y) * S + xx) * 3);
      } else if (fast == 2) {
        for (int b = 0; b < d.Ws * 3; ++b) row[b] = canvas_byte(img, d, y, b, bg3);
        for (int e = 0; e < S * 3; ++e) tmp[static_cast<int64_t>(y) * S * 3 + e] = resample_h_staged(row.data(), d, tab, S, e);
      } else {
        for (int xx = 0; xx < S; ++xx)
          for (int c = 0; c < 3; ++c) tmp[(static_cast<int64_t>(y) * S + xx) * 3 + c] = resample_h(img, d, tab, S, y, xx, c, bg3);
      }
    }
    for (int yy = 0; yy < S; ++yy) {
      if (fast && (S * 3) % 4 == 0) {
        for (int q = 0; q < S * 3 / 4; ++q) {
          uint8_t r[4];
          resample_v_quad(tmp, d, tab, S, yy, q, r);
          for (int j = 0; j < 4; ++j) {
            const int e = 4 * q + j, xx = e / 3, c = e - xx * 3;
            out[((static_cast<int64_t>(i) * 3 + c) * S + yy) * S + xx] = normalise_pixel(r[j], mean3[c], std3[c]);
          }
        }
      } else {
        for (int xx = 0; xx < S; ++xx)
          for (int c = 0; c < 3; ++c)
            out[((static_cast<int64_t>(i) * 3 + c) * S + yy) * S + xx] = resample_v(tmp, d, tab, S, yy, xx, c, mean3[c], std3[c]);
      }
    }
  }
}

// the truncate kernel's logic: first occurrences (what the CTA finds with atomicMin), then the shared decision function
void host_truncate_after_eos(const int64_t* in, int64_t* out, int b, int R, int64_t eos, int64_t pad, const int64_t* extra,
                             int n_extra) {
  for (int r = 0; r < b; ++r) {
    int first[1 + kMaxStopTokens];
    for (int k = 0; k <= n_extra; ++k) first[k] = R;
    const int64_t* row = in + static_cast<int64_t>(r) * R;
    for (int p = R - 1; p >= 0; --p) {   // any visiting order gives the minimum
      if (row[p] == eos) first[0] = p;
      for (int k = 0; k < n_extra; ++k)
        if (row[p] == extra[k]) first[1 + k] = p;
    }
    const int end = truncate_end_from_firsts(first, n_extra, R);
    for (int p = 0; p < R; ++p) out[static_cast<int64_t>(r) * R + p] = (end >= 0 && p > end) ? pad : row[p];
  }
}

float host_round_bf16(float x) { return round_bf16(x); }

}  // extern "C"
