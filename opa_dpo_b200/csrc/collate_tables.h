// Host-side planning for the image-preprocess kernels: geometry of expand2square / shortest-edge resize / centre crop and
// Pillow's bicubic resampling tables (precompute_coeffs + normalize_coeffs_8bpc [UPSTREAM Pillow, ImagingResample]) in
// the exact double arithmetic Pillow uses, so the fixed-point coefficients are identical. Plain C++ (no CUDA): included
// by collate.cu and by the CPU test harness.
#pragma once
#include <cmath>
#include <cstdint>
#include <vector>

#include "collate_core.h"

namespace opadpo {
namespace collate {

inline double bicubic_filter(double x) {
  const double a = -0.5;
  if (x < 0.0) x = -x;
  if (x < 1.0) return ((a + 2.0) * x - (a + 3.0)) * x * x + 1;
  if (x < 2.0) return (((x - 5) * x + 8) * x - 4) * a;
  return 0.0;
}

// Appends kk [count, ksize] then bounds [count, 2] for output indices [first, first + count) of a resize of in_size
// samples to out_size samples. Returns ksize.
inline int append_resample_table(int in_size, int out_size, int first, int count, std::vector<int32_t>* out) {
  const double scale = static_cast<double>(in_size) / static_cast<double>(out_size);
  double filterscale = scale;
  if (filterscale < 1.0) filterscale = 1.0;
  const double support = 2.0 * filterscale;
  const int ksize = static_cast<int>(std::ceil(support)) * 2 + 1;
  const double ss = 1.0 / filterscale;
  const size_t base = out->size();
  out->resize(base + static_cast<size_t>(count) * ksize + 2 * static_cast<size_t>(count), 0);
  int32_t* kk = out->data() + base;
  int32_t* bounds = kk + static_cast<size_t>(count) * ksize;
  std::vector<double> w(ksize);
  for (int i = 0; i < count; ++i) {
    const int xx = first + i;
    const double center = (xx + 0.5) * scale;
    int xmin = static_cast<int>(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = static_cast<int>(center + support + 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    double ww = 0.0;
    for (int x = 0; x < xmax; ++x) {
      w[x] = bicubic_filter((x + xmin - center + 0.5) * ss);
      ww += w[x];
    }
    for (int x = 0; x < xmax; ++x) {
      double v = w[x];
      if (ww != 0.0) v /= ww;
      kk[static_cast<size_t>(i) * ksize + x] =
          v < 0 ? static_cast<int32_t>(-0.5 + v * (1 << kPrecisionBits)) : static_cast<int32_t>(0.5 + v * (1 << kPrecisionBits));
    }
    bounds[2 * i] = xmin;
    bounds[2 * i + 1] = xmax;
  }
  return ksize;
}

// Geometry + tables of one image (data_utils_dpo.py:311-333; get_resize_output_image_size / center_crop of the pinned
// image processor [UPSTREAM transformers 4.34.1]). tmp_bytes: size of its horizontal-pass scratch.
inline void plan_image(int H, int W, int S, bool pad_to_square, ImageDesc* d, std::vector<int32_t>* tables, int64_t* tmp_bytes) {
  d->H = H;
  d->W = W;
  d->Hs = H;
  d->Ws = W;
  d->ox = d->oy = 0;
  if (pad_to_square && W != H) {
    const int L = W > H ? W : H;
    d->Hs = d->Ws = L;
    if (W > H)
      d->oy = (W - H) / 2;
    else
      d->ox = (H - W) / 2;
  }
  const int w = d->Ws, h = d->Hs;
  const int short_ = w <= h ? w : h, long_ = w <= h ? h : w;
  const int new_long = static_cast<int>(static_cast<double>(static_cast<int64_t>(S) * long_) / static_cast<double>(short_));
  const int new_w = w <= h ? S : new_long, new_h = w <= h ? new_long : S;
  const int top = (new_h - S) / 2, left = (new_w - S) / 2;
  d->tab_off = static_cast<int64_t>(tables->size());
  d->ksx = append_resample_table(w, new_w, left, S, tables);
  d->ksy = append_resample_table(h, new_h, top, S, tables);
  *tmp_bytes = static_cast<int64_t>(h) * S * 3;
}

}  // namespace collate
}  // namespace opadpo
