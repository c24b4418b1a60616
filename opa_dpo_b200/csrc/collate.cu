// Collator-stage kernels. All of them are HBM-bound integer / byte passes: one read of the ragged
// input, one coalesced write of the padded output. Per-element logic is shared with the CPU test harness through
// collate_core.h; this file holds the thread mapping, the reductions and the launchers.
#include "collate.h"

#include <map>
#include <mutex>
#include <tuple>
#include <vector>

#include "collate_tables.h"
#include "common.h"

namespace opadpo {
namespace {

using namespace collate;

inline int launch_ok(const char* what) {
  count_launch();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s launch: %s", what, cudaGetErrorString(e));
    return OPADPO_ECUDA;
  }
  return 0;
}

// One CTA per response row. The row's segments are resolved once into shared memory (effective begin + running length;
// rows with more than kMaxSegs entries fall back to walking the segments per position), every position then finds its
// segment by binary search; first pad / first EOS come from shared-memory atomicMin; one thread applies the EOS rules;
// pass 2 writes the attention mask from the final ids (or, MODE_TOKEN_MASK, pass 1 writes the tokenizer's own mask).
constexpr int kMaxSegs = 512;
__global__ void __launch_bounds__(256) collate_responses_kernel(Segments s, int R, int pad_id, int eos_id, int mode,
                                                                int64_t* __restrict__ ids, uint8_t* __restrict__ mask,
                                                                float* __restrict__ scores, float* __restrict__ relations,
                                                                int32_t* __restrict__ overflow) {
  __shared__ int first_pad, first_eos;
  __shared__ int64_t seg_begin[kMaxSegs];
  __shared__ int64_t seg_cum[kMaxSegs + 1];  // seg_cum[j] = tokens of the row before segment j
  const int row = blockIdx.x;
  const int64_t j0 = s.row_offsets[row];
  const int nseg = static_cast<int>(s.row_offsets[row + 1] - j0);
  const bool indexed = nseg <= kMaxSegs;
  if (threadIdx.x == 0) {
    first_pad = R;
    first_eos = R;
  }
  if (indexed) {
    for (int j = threadIdx.x; j < nseg; j += blockDim.x) {
      int64_t b, e;
      segment_range(s, j0 + j, &b, &e);
      seg_begin[j] = b;
      seg_cum[j + 1] = e - b;  // lengths first, turned into a running sum below
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      seg_cum[0] = 0;
      for (int j = 0; j < nseg; ++j) seg_cum[j + 1] += seg_cum[j];
    }
  }
  __syncthreads();
  const int64_t len = indexed ? seg_cum[nseg] : row_length(s, row);
  int64_t* ids_row = ids + static_cast<int64_t>(row) * R;
  float* sc_row = scores ? scores + static_cast<int64_t>(row) * R : nullptr;
  float* rel_row = relations ? relations + static_cast<int64_t>(row) * R : nullptr;
  uint8_t* mask_row = mask ? mask + static_cast<int64_t>(row) * R : nullptr;
  const bool token_mask = (mode & MODE_TOKEN_MASK) != 0;  // the tokenizer's mask: positions that hold a token
  for (int p = threadIdx.x; p < R; p += blockDim.x) {
    int32_t tok = 0;
    float sc = 0.f, rel = 0.f;
    bool ok;
    if (indexed) {
      ok = p < len;
      if (ok) {  // largest j with seg_cum[j] <= p (empty segments share a boundary: the search lands on the non-empty one)
        int lo = 0, hi = nseg - 1;
        while (lo < hi) {
          const int mid = (lo + hi + 1) >> 1;
          if (seg_cum[mid] <= p)
            lo = mid;
          else
            hi = mid - 1;
        }
        tok = s.tokens[seg_begin[lo] + (p - seg_cum[lo])];
        sc = s.seg_score ? s.seg_score[j0 + lo] : 0.f;
        rel = s.seg_relation ? s.seg_relation[j0 + lo] : 0.f;
      }
    } else {
      ok = token_at(s, row, p, &tok, &sc, &rel);
    }
    const int64_t id = ok ? tok : pad_id;
    ids_row[p] = id;
    if (sc_row) sc_row[p] = ok ? sc : 0.f;
    if (rel_row) rel_row[p] = ok ? rel : 0.f;
    if (mask_row && token_mask) mask_row[p] = ok ? 1 : 0;
    if (id == pad_id) atomicMin(&first_pad, p);
    if (id == eos_id) atomicMin(&first_eos, p);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    finish_row(ids_row, sc_row, rel_row, R, first_pad, first_eos, mode, eos_id);
    if (overflow) overflow[row] = (!(mode & MODE_TRUNCATE) && len > R) ? 1 : 0;
    // add_eos also walks the tokenizer's attention mask (data_utils_dpo.py:76-79): its first 0 becomes eos_id
    if (mask_row && token_mask && (mode & MODE_ADD_EOS) && len < R) mask_row[len] = static_cast<uint8_t>(eos_id);
  }
  __syncthreads();
  if (mask_row && !token_mask)
    for (int p = threadIdx.x; p < R; p += blockDim.x) mask_row[p] = ids_row[p] != pad_id ? 1 : 0;
}

__global__ void collate_queries_kernel(const int32_t* __restrict__ tokens, const int64_t* __restrict__ offsets, int b, int Q,
                                       int pad_id, int placeholder_id, int image_token, int64_t* __restrict__ ids,
                                       uint8_t* __restrict__ mask) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<int64_t>(b) * Q) return;
  const int row = static_cast<int>(i / Q), p = static_cast<int>(i % Q);
  int64_t id;
  uint8_t m;
  query_at(tokens, offsets, row, p, Q, pad_id, placeholder_id, image_token, &id, &m);
  ids[i] = id;
  if (mask) mask[i] = m;
}

// image.mean() of dpo_trainer.py:85, fused with the copy into `out`: grid (kMeanParts, image); each CTA reads its slice of
// the image once, writes it to `out` (bf16-rounded when the trainer's images are bf16) and leaves one float64 partial
// sum; the partials are combined in a fixed order (deterministic), in float64 — closer to the exact mean than any
// float32 summation order, and the reference's own order is not part of torch's contract.
constexpr int kMeanParts = 16;
__global__ void __launch_bounds__(512) image_sum_copy_kernel(const float* __restrict__ images, float* __restrict__ out, int64_t chw,
                                                             int bf16_pixels, int do_copy, double* __restrict__ partials) {
  __shared__ double red[16];
  const int64_t per = (chw + kMeanParts - 1) / kMeanParts;
  const int64_t begin = blockIdx.x * per, end = begin + per < chw ? begin + per : chw;
  const float* img = images + static_cast<int64_t>(blockIdx.y) * chw;
  float* dst = out + static_cast<int64_t>(blockIdx.y) * chw;
  double acc = 0.0;
  for (int64_t i = begin + threadIdx.x; i < end; i += blockDim.x) {
    float v = img[i];
    if (bf16_pixels) v = round_bf16(v);
    if (do_copy) dst[i] = v;
    acc += static_cast<double>(v);
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    acc = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.0;
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
    if (threadIdx.x == 0) partials[static_cast<int64_t>(blockIdx.y) * kMeanParts + blockIdx.x] = acc;
  }
}

__global__ void image_mean_finish_kernel(const double* __restrict__ partials, int n, int64_t chw, int bf16_pixels,
                                         float* __restrict__ means) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double acc = 0.0;
  for (int j = 0; j < kMeanParts; ++j) acc += partials[static_cast<int64_t>(i) * kMeanParts + j];
  float m = static_cast<float>(acc / static_cast<double>(chw));
  if (bf16_pixels) m = round_bf16(m);
  means[i] = m;
}

// grid (items, image): item = ((k * C + c) * block + dy) * block + dx
__global__ void mask_fill_kernel(float* __restrict__ out, const float* __restrict__ means, const int64_t* __restrict__ idx, int C,
                                 int H, int W, int m, int block) {
  const int img = blockIdx.y;
  const int64_t item = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t per_k = static_cast<int64_t>(C) * block * block;
  if (item >= per_k * m) return;
  const int k = static_cast<int>(item / per_k);
  int rem = static_cast<int>(item % per_k);
  const int c = rem / (block * block);
  rem %= block * block;
  const int dy = rem / block, dx = rem % block;
  const int64_t sel = idx[static_cast<int64_t>(img) * m + k];
  int y, x;
  if (block == 1) {
    if (sel < 0 || sel >= static_cast<int64_t>(H) * W) return;
    y = static_cast<int>(sel / W);
    x = static_cast<int>(sel % W);
  } else {
    const int wb = W / block, hb = H / block;
    if (sel < 0 || sel >= static_cast<int64_t>(hb) * wb) return;
    y = static_cast<int>(sel / wb) * block + dy;
    x = static_cast<int>(sel % wb) * block + dx;
  }
  out[((static_cast<int64_t>(img) * C + c) * H + y) * W + x] = means[img];
}

__global__ void holes_kernel(uint8_t* __restrict__ out, const int64_t* __restrict__ idx, int n, int P, int m) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<int64_t>(n) * m) return;
  const int row = static_cast<int>(i / m);
  const int64_t col = idx[i];
  if (col >= 0 && col < P) out[static_cast<int64_t>(row) * P + col] = 0;
}

struct StopTokens {
  int64_t eos;
  int64_t extra[kMaxStopTokens];
  int n_extra;
};

// One CTA per rollout row: first occurrence of every stop token by shared-memory atomicMin, then the tail goes to pad.
__global__ void __launch_bounds__(128) truncate_after_eos_kernel(const int64_t* __restrict__ in, int64_t* __restrict__ out, int R,
                                                                 int64_t pad_id, StopTokens st) {
  __shared__ int first[1 + kMaxStopTokens];
  __shared__ int end_s;
  if (threadIdx.x <= st.n_extra) first[threadIdx.x] = R;
  __syncthreads();
  const int64_t* row = in + static_cast<int64_t>(blockIdx.x) * R;
  int64_t* orow = out + static_cast<int64_t>(blockIdx.x) * R;
  for (int p = threadIdx.x; p < R; p += blockDim.x) {
    const int64_t t = row[p];
    if (t == st.eos) atomicMin(&first[0], p);
    for (int k = 0; k < st.n_extra; ++k)
      if (t == st.extra[k]) atomicMin(&first[1 + k], p);
  }
  __syncthreads();
  if (threadIdx.x == 0) end_s = truncate_end_from_firsts(first, st.n_extra, R);
  __syncthreads();
  const int end = end_s;
  for (int p = threadIdx.x; p < R; p += blockDim.x) {
    const int64_t t = row[p];
    orow[p] = (end >= 0 && p > end) ? pad_id : t;
  }
}

struct Norm3 {
  float mean[3], std[3];
  uint8_t bg[4];
};

// Horizontal pass, generic form: grid (ceil(max Hs * S / 256), image); thread = (canvas row, output column), three
// channels each, every tap bounds-checked against the canvas. Used when a canvas row does not fit shared memory.
__global__ void __launch_bounds__(256) resample_h_kernel(const uint8_t* __restrict__ pixels, const ImageDesc* __restrict__ descs,
                                                         const int32_t* __restrict__ tables, uint8_t* __restrict__ tmp, int S,
                                                         Norm3 nm) {
  const ImageDesc d = descs[blockIdx.y];
  const int64_t item = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (item >= static_cast<int64_t>(d.Hs) * S) return;
  const int y = static_cast<int>(item / S), xx = static_cast<int>(item % S);
  if (!h_row_needed(d, y)) return;
  const uint8_t* img = pixels + d.pix_off;
  const int32_t* tab = tables + d.tab_off;
  uint8_t* o = tmp + d.tmp_off + item * 3;
#pragma unroll
  for (int c = 0; c < 3; ++c) o[c] = resample_h(img, d, tab, S, y, xx, c, nm.bg);
}

// Horizontal pass, staged form: one CTA per canvas row (grid (max Hs, image), 128 threads). The row — image bytes with
// the background colour left and right of them — is copied into shared memory with coalesced loads, so the tap loop has
// no bounds checks; each thread then produces the three channels of output columns xx, xx + 128, ... with one
// coefficient load per tap. Background-only rows are done once per image (h_row_needed).
constexpr int kHThreads = 128;
__global__ void __launch_bounds__(kHThreads) resample_h_staged_kernel(const uint8_t* __restrict__ pixels,
                                                                      const ImageDesc* __restrict__ descs,
                                                                      const int32_t* __restrict__ tables,
                                                                      uint8_t* __restrict__ tmp, int S, Norm3 nm) {
  extern __shared__ uint8_t canvas_row[];
  const ImageDesc d = descs[blockIdx.y];
  const int y = blockIdx.x;
  if (y >= d.Hs || !h_row_needed(d, y)) return;
  const uint8_t* src = is_padding_row(d, y) ? nullptr : pixels + d.pix_off + static_cast<int64_t>(y - d.oy) * d.W * 3;
  const int lo = d.ox * 3, hi = (d.ox + d.W) * 3, row_bytes = d.Ws * 3;
  for (int i = threadIdx.x; i < row_bytes; i += blockDim.x) canvas_row[i] = canvas_byte_row(src, lo, hi, i, nm.bg);
  __syncthreads();
  const int32_t* tab = tables + d.tab_off;
  uint8_t* o = tmp + d.tmp_off + static_cast<int64_t>(y) * S * 3;
  for (int xx = threadIdx.x; xx < S; xx += blockDim.x) {
    uint8_t r[3];
    resample_h_staged3(canvas_row, d, tab, S, xx, r);
    o[xx * 3] = r[0];
    o[xx * 3 + 1] = r[1];
    o[xx * 3 + 2] = r[2];
  }
}

// Vertical pass + rescale + normalise, generic form: grid (ceil(S * S / 256), image); planar f32 output.
__global__ void __launch_bounds__(256) resample_v_kernel(const ImageDesc* __restrict__ descs, const int32_t* __restrict__ tables,
                                                         const uint8_t* __restrict__ tmp, float* __restrict__ out, int S,
                                                         Norm3 nm) {
  const ImageDesc d = descs[blockIdx.y];
  const int item = blockIdx.x * blockDim.x + threadIdx.x;
  if (item >= S * S) return;
  const int yy = item / S, xx = item % S;
  const int32_t* tab = tables + d.tab_off;
  const uint8_t* t = tmp + d.tmp_off;
#pragma unroll
  for (int c = 0; c < 3; ++c)
    out[(static_cast<int64_t>(blockIdx.y) * 3 + c) * S * S + item] = resample_v(t, d, tab, S, yy, xx, c, nm.mean[c], nm.std[c]);
}

// Vertical pass, quad form (S * 3 % 4 == 0): one CTA per output row (grid (S, image)); a thread resamples four adjacent
// bytes with one 32-bit load per tap row, the row's 3*S normalised values are transposed through shared memory so each
// colour plane is written with fully coalesced stores.
__global__ void __launch_bounds__(256) resample_v_quad_kernel(const ImageDesc* __restrict__ descs,
                                                              const int32_t* __restrict__ tables,
                                                              const uint8_t* __restrict__ tmp, float* __restrict__ out, int S,
                                                              Norm3 nm) {
  extern __shared__ float planes[];  // [3][S]
  const ImageDesc d = descs[blockIdx.y];
  const int yy = blockIdx.x;
  const int32_t* tab = tables + d.tab_off;
  const uint8_t* t = tmp + d.tmp_off;
  for (int q = threadIdx.x; q < S * 3 / 4; q += blockDim.x) {
    uint8_t r[4];
    resample_v_quad(t, d, tab, S, yy, q, r);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int e = 4 * q + j, xx = e / 3, c = e - xx * 3;
      planes[c * S + xx] = normalise_pixel(r[j], nm.mean[c], nm.std[c]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 3 * S; i += blockDim.x) {
    const int c = i / S, xx = i - c * S;
    out[((static_cast<int64_t>(blockIdx.y) * 3 + c) * S + yy) * S + xx] = planes[i];
  }
}

inline unsigned nblocks(int64_t n, int threads) { return static_cast<unsigned>((n + threads - 1) / threads); }

// The per-call scratch (cudaMallocAsync) must stay in the device's pool between calls: with the default release
// threshold of 0 every synchronisation hands it back to the driver and the next call pays a fresh mapping (~0.1 ms per
// allocation, several times the kernels). The engine does the same at init; the collator ops work without an engine.
void keep_pool_memory() {
  static std::once_flag once[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return;
  std::call_once(once[dev], [dev]() {
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
      uint64_t thr = UINT64_MAX;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
  });
}

}  // namespace

int collate_responses(const collate::Segments& s, int b, int R, int pad_id, int eos_id, int mode, int64_t* ids, uint8_t* mask,
                      float* scores, float* relations, int32_t* overflow, cudaStream_t st) {
  if (b == 0) return 0;
  collate_responses_kernel<<<b, 256, 0, st>>>(s, R, pad_id, eos_id, mode, ids, mask, scores, relations, overflow);
  return launch_ok("collate_responses");
}

int collate_queries(const int32_t* tokens, const int64_t* offsets, int b, int Q, int pad_id, int placeholder_id, int image_token,
                    int64_t* ids, uint8_t* mask, cudaStream_t st) {
  if (b == 0) return 0;
  collate_queries_kernel<<<nblocks(static_cast<int64_t>(b) * Q, 256), 256, 0, st>>>(tokens, offsets, b, Q, pad_id, placeholder_id,
                                                                                   image_token, ids, mask);
  return launch_ok("collate_queries");
}

int mask_images(const float* images, float* out, int n, int C, int H, int W, const int64_t* idx, int m, int block, int bf16_pixels,
                float* means_out, cudaStream_t st) {
  if (n == 0) return 0;
  keep_pool_memory();
  const int64_t chw = static_cast<int64_t>(C) * H * W;
  float* means = means_out;
  double* partials = nullptr;
  if (!means) OPADPO_CUDA_CHECK(cudaMallocAsync(reinterpret_cast<void**>(&means), sizeof(float) * n, st));
  OPADPO_CUDA_CHECK(cudaMallocAsync(reinterpret_cast<void**>(&partials), sizeof(double) * n * kMeanParts, st));
  const int do_copy = (out != images || bf16_pixels) ? 1 : 0;
  image_sum_copy_kernel<<<dim3(kMeanParts, n), 512, 0, st>>>(images, out, chw, bf16_pixels, do_copy, partials);
  int rc = launch_ok("image_sum_copy");
  if (rc == 0) {
    image_mean_finish_kernel<<<nblocks(n, 128), 128, 0, st>>>(partials, n, chw, bf16_pixels, means);
    rc = launch_ok("image_mean_finish");
  }
  if (rc == 0 && m > 0) {
    const int64_t items = static_cast<int64_t>(m) * C * block * block;
    mask_fill_kernel<<<dim3(nblocks(items, 256), n), 256, 0, st>>>(out, means, idx, C, H, W, m, block);
    rc = launch_ok("mask_fill");
  }
  cudaFreeAsync(partials, st);
  if (!means_out) cudaFreeAsync(means, st);
  return rc;
}

int truncate_after_eos(const int64_t* in, int64_t* out, int b, int R, int64_t eos_id, int64_t pad_id, const int64_t* h_extra,
                       int n_extra, cudaStream_t st) {
  if (b == 0) return 0;
  if (n_extra < 0 || n_extra > kMaxStopTokens) {
    set_error("truncate_after_eos: at most %d additional stop tokens", kMaxStopTokens);
    return OPADPO_EINVAL;
  }
  StopTokens s{};
  s.eos = eos_id;
  s.n_extra = n_extra;
  for (int k = 0; k < n_extra; ++k) s.extra[k] = h_extra[k];
  truncate_after_eos_kernel<<<b, 128, 0, st>>>(in, out, R, pad_id, s);
  return launch_ok("truncate_after_eos");
}

int attention_holes(uint8_t* out, const int64_t* idx, int n, int P, int m, cudaStream_t st) {
  if (n == 0) return 0;
  OPADPO_CUDA_CHECK(cudaMemsetAsync(out, 1, static_cast<size_t>(n) * P, st));
  if (m == 0) return 0;
  holes_kernel<<<nblocks(static_cast<int64_t>(n) * m, 256), 256, 0, st>>>(out, idx, n, P, m);
  return launch_ok("attention_holes");
}

// Resampling tables depend only on (H, W, S, pad_to_square) and a dataset has a handful of distinct image sizes, so each
// geometry is planned once (Pillow's double arithmetic, ~30 us of host time) and its tables stay in a per-device arena;
// a call then only uploads the n image descriptors.
namespace {
struct PlanKey {
  int dev, H, W, S, pad;
  bool operator<(const PlanKey& o) const {
    return std::tie(dev, H, W, S, pad) < std::tie(o.dev, o.H, o.W, o.S, o.pad);
  }
};
struct TableArena {
  int32_t* base = nullptr;
  int64_t used = 0;
};
constexpr int64_t kArenaInts = 8 * 1024 * 1024;  // 32 MB of coefficients per device
std::mutex g_plan_mu;
std::map<PlanKey, ImageDesc> g_plans;
std::map<int, TableArena> g_arenas;

int cached_plan(int dev, int H, int W, int S, bool pad, cudaStream_t st, ImageDesc* out) {
  std::lock_guard<std::mutex> lock(g_plan_mu);
  const PlanKey key{dev, H, W, S, pad ? 1 : 0};
  auto it = g_plans.find(key);
  if (it != g_plans.end()) {
    *out = it->second;
    return 0;
  }
  TableArena& ar = g_arenas[dev];
  if (!ar.base) OPADPO_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&ar.base), sizeof(int32_t) * kArenaInts));
  ImageDesc d{};
  std::vector<int32_t> tables;
  int64_t tmp_bytes = 0;
  plan_image(H, W, S, pad, &d, &tables, &tmp_bytes);
  const int64_t need = static_cast<int64_t>(tables.size());
  if (need > kArenaInts) {
    set_error("preprocess_images: a %dx%d image needs %lld table entries (arena holds %lld)", H, W, (long long)need,
              (long long)kArenaInts);
    return OPADPO_EINVAL;
  }
  if (ar.used + need > kArenaInts) {  // arena full: wait for everything that may still read it, then start over
    OPADPO_CUDA_CHECK(cudaDeviceSynchronize());
    for (auto p = g_plans.begin(); p != g_plans.end();) p = p->first.dev == dev ? g_plans.erase(p) : std::next(p);
    ar.used = 0;
  }
  d.tab_off = ar.used;
  // blocking copy (first use of a geometry only): the tables must be visible to whichever stream uses the plan next
  OPADPO_CUDA_CHECK(cudaMemcpy(ar.base + ar.used, tables.data(), sizeof(int32_t) * need, cudaMemcpyHostToDevice));
  (void)st;
  ar.used += need;
  g_plans[key] = d;
  *out = d;
  return 0;
}
}  // namespace

int preprocess_images(const uint8_t* pixels, const int64_t* h_offsets, const int32_t* h_heights, const int32_t* h_widths, int n,
                      int S, const float* mean3, const float* std3, const uint8_t* bg3, int pad_to_square, float* out,
                      cudaStream_t st) {
  if (n == 0) return 0;
  keep_pool_memory();
  int dev = 0;
  OPADPO_CUDA_CHECK(cudaGetDevice(&dev));
  std::vector<ImageDesc> descs(n);
  int64_t tmp_total = 0;
  int max_hs = 0, max_ws = 0;
  for (int i = 0; i < n; ++i) {
    if (h_heights[i] <= 0 || h_widths[i] <= 0) {
      set_error("preprocess_images: image %d has size %dx%d", i, h_heights[i], h_widths[i]);
      return OPADPO_EINVAL;
    }
    int rc = cached_plan(dev, h_heights[i], h_widths[i], S, pad_to_square != 0, st, &descs[i]);
    if (rc != 0) return rc;
    descs[i].pix_off = h_offsets[i];
    descs[i].tmp_off = tmp_total;
    tmp_total += (static_cast<int64_t>(descs[i].Hs) * S * 3 + 15) & ~int64_t(15);
    if (descs[i].Hs > max_hs) max_hs = descs[i].Hs;
    if (descs[i].Ws > max_ws) max_ws = descs[i].Ws;
  }
  Norm3 nm;
  for (int c = 0; c < 3; ++c) {
    nm.mean[c] = mean3[c];
    nm.std[c] = std3[c];
    // expand2square's background: tuple(int(x * 255) for x in image_mean) (data_utils_dpo.py:328)
    nm.bg[c] = bg3 ? bg3[c] : static_cast<uint8_t>(static_cast<int>(static_cast<double>(mean3[c]) * 255));
  }
  nm.bg[3] = 0;
  const int32_t* d_tables = nullptr;
  {
    std::lock_guard<std::mutex> lock(g_plan_mu);
    d_tables = g_arenas[dev].base;
  }
  ImageDesc* d_descs = nullptr;
  uint8_t* d_tmp = nullptr;
  OPADPO_CUDA_CHECK(cudaMallocAsync(reinterpret_cast<void**>(&d_descs), sizeof(ImageDesc) * n, st));
  OPADPO_CUDA_CHECK(cudaMallocAsync(reinterpret_cast<void**>(&d_tmp), static_cast<size_t>(tmp_total), st));
  // pageable source: cudaMemcpyAsync stages it before returning, so the vector may go out of scope afterwards
  OPADPO_CUDA_CHECK(cudaMemcpyAsync(d_descs, descs.data(), sizeof(ImageDesc) * n, cudaMemcpyHostToDevice, st));
  int rc;
  if (max_ws * 3 <= 48 * 1024) {
    resample_h_staged_kernel<<<dim3(max_hs, n), kHThreads, max_ws * 3, st>>>(pixels, d_descs, d_tables, d_tmp, S, nm);
    rc = launch_ok("resample_h_staged");
  } else {
    resample_h_kernel<<<dim3(nblocks(static_cast<int64_t>(max_hs) * S, 256), n), 256, 0, st>>>(pixels, d_descs, d_tables, d_tmp, S, nm);
    rc = launch_ok("resample_h");
  }
  if (rc == 0) {
    if ((S * 3) % 4 == 0 && S * 3 * sizeof(float) <= 48 * 1024) {
      resample_v_quad_kernel<<<dim3(S, n), 256, S * 3 * sizeof(float), st>>>(d_descs, d_tables, d_tmp, out, S, nm);
      rc = launch_ok("resample_v_quad");
    } else {
      resample_v_kernel<<<dim3(nblocks(static_cast<int64_t>(S) * S, 256), n), 256, 0, st>>>(d_descs, d_tables, d_tmp, out, S, nm);
      rc = launch_ok("resample_v");
    }
  }
  cudaFreeAsync(d_descs, st);
  cudaFreeAsync(d_tmp, st);
  return rc;
}

}  // namespace opadpo
