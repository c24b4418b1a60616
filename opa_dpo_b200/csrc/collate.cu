// Collator-stage kernels. All of them are HBM-bound integer / byte passes: one read of the ragged
// input, one coalesced write of the padded output. Per-element logic is shared with the CPU test harness through
// collate_core.h; this file holds the thread mapping, the reductions and the launchers.
#include "collate.h"

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

// One CTA per response row. Pass 1 places every position (token_at walks the row's segments; a row has a few dozen), and
// finds the first pad / first EOS with shared-memory atomicMin; one thread then applies the EOS rules; pass 2 writes the
// attention mask from the final ids (or, MODE_TOKEN_MASK, pass 1 writes the tokenizer's own mask).
__global__ void __launch_bounds__(256) collate_responses_kernel(Segments s, int R, int pad_id, int eos_id, int mode,
                                                                int64_t* __restrict__ ids, uint8_t* __restrict__ mask,
                                                                float* __restrict__ scores, float* __restrict__ relations,
                                                                int32_t* __restrict__ overflow) {
  __shared__ int first_pad, first_eos;
  const int row = blockIdx.x;
  if (threadIdx.x == 0) {
    first_pad = R;
    first_eos = R;
  }
  __syncthreads();
  int64_t* ids_row = ids + static_cast<int64_t>(row) * R;
  float* sc_row = scores ? scores + static_cast<int64_t>(row) * R : nullptr;
  float* rel_row = relations ? relations + static_cast<int64_t>(row) * R : nullptr;
  uint8_t* mask_row = mask ? mask + static_cast<int64_t>(row) * R : nullptr;
  const bool token_mask = (mode & MODE_TOKEN_MASK) != 0;  // the tokenizer's mask: positions that hold a token
  for (int p = threadIdx.x; p < R; p += blockDim.x) {
    int32_t tok = 0;
    float sc = 0.f, rel = 0.f;
    const bool ok = token_at(s, row, p, &tok, &sc, &rel);
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
    const int64_t len = row_length(s, row);
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

// image.mean() of dpo_trainer.py:85: one CTA per image, float64 accumulation (the value is then closer to the exact mean
// than any float32 summation order; the reference's own order is not part of torch's contract).
__global__ void __launch_bounds__(1024) image_mean_kernel(const float* __restrict__ images, int64_t chw, int bf16_pixels,
                                                          float* __restrict__ means) {
  __shared__ double red[32];
  const float* img = images + static_cast<int64_t>(blockIdx.x) * chw;
  double acc = 0.0;
  for (int64_t i = threadIdx.x; i < chw; i += blockDim.x) {
    const float v = img[i];
    acc += static_cast<double>(bf16_pixels ? round_bf16(v) : v);
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    acc = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.0;
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
    if (threadIdx.x == 0) {
      float m = static_cast<float>(acc / static_cast<double>(chw));
      if (bf16_pixels) m = round_bf16(m);
      means[blockIdx.x] = m;
    }
  }
}

__global__ void copy_pixels_kernel(const float* __restrict__ in, float* __restrict__ out, int64_t n, int bf16_pixels) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float v = in[i];
    out[i] = bf16_pixels ? round_bf16(v) : v;
  }
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

struct Norm3 {
  float mean[3], std[3];
  uint8_t bg[4];
};

// Horizontal pass: grid (ceil(max Hs * S / 256), image); thread = (canvas row, output column), three channels each.
__global__ void __launch_bounds__(256) resample_h_kernel(const uint8_t* __restrict__ pixels, const ImageDesc* __restrict__ descs,
                                                         const int32_t* __restrict__ tables, uint8_t* __restrict__ tmp, int S,
                                                         Norm3 nm) {
  const ImageDesc d = descs[blockIdx.y];
  const int64_t item = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (item >= static_cast<int64_t>(d.Hs) * S) return;
  const int y = static_cast<int>(item / S), xx = static_cast<int>(item % S);
  const uint8_t* img = pixels + d.pix_off;
  const int32_t* tab = tables + d.tab_off;
  uint8_t* o = tmp + d.tmp_off + item * 3;
#pragma unroll
  for (int c = 0; c < 3; ++c) o[c] = resample_h(img, d, tab, S, y, xx, c, nm.bg);
}

// Vertical pass + rescale + normalise: grid (ceil(S * S / 256), image); planar f32 output, coalesced along x.
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

inline unsigned nblocks(int64_t n, int threads) { return static_cast<unsigned>((n + threads - 1) / threads); }

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
  const int64_t chw = static_cast<int64_t>(C) * H * W;
  float* means = means_out;
  if (!means) OPADPO_CUDA_CHECK(cudaMallocAsync(reinterpret_cast<void**>(&means), sizeof(float) * n, st));
  image_mean_kernel<<<n, 1024, 0, st>>>(images, chw, bf16_pixels, means);
  int rc = launch_ok("image_mean");
  if (rc == 0 && (out != images || bf16_pixels)) {
    const int64_t total = chw * n;
    unsigned blocks = nblocks(total, 256);
    if (blocks > 148u * 16u) blocks = 148u * 16u;
    copy_pixels_kernel<<<blocks, 256, 0, st>>>(images, out, total, bf16_pixels);
    rc = launch_ok("copy_pixels");
  }
  if (rc == 0 && m > 0) {
    const int64_t items = static_cast<int64_t>(m) * C * block * block;
    mask_fill_kernel<<<dim3(nblocks(items, 256), n), 256, 0, st>>>(out, means, idx, C, H, W, m, block);
    rc = launch_ok("mask_fill");
  }
  if (!means_out) cudaFreeAsync(means, st);
  return rc;
}

int attention_holes(uint8_t* out, const int64_t* idx, int n, int P, int m, cudaStream_t st) {
  if (n == 0) return 0;
  OPADPO_CUDA_CHECK(cudaMemsetAsync(out, 1, static_cast<size_t>(n) * P, st));
  if (m == 0) return 0;
  holes_kernel<<<nblocks(static_cast<int64_t>(n) * m, 256), 256, 0, st>>>(out, idx, n, P, m);
  return launch_ok("attention_holes");
}

int preprocess_images(const uint8_t* pixels, const int64_t* h_offsets, const int32_t* h_heights, const int32_t* h_widths, int n,
                      int S, const float* mean3, const float* std3, const uint8_t* bg3, int pad_to_square, float* out,
                      cudaStream_t st) {
  if (n == 0) return 0;
  std::vector<ImageDesc> descs(n);
  std::vector<int32_t> tables;
  int64_t tmp_total = 0;
  int max_hs = 0;
  for (int i = 0; i < n; ++i) {
    if (h_heights[i] <= 0 || h_widths[i] <= 0) {
      set_error("preprocess_images: image %d has size %dx%d", i, h_heights[i], h_widths[i]);
      return OPADPO_EINVAL;
    }
    int64_t tmp_bytes = 0;
    plan_image(h_heights[i], h_widths[i], S, pad_to_square != 0, &descs[i], &tables, &tmp_bytes);
    descs[i].pix_off = h_offsets[i];
    descs[i].tmp_off = tmp_total;
    tmp_total += (tmp_bytes + 15) & ~int64_t(15);
    if (descs[i].Hs > max_hs) max_hs = descs[i].Hs;
  }
  Norm3 nm;
  for (int c = 0; c < 3; ++c) {
    nm.mean[c] = mean3[c];
    nm.std[c] = std3[c];
    // expand2square's background: tuple(int(x * 255) for x in image_mean) (data_utils_dpo.py:328)
    nm.bg[c] = bg3 ? bg3[c] : static_cast<uint8_t>(static_cast<int>(static_cast<double>(mean3[c]) * 255));
  }
  nm.bg[3] = 0;
  ImageDesc* d_descs = nullptr;
  int32_t* d_tables = nullptr;
  uint8_t* d_tmp = nullptr;
  OPADPO_CUDA_CHECK(cudaMallocAsync(reinterpret_cast<void**>(&d_descs), sizeof(ImageDesc) * n, st));
  OPADPO_CUDA_CHECK(cudaMallocAsync(reinterpret_cast<void**>(&d_tables), sizeof(int32_t) * tables.size(), st));
  OPADPO_CUDA_CHECK(cudaMallocAsync(reinterpret_cast<void**>(&d_tmp), static_cast<size_t>(tmp_total), st));
  // pageable sources: cudaMemcpyAsync stages them before returning, so the vectors may go out of scope afterwards
  OPADPO_CUDA_CHECK(cudaMemcpyAsync(d_descs, descs.data(), sizeof(ImageDesc) * n, cudaMemcpyHostToDevice, st));
  OPADPO_CUDA_CHECK(cudaMemcpyAsync(d_tables, tables.data(), sizeof(int32_t) * tables.size(), cudaMemcpyHostToDevice, st));
  resample_h_kernel<<<dim3(nblocks(static_cast<int64_t>(max_hs) * S, 256), n), 256, 0, st>>>(pixels, d_descs, d_tables, d_tmp, S, nm);
  int rc = launch_ok("resample_h");
  if (rc == 0) {
    resample_v_kernel<<<dim3(nblocks(static_cast<int64_t>(S) * S, 256), n), 256, 0, st>>>(d_descs, d_tables, d_tmp, out, S, nm);
    rc = launch_ok("resample_v");
  }
  cudaFreeAsync(d_descs, st);
  cudaFreeAsync(d_tables, st);
  cudaFreeAsync(d_tmp, st);
  return rc;
}

}  // namespace opadpo
