// HBM-bound kernels of the OPA-DPO hot path (see elementwise.h). sm_100a; 16-byte vector accesses, fp32 math.
#include "elementwise.h"

#include <cmath>
#include <cstdio>

#include "common.h"

namespace opadpo {
namespace {

constexpr int kIgnore = -1;

__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float2 t = __bfloat1622float2(h[e]);
    f[2 * e] = t.x;
    f[2 * e + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int e = 0; e < 4; ++e) h[e] = __floats2bfloat162_rn(f[2 * e], f[2 * e + 1]);
  return u;
}
__device__ __forceinline__ float bf16_round(float x) { return __bfloat162float(__float2bfloat16(x)); }
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// Sum over the block (all threads get the result). red must hold 32 floats.
__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  float t = (lane < nw) ? red[lane] : 0.f;
  t = warp_sum(t);
  return t;
}

// Two sums in one pass through shared memory (same reduction order per quantity as block_sum). red must hold 64 floats.
__device__ __forceinline__ void block_sum2(float& a, float& b, float* red) {
  a = warp_sum(a);
  b = warp_sum(b);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) {
    red[warp] = a;
    red[32 + warp] = b;
  }
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  a = warp_sum((lane < nw) ? red[lane] : 0.f);
  b = warp_sum((lane < nw) ? red[32 + lane] : 0.f);
}

__device__ __forceinline__ int64_t map_row(int64_t r, const int32_t* idx) { return idx ? idx[r] : r; }

constexpr int NORM_THREADS = 256;
constexpr int NORM_MAXV = 4;  // up to 8192 columns held in registers

// ---------------------------------------------------------------------------------------------- RMSNorm
__global__ void __launch_bounds__(NORM_THREADS) rmsnorm_fwd_kernel(const bf16* __restrict__ x, const bf16* __restrict__ w,
                                                                    bf16* __restrict__ y, int d, float eps,
                                                                    const int32_t* __restrict__ x_idx) {
  __shared__ float red[32];
  const int64_t r = blockIdx.x;
  const uint4* xr = reinterpret_cast<const uint4*>(x + map_row(r, x_idx) * d);
  const int nv = d >> 3;
  uint4 xv[NORM_MAXV];
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < NORM_MAXV; ++i) {
    const int v = threadIdx.x + i * NORM_THREADS;
    if (v < nv) {
      xv[i] = xr[v];
      float f[8];
      unpack8(xv[i], f);
#pragma unroll
      for (int e = 0; e < 8; ++e) ss += f[e] * f[e];
    }
  }
  ss = block_sum(ss, red);
  const float rstd = rsqrtf(ss / d + eps);
  uint4* yr = reinterpret_cast<uint4*>(y + r * d);
  const uint4* wr = reinterpret_cast<const uint4*>(w);
#pragma unroll
  for (int i = 0; i < NORM_MAXV; ++i) {
    const int v = threadIdx.x + i * NORM_THREADS;
    if (v < nv) {
      float f[8], g[8];
      unpack8(xv[i], f);
      unpack8(__ldg(wr + v), g);
#pragma unroll
      for (int e = 0; e < 8; ++e) f[e] = g[e] * bf16_round(f[e] * rstd);
      yr[v] = pack8(f);
    }
  }
}

// NV = 16 B vectors per thread (d <= NV * 8 * NORM_THREADS): sized per width so the three row buffers stay in registers
template <int NV>
__global__ void __launch_bounds__(NORM_THREADS)
    rmsnorm_bwd_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ x, const bf16* __restrict__ w,
                       const bf16* dres, bf16* dx, int d, float eps, const int32_t* __restrict__ x_idx,
                       const int32_t* __restrict__ dx_idx) {
  __shared__ float red[64];
  const int64_t r = blockIdx.x;
  const int64_t xrw = map_row(r, x_idx);
  const int64_t fr = map_row(r, dx_idx);
  const uint4* xr = reinterpret_cast<const uint4*>(x + xrw * d);
  const uint4* dyr = reinterpret_cast<const uint4*>(dy + r * d);
  const uint4* wr = reinterpret_cast<const uint4*>(w);
  const uint4* drr = dres ? reinterpret_cast<const uint4*>(dres + fr * d) : nullptr;
  const int nv = d >> 3;
  uint4 xv[NV], gv[NV];  // gv = w*dy packed as bf16 is lossy; keep dy and re-multiply
  uint4 rv[NV];                 // the incoming residual gradient: requested now, needed after the row reductions
  float ss = 0.f, dot = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int v = threadIdx.x + i * NORM_THREADS;
    if (v < nv) {
      xv[i] = xr[v];
      gv[i] = dyr[v];
      rv[i] = drr ? drr[v] : make_uint4(0u, 0u, 0u, 0u);  // bf16 +0 bits
    }
  }
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int v = threadIdx.x + i * NORM_THREADS;
    if (v < nv) {
      float f[8], g[8], ww[8];
      unpack8(xv[i], f);
      unpack8(gv[i], g);
      unpack8(__ldg(wr + v), ww);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        ss += f[e] * f[e];
        dot += ww[e] * g[e] * f[e];
      }
    }
  }
  block_sum2(ss, dot, red);
  const float rstd = rsqrtf(ss / d + eps);
  const float c = dot / d * rstd * rstd * rstd;
  uint4* dxr = reinterpret_cast<uint4*>(dx + fr * d);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int v = threadIdx.x + i * NORM_THREADS;
    if (v < nv) {
      float f[8], g[8], ww[8], o[8];
      unpack8(xv[i], f);
      unpack8(gv[i], g);
      unpack8(__ldg(wr + v), ww);
      unpack8(rv[i], o);
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] += rstd * ww[e] * g[e] - f[e] * c;
      dxr[v] = pack8(o);
    }
  }
}

__global__ void __launch_bounds__(NORM_THREADS)
    layernorm_fwd_kernel(const bf16* __restrict__ x, const bf16* __restrict__ w, const bf16* __restrict__ b,
                         bf16* __restrict__ y, int d, float eps) {
  __shared__ float red[32];
  const int64_t r = blockIdx.x;
  const uint4* xr = reinterpret_cast<const uint4*>(x + r * d);
  const int nv = d >> 3;
  uint4 xv[NORM_MAXV];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NORM_MAXV; ++i) {
    const int v = threadIdx.x + i * NORM_THREADS;
    if (v < nv) {
      xv[i] = xr[v];
      float f[8];
      unpack8(xv[i], f);
#pragma unroll
      for (int e = 0; e < 8; ++e) s += f[e];
    }
  }
  const float mean = block_sum(s, red) / d;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NORM_MAXV; ++i) {
    const int v = threadIdx.x + i * NORM_THREADS;
    if (v < nv) {
      float f[8];
      unpack8(xv[i], f);
#pragma unroll
      for (int e = 0; e < 8; ++e) q += (f[e] - mean) * (f[e] - mean);
    }
  }
  const float rstd = rsqrtf(block_sum(q, red) / d + eps);
  uint4* yr = reinterpret_cast<uint4*>(y + r * d);
  const uint4* wr = reinterpret_cast<const uint4*>(w);
  const uint4* br = reinterpret_cast<const uint4*>(b);
#pragma unroll
  for (int i = 0; i < NORM_MAXV; ++i) {
    const int v = threadIdx.x + i * NORM_THREADS;
    if (v < nv) {
      float f[8], g[8], h[8];
      unpack8(xv[i], f);
      unpack8(__ldg(wr + v), g);
      unpack8(__ldg(br + v), h);
#pragma unroll
      for (int e = 0; e < 8; ++e) f[e] = (f[e] - mean) * rstd * g[e] + h[e];
      yr[v] = pack8(f);
    }
  }
}

// ---------------------------------------------------------------------------------------------- RoPE
__global__ void rope_kernel(bf16* x, int64_t ld, int64_t T, int S, const int32_t* __restrict__ pos, int pos_offset,
                            const int32_t* __restrict__ pos_scalar, int nheads, int hd, const float2* __restrict__ cs,
                            int inverse) {
  const int vec_per_head = hd >> 4;  // uint4 vectors in half a head
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t total = T * nheads * vec_per_head;
  if (idx >= total) return;
  const int v = static_cast<int>(idx % vec_per_head);
  const int h = static_cast<int>((idx / vec_per_head) % nheads);
  const int64_t t = idx / (static_cast<int64_t>(vec_per_head) * nheads);
  const int p = (pos ? pos[t] : static_cast<int>(t % S)) + pos_offset + (pos_scalar ? *pos_scalar : 0);
  bf16* base = x + t * ld + h * hd + v * 8;
  uint4 lo = *reinterpret_cast<uint4*>(base);
  uint4 hi = *reinterpret_cast<uint4*>(base + (hd >> 1));
  float a[8], b[8];
  unpack8(lo, a);
  unpack8(hi, b);
  const float2* c = cs + static_cast<int64_t>(p) * (hd >> 1) + v * 8;
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const float2 k = __ldg(c + e);
    const float sn = inverse ? -k.y : k.y;
    const float na = a[e] * k.x - b[e] * sn;
    const float nb = b[e] * k.x + a[e] * sn;
    a[e] = na;
    b[e] = nb;
  }
  *reinterpret_cast<uint4*>(base) = pack8(a);
  *reinterpret_cast<uint4*>(base + (hd >> 1)) = pack8(b);
}

__global__ void rope_table_kernel(float2* cs, int max_pos, int half, float theta) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= max_pos * half) return;
  const int p = idx / half, j = idx % half;
  // inv_freq = theta^(-2j/hd) in fp32, angle = p * inv_freq in fp32 (HF LlamaRotaryEmbedding arithmetic)
  const float inv = 1.0f / powf(theta, static_cast<float>(2 * j) / static_cast<float>(2 * half));
  const float ang = static_cast<float>(p) * inv;
  cs[idx] = make_float2(cosf(ang), sinf(ang));
}

// ---------------------------------------------------------------------------------------------- SwiGLU / GELU
// (token row, 16 B vector) of a flat thread index. The 64-bit division by a runtime F/8 costs more instructions than the
// rest of these kernels together; every real shape fits 32 bits.
__device__ __forceinline__ bool row_and_vec(int64_t T, int fv, int64_t& t, int& v) {
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t total = T * fv;
  if (idx >= total) return false;
  if (total <= 0xffffffffLL) {
    const uint32_t i = static_cast<uint32_t>(idx), q = i / static_cast<uint32_t>(fv);
    t = q;
    v = static_cast<int>(i - q * static_cast<uint32_t>(fv));
  } else {
    t = idx / fv;
    v = static_cast<int>(idx % fv);
  }
  return true;
}

__global__ void swiglu_fwd_kernel(const bf16* __restrict__ gu, bf16* __restrict__ act, int64_t T, int F) {
  const int fv = F >> 3;
  int64_t t;
  int v;
  if (!row_and_vec(T, fv, t, v)) return;
  const uint4* row = reinterpret_cast<const uint4*>(gu + t * 2 * F);
  float g[8], u[8];
  unpack8(row[v], g);
  unpack8(row[fv + v], u);
#pragma unroll
  for (int e = 0; e < 8; ++e) g[e] = bf16_round(g[e] / (1.f + __expf(-g[e]))) * u[e];
  reinterpret_cast<uint4*>(act + t * F)[v] = pack8(g);
}

// W [2F, d] = [gate rows | up rows]  ->  out [2F, d] with rows in blocks of 128: gate block t, up block t, gate block t+1, ...
__global__ void interleave_gate_up_kernel(const bf16* __restrict__ W, bf16* __restrict__ out, int F, int d) {
  const int dv = d >> 3;
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<int64_t>(2) * F * dv) return;
  const int64_t r = idx / dv;              // destination row
  const int v = static_cast<int>(idx % dv);
  const int blk = static_cast<int>(r / 256), within = static_cast<int>(r % 256);
  const int64_t src = within < 128 ? static_cast<int64_t>(blk) * 128 + within : static_cast<int64_t>(F) + blk * 128 + (within - 128);
  reinterpret_cast<uint4*>(out + r * d)[v] = reinterpret_cast<const uint4*>(W + src * d)[v];
}

__global__ void swiglu_bwd_kernel(const bf16* __restrict__ dact, const bf16* __restrict__ gu, bf16* __restrict__ dgu,
                                  int64_t T, int F) {
  const int fv = F >> 3;
  int64_t t;
  int v;
  if (!row_and_vec(T, fv, t, v)) return;
  const uint4* row = reinterpret_cast<const uint4*>(gu + t * 2 * F);
  float g[8], u[8], da[8], dg[8], du[8];
  unpack8(row[v], g);
  unpack8(row[fv + v], u);
  unpack8(reinterpret_cast<const uint4*>(dact + t * F)[v], da);
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const float sg = 1.f / (1.f + __expf(-g[e]));
    const float silu = g[e] * sg;
    du[e] = da[e] * silu;
    dg[e] = da[e] * u[e] * sg * (1.f + g[e] * (1.f - sg));
  }
  uint4* orow = reinterpret_cast<uint4*>(dgu + t * 2 * F);
  orow[v] = pack8(dg);
  orow[fv + v] = pack8(du);
}

__global__ void gelu_fwd_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, int64_t nv) {
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= nv) return;
  float f[8];
  unpack8(reinterpret_cast<const uint4*>(x)[idx], f);
#pragma unroll
  for (int e = 0; e < 8; ++e) f[e] = 0.5f * f[e] * (1.f + erff(f[e] * 0.70710678118654752f));
  reinterpret_cast<uint4*>(y)[idx] = pack8(f);
}
__global__ void gelu_bwd_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ x, bf16* __restrict__ dx, int64_t nv) {
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= nv) return;
  float f[8], g[8];
  unpack8(reinterpret_cast<const uint4*>(x)[idx], f);
  unpack8(reinterpret_cast<const uint4*>(dy)[idx], g);
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const float cdf = 0.5f * (1.f + erff(f[e] * 0.70710678118654752f));
    const float pdf = 0.3989422804014327f * __expf(-0.5f * f[e] * f[e]);
    g[e] *= cdf + f[e] * pdf;
  }
  reinterpret_cast<uint4*>(dx)[idx] = pack8(g);
}

// ---------------------------------------------------------------------------------------------- embedding splice
__global__ void find_image_pos_kernel(const int32_t* __restrict__ ids, int32_t* __restrict__ img_pos, int N, int L,
                                      int image_token) {
  const int n = blockIdx.x;
  if (n >= N) return;
  int found = L;  // no image token: slot placed past the end (all rows are text)
  for (int i = threadIdx.x; i < L; i += blockDim.x)
    if (ids[static_cast<int64_t>(n) * L + i] == image_token) found = min(found, i);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) found = min(found, __shfl_xor_sync(0xffffffffu, found, o));
  if (threadIdx.x == 0) img_pos[n] = found;
}

__global__ void embed_splice_kernel(const int32_t* __restrict__ ids, const uint8_t* __restrict__ tok_mask,
                                    const uint8_t* __restrict__ img_mask, const int32_t* __restrict__ img_of_row,
                                    const bf16* __restrict__ table, const bf16* __restrict__ feats, bf16* __restrict__ out,
                                    uint8_t* __restrict__ out_mask, const int32_t* __restrict__ img_pos, int L, int P,
                                    int d, int vocab) {
  const int S = L - 1 + P;
  const int64_t row = blockIdx.x;  // n * S + s
  const int n = static_cast<int>(row / S);
  const int s = static_cast<int>(row % S);
  const int p0 = img_pos[n];
  const uint4* src;
  uint8_t m;
  if (s >= p0 && s < p0 + P) {
    src = reinterpret_cast<const uint4*>(feats + (static_cast<int64_t>(img_of_row[n]) * P + (s - p0)) * d);
    m = img_mask ? img_mask[static_cast<int64_t>(img_of_row[n]) * P + (s - p0)] : 1;
  } else {
    const int l = s < p0 ? s : s - P + 1;
    int id = ids[static_cast<int64_t>(n) * L + l];
    id = min(max(id, 0), vocab - 1);
    src = reinterpret_cast<const uint4*>(table + static_cast<int64_t>(id) * d);
    m = tok_mask[static_cast<int64_t>(n) * L + l];
  }
  uint4* dst = reinterpret_cast<uint4*>(out + row * d);
  for (int v = threadIdx.x; v < (d >> 3); v += blockDim.x) dst[v] = __ldg(src + v);
  if (threadIdx.x == 0 && out_mask) out_mask[row] = m;
}

__global__ void embed_splice_bwd_kernel(const bf16* __restrict__ dX, const int32_t* __restrict__ img_of_row,
                                        const int32_t* __restrict__ img_pos, bf16* __restrict__ dfeats, int N, int S,
                                        int P, int d) {
  const int64_t fr = blockIdx.x;  // img * P + p
  const int img = static_cast<int>(fr / P), p = static_cast<int>(fr % P);
  for (int v = threadIdx.x; v < (d >> 3); v += blockDim.x) {
    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = 0.f;
    for (int n = 0; n < N; ++n) {
      if (img_of_row[n] != img) continue;
      float f[8];
      unpack8(reinterpret_cast<const uint4*>(dX + (static_cast<int64_t>(n) * S + img_pos[n] + p) * d)[v], f);
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] += f[e];
    }
    reinterpret_cast<uint4*>(dfeats + fr * d)[v] = pack8(acc);
  }
}

__global__ void embed_rows_kernel(const int32_t* __restrict__ ids, const bf16* __restrict__ table, bf16* __restrict__ out,
                                  int d, int vocab) {
  const int r = blockIdx.x;
  const int id = min(max(ids[r], 0), vocab - 1);
  const uint4* src = reinterpret_cast<const uint4*>(table + static_cast<int64_t>(id) * d);
  uint4* dst = reinterpret_cast<uint4*>(out + static_cast<int64_t>(r) * d);
  for (int v = threadIdx.x; v < (d >> 3); v += blockDim.x) dst[v] = __ldg(src + v);
}

// ---------------------------------------------------------------------------------------------- CLIP front end
__global__ void clip_im2col_kernel(const float* __restrict__ img, bf16* __restrict__ patches, int n, int R, int ps, int Kp) {
  const int g = R / ps;
  const int P = g * g;
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t total = static_cast<int64_t>(n) * P * Kp;
  if (idx >= total) return;
  const int k = static_cast<int>(idx % Kp);
  const int64_t pr = idx / Kp;
  const int p = static_cast<int>(pr % P);
  const int im = static_cast<int>(pr / P);
  float v = 0.f;
  if (k < 3 * ps * ps) {
    const int c = k / (ps * ps), rem = k % (ps * ps);
    const int ky = rem / ps, kx = rem % ps;
    const int py = p / g, px = p % g;
    v = img[((static_cast<int64_t>(im) * 3 + c) * R + (py * ps + ky)) * R + (px * ps + kx)];
  }
  patches[idx] = __float2bfloat16(v);
}

__global__ void clip_assemble_kernel(const bf16* __restrict__ pe, const bf16* __restrict__ cls, const bf16* __restrict__ pos,
                                     bf16* __restrict__ out, int P, int d) {
  const int64_t row = blockIdx.x;  // n * (P+1) + t
  const int t = static_cast<int>(row % (P + 1));
  const int n = static_cast<int>(row / (P + 1));
  const uint4* a = t == 0 ? reinterpret_cast<const uint4*>(cls)
                          : reinterpret_cast<const uint4*>(pe + (static_cast<int64_t>(n) * P + (t - 1)) * d);
  const uint4* b = reinterpret_cast<const uint4*>(pos + static_cast<int64_t>(t) * d);
  uint4* o = reinterpret_cast<uint4*>(out + row * d);
  for (int v = threadIdx.x; v < (d >> 3); v += blockDim.x) {
    float f[8], g[8];
    unpack8(a[v], f);
    unpack8(b[v], g);
#pragma unroll
    for (int e = 0; e < 8; ++e) f[e] += g[e];
    o[v] = pack8(f);
  }
}

__global__ void clip_drop_cls_kernel(const bf16* __restrict__ x, bf16* __restrict__ out, int P, int d) {
  const int64_t row = blockIdx.x;  // n * P + p
  const int p = static_cast<int>(row % P);
  const int n = static_cast<int>(row / P);
  const uint4* a = reinterpret_cast<const uint4*>(x + (static_cast<int64_t>(n) * (P + 1) + 1 + p) * d);
  uint4* o = reinterpret_cast<uint4*>(out + row * d);
  for (int v = threadIdx.x; v < (d >> 3); v += blockDim.x) o[v] = a[v];
}

// ---------------------------------------------------------------------------------------------- log-prob finalise
__global__ void logprob_finalize_kernel(const float4* __restrict__ partials, int ntiles, const float* __restrict__ label_logit,
                                        const int32_t* __restrict__ labels, float* __restrict__ logprob,
                                        float* __restrict__ entropy, float* __restrict__ lse_out, float* __restrict__ ent_raw,
                                        int64_t M) {
  const int64_t r = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (r >= M) return;
  constexpr float LN2 = 0.6931471805599453f;
  float m = -INFINITY;
  for (int i = 0; i < ntiles; ++i) m = fmaxf(m, partials[r * ntiles + i].x);
  float s = 0.f, zs = 0.f;
  for (int i = 0; i < ntiles; ++i) {
    const float4 p = partials[r * ntiles + i];
    const float sc = exp2f(p.x - m);
    s += p.y * sc;
    zs += p.z * sc;
  }
  const float lse2 = m + log2f(s);
  const float lse = lse2 * LN2;
  const int label = labels[r];
  const bool ok = label >= 0;
  logprob[r] = ok ? (label_logit[r] - lse) : 0.f;
  const float h = lse - (zs / s) * LN2;
  if (entropy) entropy[r] = ok ? h : 0.f;
  if (ent_raw) ent_raw[r] = h;  // unmasked, kept for the entropy head's backward
  if (lse_out) lse_out[r] = lse;
}

// ---------------------------------------------------------------------------------------------- DPO loss
__device__ __forceinline__ float softplus_f(float x) { return x > 20.f ? x : log1pf(__expf(x)); }
__device__ __forceinline__ float logsigmoid_f(float x) { return fminf(x, 0.f) - log1pf(__expf(-fabsf(x))); }
__device__ __forceinline__ float sigmoid_f(float x) { return 1.f / (1.f + __expf(-x)); }

__global__ void __launch_bounds__(1024) dpo_pair_kernel(DpoPairArgs a) {
  __shared__ float red[32];
  constexpr float EXP_CAP = 88.7228f;  // floor(log(FLT_MAX)*1e4)/1e4  (dpo_trainer.py:74-81, fp32)
  float sl = 0.f, sa = 0.f, scr = 0.f, smc = 0.f, srr = 0.f, smr = 0.f;
  const float inv_n = 1.f / static_cast<float>(a.n);
  for (int64_t i = threadIdx.x; i < a.n; i += blockDim.x) {
    const float rc = a.ref_c[i], rr = a.ref_r[i];
    const float sc = a.s_c ? a.s_c[i] : 1.f, sr = a.s_r ? a.s_r[i] : 1.f;
    const float dc = a.pi_c[i] - (a.reference_free ? 0.f : rc);
    const float dr = a.pi_r[i] - (a.reference_free ? 0.f : rr);
    float u, du_dc, du_dr;
    if (a.divergence == 2) {
      const float al = a.alpha_div;
      const float xr = -al * dr, xc = -al * dc;
      const float er = __expf(fminf(xr, EXP_CAP)), ec = __expf(fminf(xc, EXP_CAP));
      u = (er - ec) / al;
      du_dc = xc <= EXP_CAP ? ec : 0.f;
      du_dr = xr <= EXP_CAP ? -er : 0.f;
    } else {
      u = sc * dc - sr * dr;
      du_dc = sc;
      du_dr = -sr;
      if (a.divergence == 1) {
        u -= softplus_f(dc) - softplus_f(dr);
        du_dc -= sigmoid_f(dc);
        du_dr += sigmoid_f(dr);
      }
    }
    const float bu = a.beta * u;
    const float loss = -logsigmoid_f(bu) * (1.f - a.label_smoothing) - logsigmoid_f(-bu) * a.label_smoothing;
    const float dl_du = a.beta * (-(1.f - a.label_smoothing) * sigmoid_f(-bu) + a.label_smoothing * sigmoid_f(bu));
    sl += loss;
    float gc = a.loss_coef * inv_n * dl_du * du_dc;
    float gr = a.loss_coef * inv_n * dl_du * du_dr;
    const float cr = a.beta * dc, rrw = a.beta * dr;
    if (a.anchor_mode == 1) {
      const float z = a.anchor_sign * (cr - a.anchor_value);
      sa += -logsigmoid_f(z);
      gc += a.anchor_coef * inv_n * (-sigmoid_f(-z)) * a.anchor_sign * a.beta;
    } else if (a.anchor_mode == 2) {
      const float z = cr - a.anchor_value;
      sa += z * z;
      gc += a.anchor_coef * inv_n * 2.f * z * a.beta;
    }
    if (a.d_c) a.d_c[i] += gc * a.grad_scale;
    if (a.d_r) a.d_r[i] += gr * a.grad_scale;
    const float mc = rc != 0.f ? 1.f : 0.f, mr = rr != 0.f ? 1.f : 0.f;
    scr += cr * mc;
    smc += mc;
    srr += rrw * mr;
    smr += mr;
  }
  sl = block_sum(sl, red);
  sa = block_sum(sa, red);
  scr = block_sum(scr, red);
  smc = block_sum(smc, red);
  srr = block_sum(srr, red);
  smr = block_sum(smr, red);
  if (threadIdx.x == 0) {
    a.out[0] = sl * inv_n;
    a.out[1] = sa * inv_n;
    a.out[2] = scr / smc;
    a.out[3] = srr / smr;
    a.out[4] = smc;
    a.out[5] = smr;
    a.out[6] = 0.f;
    a.out[7] = 0.f;
  }
}

// ---------------------------------------------------------------------------------------------- optimiser
__global__ void sumsq_kernel(const float* __restrict__ g, int64_t n, float* out) {
  __shared__ float red[32];
  float s = 0.f;
  const int64_t n4 = n >> 2;
  const float4* g4 = reinterpret_cast<const float4*>(g);
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float4 v = g4[i];
    s += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  if (blockIdx.x == 0)
    for (int64_t i = (n4 << 2) + threadIdx.x; i < n; i += blockDim.x) s += g[i] * g[i];
  s = block_sum(s, red);
  if (threadIdx.x == 0) atomicAdd(out, s);
}

// One element of the update; the vector and scalar paths share it so both produce the same bits.
__device__ __forceinline__ float adamw_one(float g, float p, float& mi, float& vi, float lr, float b1, float b2, float eps,
                                           float wd, float bc1, float bc2, float gs) {
  g *= gs;
  p *= 1.f - lr * wd;
  mi = b1 * mi + (1.f - b1) * g;
  vi = b2 * vi + (1.f - b2) * g * g;
  const float denom = sqrtf(vi) / sqrtf(bc2) + eps;
  return p - (lr / bc1) * (mi / denom);
}

// 16 B per thread and array (4 streams read, 4 written: 30 B per element, touched once -> streaming loads / stores).
__global__ void __launch_bounds__(256) adamw_kernel(float* __restrict__ master, const float* __restrict__ grad,
                                                    float* __restrict__ m, float* __restrict__ v, bf16* __restrict__ shadow,
                                                    int64_t n, float lr, float b1, float b2, float eps, float wd, float bc1,
                                                    float bc2, float gscale, const float* sumsq, float max_norm) {
  float gs = gscale;
  if (sumsq && max_norm > 0.f) {
    const float norm = sqrtf(*sumsq) * gscale;
    gs *= fminf(1.f, max_norm / (norm + 1e-6f));
  }
  const int64_t n4 = n >> 2;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 g = __ldcs(reinterpret_cast<const float4*>(grad) + i);
    float4 p = __ldcs(reinterpret_cast<const float4*>(master) + i);
    float4 mi = __ldcs(reinterpret_cast<const float4*>(m) + i);
    float4 vi = __ldcs(reinterpret_cast<const float4*>(v) + i);
    p.x = adamw_one(g.x, p.x, mi.x, vi.x, lr, b1, b2, eps, wd, bc1, bc2, gs);
    p.y = adamw_one(g.y, p.y, mi.y, vi.y, lr, b1, b2, eps, wd, bc1, bc2, gs);
    p.z = adamw_one(g.z, p.z, mi.z, vi.z, lr, b1, b2, eps, wd, bc1, bc2, gs);
    p.w = adamw_one(g.w, p.w, mi.w, vi.w, lr, b1, b2, eps, wd, bc1, bc2, gs);
    __stcs(reinterpret_cast<float4*>(m) + i, mi);
    __stcs(reinterpret_cast<float4*>(v) + i, vi);
    __stcs(reinterpret_cast<float4*>(master) + i, p);
    __nv_bfloat162 lo = __floats2bfloat162_rn(p.x, p.y), hi = __floats2bfloat162_rn(p.z, p.w);
    uint2 pk;
    pk.x = *reinterpret_cast<uint32_t*>(&lo);
    pk.y = *reinterpret_cast<uint32_t*>(&hi);
    *reinterpret_cast<uint2*>(shadow + 4 * i) = pk;  // the next forward reads the shadow: keep it cacheable
  }
  for (int64_t i = (n4 << 2) + static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    float mi = m[i], vi = v[i];
    const float p = adamw_one(grad[i], master[i], mi, vi, lr, b1, b2, eps, wd, bc1, bc2, gs);
    m[i] = mi;
    v[i] = vi;
    master[i] = p;
    shadow[i] = __float2bfloat16(p);
  }
}

__global__ void f32_to_bf16_kernel(const float* __restrict__ src, bf16* __restrict__ dst, int64_t n) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    dst[i] = __float2bfloat16(src[i]);
}
__global__ void bf16_to_f32_kernel(const bf16* __restrict__ src, float* __restrict__ dst, int64_t n) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    dst[i] = __bfloat162float(src[i]);
}

__global__ void norm_from_sumsq_kernel(const float* sumsq, float gscale, float* out) { *out = sqrtf(*sumsq) * gscale; }
__global__ void fill_affine_kernel(int32_t* out, int n, int stride, int offset) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = i * stride + offset;
}
__global__ void fill_mod_kernel(int32_t* out, int n, int mod) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = i % mod;
}
__global__ void queries_to_i32_kernel(const int64_t* __restrict__ q, const uint8_t* __restrict__ qm, int32_t* ids,
                                      uint8_t* mask, int64_t* out, int64_t out_ld, int N, int Q) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * Q) return;
  const int n = i / Q, l = i % Q;
  ids[i] = static_cast<int32_t>(q[i]);
  mask[i] = qm[i];
  out[static_cast<int64_t>(n) * out_ld + l] = q[i];
}

// ---------------------------------------------------------------------------------------------- input stacking
struct RespPtrs {
  const int64_t* p[8];
};
// Row layouts (N rows of L ids -> S = L-1+P positions after the image splice):
//   unpacked: N = K*b rows [query | resp_k]          (rl_models.py:95-111 as written)
//   packed  : N = b rows   [query | resp_0 | resp_1 | ...]: the image+query prefix is computed ONCE per sample and
//             shared by its K responses through the segment mask of the attention kernels (identical results).
// labels / outputs are always in the reference's stacked order c = (k*b + i)*R + t.
__global__ void stack_inputs_kernel(const int64_t* __restrict__ queries, const uint8_t* __restrict__ qmask, RespPtrs resp,
                                    int K, int b, int Q, int R, int pad_id, int packed, int32_t* __restrict__ ids,
                                    uint8_t* __restrict__ mask, int32_t* __restrict__ labels) {
  const int L = packed ? Q + K * R : Q + R;
  const int N = packed ? b : K * b;
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<int64_t>(N) * L) return;
  const int l = static_cast<int>(idx % L);
  const int row = static_cast<int>(idx / L);
  const int i = packed ? row : row % b;
  if (l < Q) {
    ids[idx] = static_cast<int32_t>(queries[static_cast<int64_t>(i) * Q + l]);
    mask[idx] = qmask[static_cast<int64_t>(i) * Q + l];
  } else {
    const int k = packed ? (l - Q) / R : row / b;
    const int t = packed ? (l - Q) % R : l - Q;
    const int32_t id = static_cast<int32_t>(resp.p[k][static_cast<int64_t>(i) * R + t]);
    ids[idx] = id;
    mask[idx] = id != pad_id;
    labels[(static_cast<int64_t>(k) * b + i) * R + t] = id != pad_id ? id : kIgnore;
  }
}

// RoPE positions: absolute index in the padded row; response segments of a packed row all start at S0.
__global__ void build_positions_kernel(int32_t* __restrict__ pos, int N, int S, int S0, int R, int packed) {
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<int64_t>(N) * S) return;
  const int s = static_cast<int>(idx % S);
  pos[idx] = (packed && s >= S0) ? S0 + (s - S0) % R : s;
}

// Hidden row that predicts response token (k, i, t) (logits[:, -R-1:-1], rl_models.py:120-121) and the row its
// gradient is written to. Packed rows share the last prefix row for t == 0: k >= 1 writes go to scratch rows
// [T, T + (K-1)*b) that fold_rows adds back afterwards.
__global__ void build_row_index_kernel(int32_t* __restrict__ src, int32_t* __restrict__ dst, int K, int b, int R, int S,
                                       int S0, int packed, int T) {
  const int64_t c = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (c >= static_cast<int64_t>(K) * b * R) return;
  const int t = static_cast<int>(c % R);
  const int kb = static_cast<int>(c / R);
  const int k = kb / b, i = kb % b;
  if (!packed) {
    const int row = kb * S + (S - R - 1 + t);
    src[c] = row;
    dst[c] = row;
  } else {
    const int row = i * S + (t == 0 ? S0 - 1 : S0 + k * R + t - 1);
    src[c] = row;
    dst[c] = (t == 0 && k >= 1) ? T + (k - 1) * b + i : row;
  }
}

__global__ void fold_rows_kernel(bf16* __restrict__ dx, int K, int b, int S, int S0, int d, int T) {
  const int i = blockIdx.x;
  uint4* dst = reinterpret_cast<uint4*>(dx + (static_cast<int64_t>(i) * S + S0 - 1) * d);
  for (int v = threadIdx.x; v < (d >> 3); v += blockDim.x) {
    float acc[8];
    unpack8(dst[v], acc);
    for (int k = 1; k < K; ++k) {
      float f[8];
      unpack8(reinterpret_cast<const uint4*>(dx + (static_cast<int64_t>(T) + (k - 1) * b + i) * d)[v], f);
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] += f[e];
    }
    dst[v] = pack8(acc);
  }
}

// ---------------------------------------------------------------------------------------------- decode
__global__ void kv_store_prefill_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ kc, bf16* __restrict__ vc, int S,
                                        int S_max, int d) {
  const int64_t row = blockIdx.x;  // n * S + s
  const int n = static_cast<int>(row / S), s = static_cast<int>(row % S);
  const uint4* k = reinterpret_cast<const uint4*>(qkv + row * 3 * d + d);
  const uint4* v = reinterpret_cast<const uint4*>(qkv + row * 3 * d + 2 * d);
  uint4* ko = reinterpret_cast<uint4*>(kc + (static_cast<int64_t>(n) * S_max + s) * d);
  uint4* vo = reinterpret_cast<uint4*>(vc + (static_cast<int64_t>(n) * S_max + s) * d);
  for (int i = threadIdx.x; i < (d >> 3); i += blockDim.x) {
    ko[i] = k[i];
    vo[i] = v[i];
  }
}

// ---- split-K partial sums [splits, M, N] fp32 (gemm EPI_F32_PARTIAL) -> bf16, summed in a fixed order
__device__ __forceinline__ float sum_partials(const float* __restrict__ part, int splits, int64_t slice, int64_t idx) {
  float a = 0.f;
  for (int sp = 0; sp < splits; ++sp) a += part[sp * slice + idx];
  return a;
}
__global__ void partial_to_bf16_kernel(const float* __restrict__ part, int splits, int64_t n, bf16* __restrict__ out) {
  const int64_t i = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) * 4;
  if (i >= n) return;  // n % 4 == 0 (checked by the launcher)
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int sp = 0; sp < splits; ++sp) {
    const float4 p = *reinterpret_cast<const float4*>(part + sp * n + i);
    a.x += p.x;
    a.y += p.y;
    a.z += p.z;
    a.w += p.w;
  }
  uint2 o;
  o.x = pack2(a.x, a.y);
  o.y = pack2(a.z, a.w);
  *reinterpret_cast<uint2*>(out + i) = o;
}
// Decode step, q|k|v projection: sum the split-K slices of [M, 3d] in slice order, rotate q and k (HF rotate_half pairing,
// position pos_base + *cur for every row) on the fp32 sums and store bf16 — one launch instead of partial_to_bf16 + rope.
// One thread per 4 consecutive columns of a head's lower half (plus their partners 64 columns on) for q and k; per 4
// consecutive columns for v.
__global__ void partial_qkv_rope_kernel(const float* __restrict__ part, int splits, int64_t slice, bf16* __restrict__ qkv,
                                        int64_t M, int d, int hd, const float2* __restrict__ cs, int pos_base,
                                        const int32_t* __restrict__ cur) {
  const int half = hd >> 1;
  const int rot_items = (2 * d / hd) * (half >> 2);  // per row: q and k heads x 4-column groups of the lower half
  const int v_items = d >> 2;
  const int per_row = rot_items + v_items;
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= M * per_row) return;
  const int64_t m = idx / per_row;
  const int it = static_cast<int>(idx - m * per_row);
  const int64_t row = m * 3 * d;
  auto sum4 = [&](int64_t off) {
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int sp = 0; sp < splits; ++sp) {
      const float4 p = *reinterpret_cast<const float4*>(part + sp * slice + off);
      a.x += p.x;
      a.y += p.y;
      a.z += p.z;
      a.w += p.w;
    }
    return a;
  };
  auto store4 = [&](int64_t off, float4 a) {
    uint2 o;
    o.x = pack2(a.x, a.y);
    o.y = pack2(a.z, a.w);
    *reinterpret_cast<uint2*>(qkv + off) = o;
  };
  if (it >= rot_items) {
    const int64_t off = row + 2 * d + (it - rot_items) * 4;
    store4(off, sum4(off));
    return;
  }
  const int g4 = half >> 2;
  const int head = it / g4, j = (it - head * g4) * 4;
  const int64_t off = row + head * hd + j;
  const float4 lo = sum4(off), hi = sum4(off + half);
  const int pos = pos_base + (cur ? *cur : 0);
  const float2* c = cs + static_cast<int64_t>(pos) * half + j;
  const float2 c0 = __ldg(c), c1 = __ldg(c + 1), c2 = __ldg(c + 2), c3 = __ldg(c + 3);
  store4(off, make_float4(lo.x * c0.x - hi.x * c0.y, lo.y * c1.x - hi.y * c1.y, lo.z * c2.x - hi.z * c2.y, lo.w * c3.x - hi.w * c3.y));
  store4(off + half, make_float4(hi.x * c0.x + lo.x * c0.y, hi.y * c1.x + lo.y * c1.y, hi.z * c2.x + lo.z * c2.y, hi.w * c3.x + lo.w * c3.y));
}

// x_out = bf16(sum(part) + res);  h = w * bf16(x_out * rstd)   (residual add fused with the next RMSNorm).
// One 1024-thread CTA per row, 4 (or 8) consecutive columns per thread, all split slices loaded as independent float4s.
constexpr int PRN_THREADS = 1024;
template <int VPT>  // float4 vectors per thread
__global__ void __launch_bounds__(PRN_THREADS)
    partial_residual_rmsnorm_kernel(const float* __restrict__ part, int splits, int64_t slice, const bf16* __restrict__ res,
                                    bf16* __restrict__ x_out, const bf16* __restrict__ w, bf16* __restrict__ h, int d,
                                    float eps) {
  __shared__ float red[32];
  const int64_t r = blockIdx.x;
  float v[VPT][4];
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < VPT; ++i) {
    const int c = (threadIdx.x + i * PRN_THREADS) * 4;
    v[i][0] = v[i][1] = v[i][2] = v[i][3] = 0.f;
    if (c < d) {
      for (int sp = 0; sp < splits; ++sp) {
        const float4 p = *reinterpret_cast<const float4*>(part + sp * slice + r * d + c);
        v[i][0] += p.x;
        v[i][1] += p.y;
        v[i][2] += p.z;
        v[i][3] += p.w;
      }
      const uint2 rr = *reinterpret_cast<const uint2*>(res + r * d + c);
      const float2 r01 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&rr.x));
      const float2 r23 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&rr.y));
      v[i][0] = bf16_round(v[i][0] + r01.x);
      v[i][1] = bf16_round(v[i][1] + r01.y);
      v[i][2] = bf16_round(v[i][2] + r23.x);
      v[i][3] = bf16_round(v[i][3] + r23.y);
      uint2 o;
      o.x = pack2(v[i][0], v[i][1]);
      o.y = pack2(v[i][2], v[i][3]);
      *reinterpret_cast<uint2*>(x_out + r * d + c) = o;
      ss += v[i][0] * v[i][0] + v[i][1] * v[i][1] + v[i][2] * v[i][2] + v[i][3] * v[i][3];
    }
  }
  ss = block_sum(ss, red);
  const float rstd = rsqrtf(ss / d + eps);
  if (h) {
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
      const int c = (threadIdx.x + i * PRN_THREADS) * 4;
      if (c < d) {
        const uint2 ww = *reinterpret_cast<const uint2*>(w + c);
        const float2 w01 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&ww.x));
        const float2 w23 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&ww.y));
        uint2 o;
        o.x = pack2(w01.x * bf16_round(v[i][0] * rstd), w01.y * bf16_round(v[i][1] * rstd));
        o.y = pack2(w23.x * bf16_round(v[i][2] * rstd), w23.y * bf16_round(v[i][3] * rstd));
        *reinterpret_cast<uint2*>(h + r * d + c) = o;
      }
    }
  }
}
// act = bf16(silu(bf16(g)) ) * bf16(u) from partial sums of the fused gate|up GEMM [M, 2F]
__global__ void partial_swiglu_kernel(const float* __restrict__ part, int splits, int64_t slice, bf16* __restrict__ act,
                                      int64_t M, int F) {
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= M * F) return;
  const int64_t m = idx / F;
  const int f = static_cast<int>(idx % F);
  const float g = bf16_round(sum_partials(part, splits, slice, m * 2 * F + f));
  const float u = bf16_round(sum_partials(part, splits, slice, m * 2 * F + F + f));
  act[idx] = __float2bfloat16(bf16_round(g / (1.f + __expf(-g))) * u);
}

// One CTA per (head, sequence), 8 warps; each half-warp owns one key per iteration: 16 lanes x 16 B = one 256 B head row
// (coalesced), online softmax per half-warp, merged through shared memory. New k/v are appended at position
// len = len_base + *cur first.
template <int HD>
__global__ void __launch_bounds__(256) decode_attention_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ kc,
                                                               bf16* __restrict__ vc, const uint8_t* __restrict__ kmask,
                                                               bf16* __restrict__ o, int H, int len_base,
                                                               const int32_t* __restrict__ cur, int S_max, float scale) {
  constexpr int LPK = HD / 8;        // lanes per key (16 for hd 128, 8 for hd 64)
  constexpr int KPW = 32 / LPK;      // keys per warp per iteration
  constexpr int GROUPS = 8 * KPW;    // key groups per CTA
  const int n = blockIdx.y, h = blockIdx.x;
  const int d = H * HD;
  const int len = len_base + (cur ? *cur : 0);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sub = lane / LPK, ln = lane % LPK;
  const int grp = warp * KPW + sub;
  const bf16* qrow = qkv + static_cast<int64_t>(n) * 3 * d + h * HD;
  bf16* kbase = kc + static_cast<int64_t>(n) * S_max * d + h * HD;
  bf16* vbase = vc + static_cast<int64_t>(n) * S_max * d + h * HD;
  if (threadIdx.x < LPK) {  // append the new key/value (16 B per lane)
    *reinterpret_cast<uint4*>(kbase + static_cast<int64_t>(len) * d + threadIdx.x * 8) =
        *reinterpret_cast<const uint4*>(qrow + d + threadIdx.x * 8);
    *reinterpret_cast<uint4*>(vbase + static_cast<int64_t>(len) * d + threadIdx.x * 8) =
        *reinterpret_cast<const uint4*>(qrow + 2 * d + threadIdx.x * 8);
  }
  __syncthreads();
  float q[8], acc[8];
  unpack8(*reinterpret_cast<const uint4*>(qrow + ln * 8), q);
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    q[e] *= scale;
    acc[e] = 0.f;
  }
  float m = -INFINITY, l = 0.f;
  const uint8_t* km = kmask + static_cast<int64_t>(n) * S_max;
  // UN keys per group and iteration: their mask bytes, then all their K / V rows (2 * UN independent 16 B loads per lane in
  // flight), then ONE running-max update for the chunk. The trip count is uniform over the CTA (invalid or out-of-range keys
  // are predicated off), so the half-warp shuffles run converged.
  constexpr int UN = 4;
  const int iters = (len + GROUPS * UN) / (GROUPS * UN);  // ceil((len + 1) / (GROUPS * UN))
  for (int it = 0; it < iters; ++it) {
    const int j0 = it * GROUPS * UN + grp;
    bool ok[UN];
    uint4 kraw[UN], vraw[UN];
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const int j = j0 + u * GROUPS;
      ok[u] = j <= len && km[min(j, len)] != 0;
    }
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const int64_t row = static_cast<int64_t>(min(j0 + u * GROUPS, len)) * d + ln * 8;  // clamped: always a written row
      kraw[u] = *reinterpret_cast<const uint4*>(kbase + row);
      vraw[u] = *reinterpret_cast<const uint4*>(vbase + row);
    }
    float sv[UN];
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      float kf[8];
      unpack8(kraw[u], kf);
      float t = 0.f;
#pragma unroll
      for (int e = 0; e < 8; ++e) t += q[e] * kf[e];
      sv[u] = t;
    }
#pragma unroll
    for (int o2 = LPK / 2; o2 > 0; o2 >>= 1) {
#pragma unroll
      for (int u = 0; u < UN; ++u) sv[u] += __shfl_xor_sync(0xffffffffu, sv[u], o2);
    }
    float nm = m;
#pragma unroll
    for (int u = 0; u < UN; ++u) nm = fmaxf(nm, ok[u] ? sv[u] : -INFINITY);
    const float base = nm == -INFINITY ? 0.f : nm;  // nothing valid yet: every factor below is exp(-inf) = 0
    const float sc = __expf(m - base);
    float pw[UN], psum = 0.f;
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      pw[u] = ok[u] ? __expf(sv[u] - base) : 0.f;
      psum += pw[u];
    }
    l = l * sc + psum;
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] *= sc;
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      float vf[8];
      unpack8(vraw[u], vf);
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] += pw[u] * vf[e];
    }
    m = nm;
  }
  __shared__ float sm[GROUPS], sl[GROUPS];
  __shared__ float sacc[GROUPS][HD];
  if (ln == 0) {
    sm[grp] = m;
    sl[grp] = l;
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) sacc[grp][ln * 8 + e] = acc[e];
  __syncthreads();
  if (threadIdx.x < HD) {
    float gm = -INFINITY;
    for (int g2 = 0; g2 < GROUPS; ++g2) gm = fmaxf(gm, sm[g2]);
    float tot = 0.f, out = 0.f;
    for (int g2 = 0; g2 < GROUPS; ++g2) {
      const float sc = sm[g2] == -INFINITY ? 0.f : __expf(sm[g2] - gm);
      tot += sl[g2] * sc;
      out += sacc[g2][threadIdx.x] * sc;
    }
    o[static_cast<int64_t>(n) * d + h * HD + threadIdx.x] = __float2bfloat16(tot > 0.f ? out / tot : 0.f);
  }
}

// HF SuppressTokensLogitsProcessor [UPSTREAM; online_generator.py:298-302, rl_models.py:173-177 pass
// suppress_tokens=[eos] under suppress_eos_at_generation]: the listed ids get logit -inf before anything else looks at them.
struct SuppressList {
  int n;
  int ids[OPADPO_MAX_SUPPRESS];
};
__device__ __forceinline__ bool suppressed(const SuppressList& sl, int c) {
  bool hit = false;
#pragma unroll
  for (int i = 0; i < OPADPO_MAX_SUPPRESS; ++i) hit |= (i < sl.n && sl.ids[i] == c);
  return hit;
}

// Greedy pick over split-K partial logits [splits, N, V] (summed in fixed order). Writes the id at column *cur.
__global__ void __launch_bounds__(1024) greedy_pick_kernel(const float* __restrict__ part, int splits, int64_t slice, int V,
                                                           int eos_id, int pad_id, SuppressList sup, uint8_t* done,
                                                           int32_t* next_ids, int64_t* out, int64_t out_ld,
                                                           const int32_t* __restrict__ cur) {
  __shared__ float bv[32];
  __shared__ int bi[32];
  const int n = blockIdx.x;
  float best = -INFINITY;
  int besti = V;
  for (int i = threadIdx.x; i < V; i += blockDim.x) {
    float x = sum_partials(part, splits, slice, static_cast<int64_t>(n) * V + i);
    if (suppressed(sup, i)) x = -INFINITY;
    if (x > best || (x == best && i < besti)) {
      best = x;
      besti = i;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
    if (ob > best || (ob == best && oi < besti)) {
      best = ob;
      besti = oi;
    }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) {
    bv[warp] = best;
    bi[warp] = besti;
  }
  __syncthreads();
  if (warp == 0) {
    best = lane < (blockDim.x >> 5) ? bv[lane] : -INFINITY;
    besti = lane < (blockDim.x >> 5) ? bi[lane] : V;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
      if (ob > best || (ob == best && oi < besti)) {
        best = ob;
        besti = oi;
      }
    }
    if (lane == 0) {
      int id = besti;
      if (done[n]) id = pad_id;  // HF: finished sequences are padded (unfinished_sequences mask)
      if (id == eos_id) done[n] = 1;
      next_ids[n] = id;
      out[static_cast<int64_t>(n) * out_ld + *cur] = id;
    }
  }
}
__global__ void advance_counter_kernel(int32_t* cur) { *cur += 1; }

// Counter-based uniform in [0,1): splitmix64 over (seed, row, step) — reproducible per (seed, sequence, position).
__device__ __forceinline__ float uniform01(uint64_t seed, uint32_t row, uint32_t step) {
  uint64_t z = seed + 0x9E3779B97F4A7C15ull * (static_cast<uint64_t>(row) * 0x100000001B3ull + step + 1);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  return static_cast<float>(z >> 40) * (1.0f / 16777216.0f);
}

// Order-preserving map float -> u32 (larger float <=> larger key; -inf is the smallest non-NaN key).
__device__ __forceinline__ uint32_t float_key(float x) {
  const uint32_t b = __float_as_uint(x);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

// HF sampling [UPSTREAM GenerationMixin.sample as driven by online_generator.py:292-309 / rl_models.py:166-188]:
// suppress_tokens, then the warpers in HF's order — temperature; top-k (TopKLogitsWarper: drop every score below the
// k-th largest, ties with it survive); top-p (TopPLogitsWarper: over the softmax of what is left, ascending cumulative
// mass <= 1 - top_p is dropped, the largest token always survives) — then a multinomial over the survivors. Any
// top_k (0 = off) and any top_p in (0, 1]. One CTA per sequence; each thread owns a strided slice of the vocabulary in
// registers. Both thresholds come from an MSB-first radix select over the order-preserving keys (4 passes of 8 bits
// through a 256-bin shared-memory histogram): counts for top-k, 2^-32 fixed-point masses for top-p (integer adds, so the
// threshold does not depend on the order the atomics land in).
constexpr int SAMPLE_THREADS = 1024;
constexpr int SAMPLE_MAX_PER_THREAD = 40;  // vocab <= 40960
__global__ void __launch_bounds__(SAMPLE_THREADS)
    sample_pick_kernel(const float* __restrict__ part, int splits, int64_t slice, int V, int top_k, float top_p,
                       float inv_temp, uint64_t seed, int eos_id, int pad_id, SuppressList sup, uint8_t* done,
                       int32_t* next_ids, int64_t* out, int64_t out_ld, const int32_t* __restrict__ cur) {
  __shared__ float bv[32];
  __shared__ int bi[32];
  __shared__ float red[32];
  __shared__ unsigned long long hist[256];
  __shared__ uint32_t sel_prefix;
  __shared__ unsigned long long sel_rest;
  __shared__ float tsum[SAMPLE_THREADS];
  __shared__ int pick_s;
  const int n = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float v[SAMPLE_MAX_PER_THREAD];
#pragma unroll
  for (int i = 0; i < SAMPLE_MAX_PER_THREAD; ++i) {
    const int c = threadIdx.x + i * SAMPLE_THREADS;
    v[i] = (c < V && !suppressed(sup, c)) ? sum_partials(part, splits, slice, static_cast<int64_t>(n) * V + c) * inv_temp
                                           : -INFINITY;
  }
  // MSB-first radix select. DESC = true: the key of the `target`-th largest element by count (top-k). DESC = false: the
  // smallest key whose ascending cumulative weight exceeds `target` (top-p); weights come from `wt(i)`.
  auto radix_select = [&](bool desc, unsigned long long target, auto wt) -> uint32_t {
    uint32_t prefix = 0, pmask = 0;
    unsigned long long rest = target;
    for (int shift = 24; shift >= 0; shift -= 8) {
      for (int b = threadIdx.x; b < 256; b += SAMPLE_THREADS) hist[b] = 0ull;
      __syncthreads();
#pragma unroll
      for (int i = 0; i < SAMPLE_MAX_PER_THREAD; ++i) {
        const int c = threadIdx.x + i * SAMPLE_THREADS;
        if (c < V) {
          const uint32_t k = float_key(v[i]);
          if ((k & pmask) == prefix) {
            const unsigned long long wgt = wt(i);
            if (wgt) atomicAdd(&hist[(k >> shift) & 255u], wgt);
          }
        }
      }
      __syncthreads();
      if (threadIdx.x == 0) {
        unsigned long long acc = 0;
        int digit = desc ? 0 : 255;
        if (desc) {
          for (int b = 255; b >= 0; --b) {
            if (acc + hist[b] >= rest) {
              digit = b;
              break;
            }
            acc += hist[b];
          }
        } else {
          for (int b = 0; b < 256; ++b) {
            if (acc + hist[b] > rest) {
              digit = b;
              break;
            }
            acc += hist[b];
          }
        }
        sel_prefix = prefix | (static_cast<uint32_t>(digit) << shift);
        sel_rest = rest - acc;
      }
      __syncthreads();
      prefix = sel_prefix;
      rest = sel_rest;
      pmask |= 255u << shift;
    }
    return prefix;
  };
  if (top_k > 0 && top_k < V) {
    const uint32_t kth = radix_select(true, static_cast<unsigned long long>(top_k), [&](int) { return 1ull; });
#pragma unroll
    for (int i = 0; i < SAMPLE_MAX_PER_THREAD; ++i)
      if (float_key(v[i]) < kth) v[i] = -INFINITY;
  }
  // softmax statistics of the survivors
  float mx = -INFINITY;
#pragma unroll
  for (int i = 0; i < SAMPLE_MAX_PER_THREAD; ++i) mx = fmaxf(mx, v[i]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if (lane == 0) bv[warp] = mx;
  __syncthreads();
  mx = bv[0];
  for (int w = 1; w < 32; ++w) mx = fmaxf(mx, bv[w]);
  __syncthreads();
  if (top_p < 1.f) {
    // fixed-point masses e_i = exp(v_i - mx) * 2^32 (the largest token weighs exactly 2^32)
    auto mass = [&](int i) -> unsigned long long {
      return v[i] == -INFINITY ? 0ull : static_cast<unsigned long long>(__expf(v[i] - mx) * 4294967296.f);
    };
    unsigned long long mine = 0;
#pragma unroll
    for (int i = 0; i < SAMPLE_MAX_PER_THREAD; ++i) mine += mass(i);
    if (threadIdx.x == 0) hist[0] = 0ull;
    __syncthreads();
    atomicAdd(&hist[0], mine);
    __syncthreads();
    const unsigned long long tot = hist[0];
    __syncthreads();
    // drop while the ascending cumulative mass stays <= (1 - top_p) * total
    const unsigned long long target = static_cast<unsigned long long>(static_cast<double>(tot) * (1.0 - static_cast<double>(top_p)));
    const uint32_t thr = radix_select(false, target, mass);
#pragma unroll
    for (int i = 0; i < SAMPLE_MAX_PER_THREAD; ++i)
      if (float_key(v[i]) < thr) v[i] = -INFINITY;
  }
  // multinomial over the survivors: inverse CDF, thread t owns columns {t + i*1024}, CDF ordered thread-major (any fixed
  // order is a valid multinomial)
  float mine = 0.f;
#pragma unroll
  for (int i = 0; i < SAMPLE_MAX_PER_THREAD; ++i) mine += __expf(v[i] - mx);  // -inf -> 0
  const float tot = block_sum(mine, red);
  tsum[threadIdx.x] = mine;
  __syncthreads();
  if (threadIdx.x == 0) {
    const float u = uniform01(seed, n, static_cast<uint32_t>(*cur)) * tot;
    float acc = 0.f, base = 0.f;
    int owner = 0;
    for (int t = 0; t < SAMPLE_THREADS; ++t) {
      if (tsum[t] > 0.f) {  // without a hit the draw falls to the last thread that holds any mass
        owner = t;
        base = acc;
        if (u < acc + tsum[t]) break;
        acc += tsum[t];
      }
    }
    bi[0] = owner;
    bv[0] = u - base;
  }
  __syncthreads();
  if (threadIdx.x == bi[0]) {
    const float rem = bv[0];
    float acc = 0.f;
    int pick = -1, last = 0;
#pragma unroll
    for (int i = 0; i < SAMPLE_MAX_PER_THREAD; ++i) {
      const int c = threadIdx.x + i * SAMPLE_THREADS;
      if (c < V && pick < 0 && v[i] != -INFINITY) {
        last = c;
        acc += __expf(v[i] - mx);
        if (rem < acc) pick = c;
      }
    }
    pick_s = pick >= 0 ? pick : last;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int id = pick_s;
    if (done[n]) id = pad_id;
    if (id == eos_id) done[n] = 1;
    next_ids[n] = id;
    out[static_cast<int64_t>(n) * out_ld + *cur] = id;
  }
}

inline int check_launch(const char* what) {
  count_launch();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    fprintf(stderr, "[opadpo] %s launch failed: %s\n", what, cudaGetErrorString(e));
    return -11;
  }
  return 0;
}
inline unsigned blocks_for(int64_t n, int threads) { return static_cast<unsigned>((n + threads - 1) / threads); }
inline unsigned capped_blocks(int64_t n, int threads, int cap) {
  const int64_t b = (n + threads - 1) / threads;
  return static_cast<unsigned>(b < cap ? (b < 1 ? 1 : b) : cap);
}

}  // namespace

int rmsnorm_fwd(const bf16* x, const bf16* w, bf16* y, int64_t rows, int d, float eps, const int32_t* x_idx, cudaStream_t s) {
  if (d % 8 != 0 || d > 8 * NORM_THREADS * NORM_MAXV) return -1;
  if (rows == 0) return 0;
  rmsnorm_fwd_kernel<<<static_cast<unsigned>(rows), NORM_THREADS, 0, s>>>(x, w, y, d, eps, x_idx);
  return check_launch("rmsnorm_fwd");
}
int rmsnorm_bwd(const bf16* dy, const bf16* x, const bf16* w, const bf16* dres, bf16* dx, int64_t rows, int d, float eps,
                const int32_t* x_idx, const int32_t* dx_idx, cudaStream_t s) {
  if (d % 8 != 0 || d > 8 * NORM_THREADS * NORM_MAXV) return -1;
  if (rows == 0) return 0;
  const int nvt = ((d >> 3) + NORM_THREADS - 1) / NORM_THREADS;
  const unsigned grid = static_cast<unsigned>(rows);
  if (nvt <= 1)
    rmsnorm_bwd_kernel<1><<<grid, NORM_THREADS, 0, s>>>(dy, x, w, dres, dx, d, eps, x_idx, dx_idx);
  else if (nvt == 2)
    rmsnorm_bwd_kernel<2><<<grid, NORM_THREADS, 0, s>>>(dy, x, w, dres, dx, d, eps, x_idx, dx_idx);
  else if (nvt == 3)
    rmsnorm_bwd_kernel<3><<<grid, NORM_THREADS, 0, s>>>(dy, x, w, dres, dx, d, eps, x_idx, dx_idx);
  else
    rmsnorm_bwd_kernel<4><<<grid, NORM_THREADS, 0, s>>>(dy, x, w, dres, dx, d, eps, x_idx, dx_idx);
  return check_launch("rmsnorm_bwd");
}
int layernorm_fwd(const bf16* x, const bf16* w, const bf16* b, bf16* y, int64_t rows, int d, float eps, cudaStream_t s) {
  if (d % 8 != 0 || d > 8 * NORM_THREADS * NORM_MAXV) return -1;
  if (rows == 0) return 0;
  layernorm_fwd_kernel<<<static_cast<unsigned>(rows), NORM_THREADS, 0, s>>>(x, w, b, y, d, eps);
  return check_launch("layernorm_fwd");
}
int rope_inplace(bf16* x, int64_t ld, int64_t T, int S, const int32_t* pos, int pos_offset, const int32_t* pos_scalar,
                 int nheads, int hd, const float2* cs, int inverse, cudaStream_t s) {
  if (hd % 16 != 0) return -1;
  const int64_t total = T * nheads * (hd >> 4);
  if (total == 0) return 0;
  rope_kernel<<<blocks_for(total, 256), 256, 0, s>>>(x, ld, T, S, pos, pos_offset, pos_scalar, nheads, hd, cs, inverse);
  return check_launch("rope");
}
int rope_table(float2* cs, int max_pos, int hd, float theta, cudaStream_t s) {
  const int total = max_pos * (hd / 2);
  rope_table_kernel<<<blocks_for(total, 256), 256, 0, s>>>(cs, max_pos, hd / 2, theta);
  return check_launch("rope_table");
}
int swiglu_fwd(const bf16* gu, bf16* act, int64_t T, int F, cudaStream_t s) {
  if (F % 8 != 0) return -1;
  swiglu_fwd_kernel<<<blocks_for(T * (F >> 3), 256), 256, 0, s>>>(gu, act, T, F);
  return check_launch("swiglu_fwd");
}
int interleave_gate_up(const bf16* W, bf16* out, int F, int d, cudaStream_t s) {
  if (F % 128 != 0 || d % 8 != 0) return -1;
  interleave_gate_up_kernel<<<blocks_for(static_cast<int64_t>(2) * F * (d >> 3), 256), 256, 0, s>>>(W, out, F, d);
  return check_launch("interleave_gate_up");
}
int swiglu_bwd(const bf16* dact, const bf16* gu, bf16* dgu, int64_t T, int F, cudaStream_t s) {
  if (F % 8 != 0) return -1;
  swiglu_bwd_kernel<<<blocks_for(T * (F >> 3), 256), 256, 0, s>>>(dact, gu, dgu, T, F);
  return check_launch("swiglu_bwd");
}
int gelu_fwd(const bf16* x, bf16* y, int64_t n, cudaStream_t s) {
  if (n % 8 != 0) return -1;
  gelu_fwd_kernel<<<blocks_for(n >> 3, 256), 256, 0, s>>>(x, y, n >> 3);
  return check_launch("gelu_fwd");
}
int gelu_bwd(const bf16* dy, const bf16* x, bf16* dx, int64_t n, cudaStream_t s) {
  if (n % 8 != 0) return -1;
  gelu_bwd_kernel<<<blocks_for(n >> 3, 256), 256, 0, s>>>(dy, x, dx, n >> 3);
  return check_launch("gelu_bwd");
}
int embed_splice(const int32_t* ids, const uint8_t* tok_mask, const uint8_t* img_mask, const int32_t* img_of_row,
                 const bf16* table, const bf16* feats, bf16* out, uint8_t* out_mask, int32_t* img_pos, int N, int L, int P,
                 int d, int vocab, cudaStream_t s) {
  if (d % 8 != 0) return -1;
  find_image_pos_kernel<<<N, 32, 0, s>>>(ids, img_pos, N, L, -200);
  const int S = L - 1 + P;
  embed_splice_kernel<<<static_cast<unsigned>(static_cast<int64_t>(N) * S), 128, 0, s>>>(
      ids, tok_mask, img_mask, img_of_row, table, feats, out, out_mask, img_pos, L, P, d, vocab);
  return check_launch("embed_splice");
}
int embed_splice_bwd(const bf16* dX, const int32_t* img_of_row, const int32_t* img_pos, bf16* dfeats, int N, int S, int n_img,
                     int P, int d, cudaStream_t s) {
  embed_splice_bwd_kernel<<<static_cast<unsigned>(static_cast<int64_t>(n_img) * P), 128, 0, s>>>(dX, img_of_row, img_pos,
                                                                                                 dfeats, N, S, P, d);
  return check_launch("embed_splice_bwd");
}
int embed_rows(const int32_t* ids, const bf16* table, bf16* out, int rows, int d, int vocab, cudaStream_t s) {
  embed_rows_kernel<<<rows, 128, 0, s>>>(ids, table, out, d, vocab);
  return check_launch("embed_rows");
}
int clip_im2col(const float* images, bf16* patches, int n, int R, int ps, int Kp, cudaStream_t s) {
  const int g = R / ps;
  const int64_t total = static_cast<int64_t>(n) * g * g * Kp;
  clip_im2col_kernel<<<blocks_for(total, 256), 256, 0, s>>>(images, patches, n, R, ps, Kp);
  return check_launch("clip_im2col");
}
int clip_assemble(const bf16* patch_emb, const bf16* cls, const bf16* pos, bf16* out, int n, int P, int d, cudaStream_t s) {
  clip_assemble_kernel<<<static_cast<unsigned>(static_cast<int64_t>(n) * (P + 1)), 128, 0, s>>>(patch_emb, cls, pos, out, P, d);
  return check_launch("clip_assemble");
}
int clip_drop_cls(const bf16* x, bf16* out, int n, int P, int d, cudaStream_t s) {
  clip_drop_cls_kernel<<<static_cast<unsigned>(static_cast<int64_t>(n) * P), 128, 0, s>>>(x, out, P, d);
  return check_launch("clip_drop_cls");
}
int logprob_finalize(const float4* partials, int ntiles, const float* label_logit, const int32_t* labels, float* logprob,
                     float* entropy, float* lse, float* ent_raw, int64_t M, cudaStream_t s) {
  if (M == 0) return 0;
  logprob_finalize_kernel<<<blocks_for(M, 128), 128, 0, s>>>(partials, ntiles, label_logit, labels, logprob, entropy, lse,
                                                             ent_raw, M);
  return check_launch("logprob_finalize");
}
int dpo_pair_loss(const DpoPairArgs& a, cudaStream_t s) {
  dpo_pair_kernel<<<1, 1024, 0, s>>>(a);
  return check_launch("dpo_pair_loss");
}
int sumsq_accum(const float* g, int64_t n, float* out, cudaStream_t s) {
  if (n == 0) return 0;
  const unsigned blocks = capped_blocks(n / 4 + 1, 256, 148 * 8);
  sumsq_kernel<<<blocks, 256, 0, s>>>(g, n, out);
  return check_launch("sumsq");
}
int adamw_step(float* master, const float* grad, float* m, float* v, bf16* shadow, int64_t n, float lr, float b1, float b2,
               float eps, float wd, int step, float gscale, const float* sumsq, float max_norm, cudaStream_t s) {
  if (n == 0) return 0;
  const float bc1 = 1.f - powf(b1, static_cast<float>(step));
  const float bc2 = 1.f - powf(b2, static_cast<float>(step));
  const unsigned blocks = capped_blocks((n + 3) / 4, 256, 148 * 8);
  adamw_kernel<<<blocks, 256, 0, s>>>(master, grad, m, v, shadow, n, lr, b1, b2, eps, wd, bc1, bc2, gscale, sumsq, max_norm);
  return check_launch("adamw");
}
int f32_to_bf16(const float* src, bf16* dst, int64_t n, cudaStream_t s) {
  if (n == 0) return 0;
  f32_to_bf16_kernel<<<capped_blocks(n, 256, 148 * 16), 256, 0, s>>>(src, dst, n);
  return check_launch("f32_to_bf16");
}
int bf16_to_f32(const bf16* src, float* dst, int64_t n, cudaStream_t s) {
  if (n == 0) return 0;
  bf16_to_f32_kernel<<<capped_blocks(n, 256, 148 * 16), 256, 0, s>>>(src, dst, n);
  return check_launch("bf16_to_f32");
}
int norm_from_sumsq(const float* sumsq, float gscale, float* out, cudaStream_t s) {
  norm_from_sumsq_kernel<<<1, 1, 0, s>>>(sumsq, gscale, out);
  return check_launch("norm_from_sumsq");
}
int fill_affine(int32_t* out, int n, int stride, int offset, cudaStream_t s) {
  fill_affine_kernel<<<blocks_for(n, 128), 128, 0, s>>>(out, n, stride, offset);
  return check_launch("fill_affine");
}
int fill_mod(int32_t* out, int n, int mod, cudaStream_t s) {
  fill_mod_kernel<<<blocks_for(n, 128), 128, 0, s>>>(out, n, mod);
  return check_launch("fill_mod");
}
int queries_to_i32(const int64_t* queries, const uint8_t* qmask, int32_t* ids, uint8_t* mask, int64_t* out, int64_t out_ld,
                   int N, int Q, cudaStream_t s) {
  queries_to_i32_kernel<<<blocks_for(static_cast<int64_t>(N) * Q, 256), 256, 0, s>>>(queries, qmask, ids, mask, out, out_ld, N, Q);
  return check_launch("queries_to_i32");
}
int stack_inputs(const int64_t* queries, const uint8_t* qmask, const int64_t* const* responses, int K, int b, int Q, int R,
                 int pad_id, int packed, int32_t* ids, uint8_t* mask, int32_t* labels, cudaStream_t s) {
  if (K > 8) return -1;
  RespPtrs rp{};
  for (int k = 0; k < K; ++k) rp.p[k] = responses[k];
  const int64_t total = packed ? static_cast<int64_t>(b) * (Q + K * R) : static_cast<int64_t>(K) * b * (Q + R);
  stack_inputs_kernel<<<blocks_for(total, 256), 256, 0, s>>>(queries, qmask, rp, K, b, Q, R, pad_id, packed, ids, mask, labels);
  return check_launch("stack_inputs");
}
int build_positions(int32_t* pos, int N, int S, int S0, int R, int packed, cudaStream_t s) {
  build_positions_kernel<<<blocks_for(static_cast<int64_t>(N) * S, 256), 256, 0, s>>>(pos, N, S, S0, R, packed);
  return check_launch("build_positions");
}
int build_row_index(int32_t* src, int32_t* dst, int K, int b, int R, int S, int S0, int packed, int T, cudaStream_t s) {
  build_row_index_kernel<<<blocks_for(static_cast<int64_t>(K) * b * R, 256), 256, 0, s>>>(src, dst, K, b, R, S, S0, packed, T);
  return check_launch("build_row_index");
}
int fold_rows(bf16* dx, int K, int b, int S, int S0, int d, int T, cudaStream_t s) {
  if (K <= 1) return 0;
  fold_rows_kernel<<<b, 128, 0, s>>>(dx, K, b, S, S0, d, T);
  return check_launch("fold_rows");
}
int kv_store_prefill(const bf16* qkv, bf16* kc, bf16* vc, int N, int S, int S_max, int d, cudaStream_t s) {
  kv_store_prefill_kernel<<<static_cast<unsigned>(static_cast<int64_t>(N) * S), 128, 0, s>>>(qkv, kc, vc, S, S_max, d);
  return check_launch("kv_store_prefill");
}
int decode_attention(const bf16* qkv, bf16* kc, bf16* vc, const uint8_t* kmask, bf16* o, int N, int H, int hd, int len_base,
                     const int32_t* cur, int S_max, float scale, cudaStream_t s) {
  dim3 grid(H, N);
  if (hd == 128)
    decode_attention_kernel<128><<<grid, 256, 0, s>>>(qkv, kc, vc, kmask, o, H, len_base, cur, S_max, scale);
  else if (hd == 64)
    decode_attention_kernel<64><<<grid, 256, 0, s>>>(qkv, kc, vc, kmask, o, H, len_base, cur, S_max, scale);
  else
    return -1;
  return check_launch("decode_attention");
}
static SuppressList make_suppress(const int32_t* ids, int n) {
  SuppressList sl{};
  sl.n = n < 0 ? 0 : (n > OPADPO_MAX_SUPPRESS ? OPADPO_MAX_SUPPRESS : n);
  for (int i = 0; i < sl.n; ++i) sl.ids[i] = ids[i];
  return sl;
}
int greedy_pick(const float* partials, int splits, int N, int V, int eos_id, int pad_id, const int32_t* suppress,
                int n_suppress, uint8_t* done, int32_t* next_ids, int64_t* out, int64_t out_ld, const int32_t* cur,
                cudaStream_t s) {
  if (n_suppress < 0 || n_suppress > OPADPO_MAX_SUPPRESS) return -1;
  greedy_pick_kernel<<<N, 1024, 0, s>>>(partials, splits, static_cast<int64_t>(N) * V, V, eos_id, pad_id,
                                        make_suppress(suppress, n_suppress), done, next_ids, out, out_ld, cur);
  return check_launch("greedy_pick");
}
int sample_pick(const float* partials, int splits, int N, int V, int top_k, float top_p, float temperature, uint64_t seed,
                int eos_id, int pad_id, const int32_t* suppress, int n_suppress, uint8_t* done, int32_t* next_ids,
                int64_t* out, int64_t out_ld, const int32_t* cur, cudaStream_t s) {
  if (V > SAMPLE_THREADS * SAMPLE_MAX_PER_THREAD || top_k < 0 || !(top_p > 0.f) || !(temperature > 0.f)) return -1;
  if (n_suppress < 0 || n_suppress > OPADPO_MAX_SUPPRESS) return -1;
  sample_pick_kernel<<<N, SAMPLE_THREADS, 0, s>>>(partials, splits, static_cast<int64_t>(N) * V, V, top_k, top_p,
                                                  1.f / temperature, seed, eos_id, pad_id,
                                                  make_suppress(suppress, n_suppress), done, next_ids, out, out_ld, cur);
  return check_launch("sample_pick");
}
int advance_counter(int32_t* cur, cudaStream_t s) {
  advance_counter_kernel<<<1, 1, 0, s>>>(cur);
  return check_launch("advance_counter");
}
int partial_to_bf16(const float* part, int splits, int64_t n, bf16* out, cudaStream_t s) {
  if (n % 4 != 0) return -1;
  partial_to_bf16_kernel<<<blocks_for(n / 4, 128), 128, 0, s>>>(part, splits, n, out);
  return check_launch("partial_to_bf16");
}
int partial_residual_rmsnorm(const float* part, int splits, const bf16* res, bf16* x_out, const bf16* w, bf16* h, int64_t M,
                             int d, float eps, cudaStream_t s) {
  if (d % 4 != 0 || d > 4 * PRN_THREADS * 2) return -1;
  if (d <= 4 * PRN_THREADS)
    partial_residual_rmsnorm_kernel<1><<<static_cast<unsigned>(M), PRN_THREADS, 0, s>>>(part, splits, M * d, res, x_out, w, h, d, eps);
  else
    partial_residual_rmsnorm_kernel<2><<<static_cast<unsigned>(M), PRN_THREADS, 0, s>>>(part, splits, M * d, res, x_out, w, h, d, eps);
  return check_launch("partial_residual_rmsnorm");
}
int partial_qkv_rope(const float* part, int splits, bf16* qkv, int64_t M, int d, int hd, const float2* cs, int pos_base,
                     const int32_t* cur, cudaStream_t s) {
  if (hd % 8 != 0 || d % hd != 0 || d % 4 != 0) return -1;
  const int64_t items = M * ((2 * d / hd) * (hd >> 3) + (d >> 2));
  partial_qkv_rope_kernel<<<blocks_for(items, 256), 256, 0, s>>>(part, splits, M * 3 * d, qkv, M, d, hd, cs, pos_base, cur);
  return check_launch("partial_qkv_rope");
}
int partial_swiglu(const float* part, int splits, bf16* act, int64_t M, int F, cudaStream_t s) {
  partial_swiglu_kernel<<<blocks_for(M * F, 256), 256, 0, s>>>(part, splits, M * 2 * F, act, M, F);
  return check_launch("partial_swiglu");
}

}  // namespace opadpo
