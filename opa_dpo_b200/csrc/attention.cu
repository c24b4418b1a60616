// tcgen05 flash attention for sm_100a. One CTA = (sequence n, head h, tile); 6 warps:
//   warps 0-3 : softmax / epilogue, ONE THREAD PER TMEM LANE (= query row, or head-dim row for dK^T/dV^T)
//   warp 4    : TMA producer (one elected lane) + TMEM allocation
//   warp 5    : MMA issuer (one elected lane)
// Score tiles live in TMEM; P / dS are written by the softmax threads into 128B-swizzled smem and consumed by the
// next MMA either as a K-major A operand (P V, dS K) or, unchanged, as an MN-major B operand (dO^T P, Q^T dS).
#include "attention.h"

#include <cstdio>
#include <cstdlib>

#include "common.h"
#include "ptx.cuh"
#include "tmap.h"

namespace opadpo {
namespace {

constexpr int BQ = 128;   // query rows per tile
constexpr int BKV = 64;   // keys per tile
constexpr float LOG2E = 1.4426950408889634f;
constexpr float LN2 = 0.6931471805599453f;
constexpr int NTHREADS = 192;      // forward: 4 softmax warps + TMA warp + MMA warp
constexpr int NTHREADS_BWD = 320;  // backward: 8 softmax warps (two threads per TMEM lane, 32 score columns each) + TMA + MMA

struct AttnParams {
  int N, S, H;
  int nwords;
  int S0, segR;  // shared-prefix packing: keys [S0, segment start of the query) are invisible (segR == 0: off)
  float scale, scale_log2;
  const uint64_t* kbits;
  __nv_bfloat16* o;
  int64_t ld_o;
  float* lse;
  const float* delta;
  __nv_bfloat16 *dq, *dk, *dv;
  int64_t ld_dqkv;
  const float2* rope_cs;
  const int32_t* rope_pos;
};

// The (cos, sin) rows of a tile's 128 token rows, staged through shared memory. Read straight from global memory, every lane
// of a warp would walk its own 512 B table row 8 bytes at a time: 32 wavefronts per load instruction, ~8000 per CTA, all on
// the exposed tail of the kernel. Staged: each warp copies 16 rows with coalesced 16 B lanes (4 wavefronts per row); rows are
// padded to 132 floats so that the per-thread float4 reads afterwards are bank-conflict free.
constexpr int CS_ROW_FLOATS = 132;  // 64 (cos, sin) pairs + 4 floats of padding
constexpr int CS_STAGE_BYTES = 128 * CS_ROW_FLOATS * 4;
// warp w of the 8 epilogue warps stages rows [16 w, 16 w + 16); pos_of(r) = rotary position of tile row r
template <typename PosFn>
__device__ __forceinline__ void stage_rope_rows(float* cs_s, const float2* table, int half_hd, int warp, PosFn pos_of) {
  const int lane = threadIdx.x & 31;
#pragma unroll 4
  for (int r = warp * 16; r < warp * 16 + 16; ++r) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(table + static_cast<int64_t>(pos_of(r)) * half_hd) + lane);
    *reinterpret_cast<float4*>(cs_s + r * CS_ROW_FLOATS + lane * 4) = v;
  }
}

// 32 columns [c0, c0+32) and their rotary partners [c0+64, c0+96) of one row, fp32 accumulators -> bf16, with the optional
// inverse rotation (cs = this row's staged (cos, sin) entries for columns c0..c0+31, or null) and a common scale.
__device__ __forceinline__ void store_rotary_pair_runs(__nv_bfloat16* row, int c0, const uint32_t (&ta)[32], const uint32_t (&tb)[32],
                                                       float scale, const float* cs) {
  float a[32], b[32];
#pragma unroll
  for (int e = 0; e < 32; e += 2) {
    float4 k = make_float4(1.f, 0.f, 1.f, 0.f);  // (cos, sin) of columns e and e + 1; identity without a table
    if (cs) k = *reinterpret_cast<const float4*>(cs + 2 * e);
    const float x0 = __uint_as_float(ta[e]) * scale, y0 = __uint_as_float(tb[e]) * scale;
    const float x1 = __uint_as_float(ta[e + 1]) * scale, y1 = __uint_as_float(tb[e + 1]) * scale;
    a[e] = x0 * k.x + y0 * k.y;
    b[e] = y0 * k.x - x0 * k.y;
    a[e + 1] = x1 * k.z + y1 * k.w;
    b[e + 1] = y1 * k.z - x1 * k.w;
  }
  uint4* da = reinterpret_cast<uint4*>(row + c0);
  uint4* db = reinterpret_cast<uint4*>(row + c0 + 64);
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    uint4 u, w;
    u.x = ptx::pack_bf16(a[q * 8 + 0], a[q * 8 + 1]);
    u.y = ptx::pack_bf16(a[q * 8 + 2], a[q * 8 + 3]);
    u.z = ptx::pack_bf16(a[q * 8 + 4], a[q * 8 + 5]);
    u.w = ptx::pack_bf16(a[q * 8 + 6], a[q * 8 + 7]);
    w.x = ptx::pack_bf16(b[q * 8 + 0], b[q * 8 + 1]);
    w.y = ptx::pack_bf16(b[q * 8 + 2], b[q * 8 + 3]);
    w.z = ptx::pack_bf16(b[q * 8 + 4], b[q * 8 + 5]);
    w.w = ptx::pack_bf16(b[q * 8 + 6], b[q * 8 + 7]);
    da[q] = u;
    db[q] = w;
  }
}

// K-major tile of `rows` rows x (nblk*64) columns -> smem [blk][rows][128 B]
__device__ __forceinline__ void load_tile(uint8_t* dst, const CUtensorMap* m, uint64_t* bar, int col0, int row0,
                                          int rows, int nblk) {
  for (int b = 0; b < nblk; ++b) ptx::tma_load_2d(dst + b * rows * 128, m, bar, col0 + 64 * b, row0);
}

// Write 32 consecutive bf16 values (columns [half*32, half*32+32) of a 64-wide, 128 B row) into a swizzled tile.
__device__ __forceinline__ void store_half_row_sw128(uint8_t* tile, int r, int half, const float (&v)[32]) {
  uint8_t* rowp = tile + r * 128;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    uint4 u;
    u.x = ptx::pack_bf16(v[q * 8 + 0], v[q * 8 + 1]);
    u.y = ptx::pack_bf16(v[q * 8 + 2], v[q * 8 + 3]);
    u.z = ptx::pack_bf16(v[q * 8 + 4], v[q * 8 + 5]);
    u.w = ptx::pack_bf16(v[q * 8 + 6], v[q * 8 + 7]);
    const int chunk = (half * 4 + q) ^ (r & 7);
    *reinterpret_cast<uint4*>(rowp + chunk * 16) = u;
  }
}

// Validity word of key tile j for sequence n (bit c = key j*64+c may be attended to). Loaded one tile ahead of its use.
__device__ __forceinline__ uint64_t load_kword(const AttnParams& p, int n, int j) {
  return (p.kbits && j < p.nwords) ? __ldg(p.kbits + static_cast<int64_t>(n) * p.nwords + j) : ~0ull;
}
__device__ __forceinline__ uint64_t key_bits_w(const AttnParams& p, uint64_t bits, int j, int i, bool causal, bool row_valid) {
  const int rem = p.S - j * BKV;  // keys beyond the sequence end
  if (rem < 64) bits &= (rem <= 0) ? 0ull : ((1ull << rem) - 1ull);
  if (causal) {
    const int lim = i - j * BKV;  // keys c <= lim visible
    if (lim < 63) bits &= (lim < 0) ? 0ull : ((2ull << lim) - 1ull);
  }
  if (p.segR > 0 && i >= p.S0) {
    // packed row = [prefix | resp_0 | resp_1 ...]: a response token sees the prefix and its own segment only
    const int ex_hi = p.S0 + ((i - p.S0) / p.segR) * p.segR;
    const int lo = max(p.S0 - j * BKV, 0), hi = min(ex_hi - j * BKV, 64);
    if (hi > lo) {
      const uint64_t m = (hi >= 64 ? ~0ull : ((1ull << hi) - 1ull)) & ~((1ull << lo) - 1ull);
      bits &= ~m;
    }
  }
  return row_valid ? bits : 0ull;
}

// True when key tile j is invisible to EVERY row of the query tile starting at q0 (another response's segment).
__device__ __forceinline__ bool tile_skipped(const AttnParams& p, int q0, int j) {
  if (p.segR <= 0 || q0 < p.S0) return false;
  const int seg_start = p.S0 + ((q0 - p.S0) / p.segR) * p.segR;
  return j * BKV >= p.S0 && (j + 1) * BKV <= seg_start;
}

// The same rules with everything that does not depend on the key tile hoisted out of the tile loop (the segment arithmetic
// needs an integer division: ~40 dependent instructions per tile when left inside).
struct RowVis {   // what query row i may see, as key intervals
  int hi;         // keys >= hi are invisible (sequence end, causal diagonal; 0 for a row beyond the sequence)
  int ex_lo, ex_hi;  // keys in [ex_lo, ex_hi) are invisible (the other responses' segments in front of the row's own); empty if ex_hi <= ex_lo
};
__device__ __forceinline__ RowVis make_row_vis(const AttnParams& p, int i, bool causal) {
  RowVis v;
  v.hi = i < p.S ? (causal ? min(i + 1, p.S) : p.S) : 0;
  v.ex_lo = v.ex_hi = 0;
  if (p.segR > 0 && i >= p.S0) {
    v.ex_lo = p.S0;
    v.ex_hi = p.S0 + ((i - p.S0) / p.segR) * p.segR;
  }
  return v;
}
__device__ __forceinline__ uint64_t low_bits(int n) {  // n in [0, 64]
  return n >= 64 ? ~0ull : ((1ull << n) - 1ull);
}
__device__ __forceinline__ uint64_t key_bits_v(const RowVis& v, uint64_t bits, int j) {
  const int base = j * BKV;
  bits &= low_bits(min(max(v.hi - base, 0), 64));
  const int lo = min(max(v.ex_lo - base, 0), 64), hi = min(max(v.ex_hi - base, 0), 64);
  if (hi > lo) bits &= ~(low_bits(hi) & ~low_bits(lo));
  return bits;
}
// Key tiles [lo, hi) are invisible to every row of the query tile at q0 (tile_skipped as an interval).
struct TileSkip {
  int lo, hi;
};
__device__ __forceinline__ TileSkip make_tile_skip(const AttnParams& p, int q0) {
  TileSkip t{0, 0};
  if (p.segR > 0 && q0 >= p.S0) {
    const int seg_start = p.S0 + ((q0 - p.S0) / p.segR) * p.segR;
    t.lo = (p.S0 + BKV - 1) / BKV;
    t.hi = seg_start / BKV;
  }
  return t;
}
__device__ __forceinline__ int next_tile_after(const TileSkip& t, int j) {
  ++j;
  return (j >= t.lo && j < t.hi) ? t.hi : j;
}

// =====================================================================================================
// Forward
// =====================================================================================================
template <int HD>
struct FwdSmem {
  static constexpr int NB = HD / 64;
  static constexpr int Q_BYTES = NB * BQ * 128;
  static constexpr int KV_BYTES = NB * BKV * 128;
  static constexpr int P_BYTES = BQ * 128;
  static constexpr int OFF_Q = 0;
  static constexpr int OFF_K = OFF_Q + Q_BYTES;
  static constexpr int OFF_V = OFF_K + 2 * KV_BYTES;
  static constexpr int OFF_P = OFF_V + 2 * KV_BYTES;
  static constexpr int OFF_BAR = OFF_P + P_BYTES;
  static constexpr int TOTAL = OFF_BAR + 128;
  static constexpr int TMEM_COLS = 256;  // S double-buffered (2 x 64) + O (HD)
};

template <int HD, bool CAUSAL>
__global__ void __launch_bounds__(NTHREADS, 2)
    attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                    const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
  using L = FwdSmem<HD>;
  constexpr int NB = L::NB;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::OFF_BAR);
  uint64_t* q_full = bars + 0;
  uint64_t* kv_full = bars + 1;   // [2]
  uint64_t* kv_empty = bars + 3;  // [2]
  uint64_t* s_full = bars + 5;    // [2]: S is double-buffered in TMEM so Q K^T of tile j+1 runs under the softmax of tile j
  uint64_t* p_full = bars + 7;
  uint64_t* o_full = bars + 8;
  uint64_t* pv_done = bars + 9;   // P V of the previous tile has completed: the P buffer and O may be touched again
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);

  const int warp = threadIdx.x >> 5;
  const int qt = CAUSAL ? (gridDim.x - 1 - blockIdx.x) : blockIdx.x;
  const int h = blockIdx.y, n = blockIdx.z;
  const int q0 = qt * BQ;
  const int row0 = n * p.S + q0;
  const int col0 = h * HD;
  int num_kv = (p.S + BKV - 1) / BKV;
  if (CAUSAL) num_kv = min(num_kv, (q0 + BQ + BKV - 1) / BKV);

  if (warp == 4) {
    if (ptx::elect_one()) {
      ptx::prefetch_tmap(&tmQ);
      ptx::prefetch_tmap(&tmK);
      ptx::prefetch_tmap(&tmV);
    }
    __syncwarp();
    ptx::tmem_alloc(tmem_slot, L::TMEM_COLS);
    ptx::tmem_relinquish();
  } else if (warp == 5) {
    if (ptx::elect_one()) {
      if ((ptx::smem_u32(smem) & 1023u) != 0) __trap();
      ptx::mbar_init(q_full, 1);
      for (int i = 0; i < 2; ++i) {
        ptx::mbar_init(&kv_full[i], 1);
        ptx::mbar_init(&kv_empty[i], 1);
      }
      ptx::mbar_init(&s_full[0], 1);
      ptx::mbar_init(&s_full[1], 1);
      ptx::mbar_init(p_full, 128);
      ptx::mbar_init(o_full, 1);
      ptx::mbar_init(pv_done, 1);
      ptx::fence_barrier_init();
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t t_s = tmem, t_o = tmem + 128;  // S buffers at columns [0, 64) and [64, 128)

  if (warp == 4) {
    if (ptx::elect_one()) {
      ptx::mbar_arrive_expect_tx(q_full, L::Q_BYTES);
      load_tile(smem + L::OFF_Q, &tmQ, q_full, col0, row0, BQ, NB);
      int it = 0;
      for (int j = 0; j < num_kv; ++j) {
        if (tile_skipped(p, q0, j)) continue;
        const int st = it & 1;
        ptx::mbar_wait(&kv_empty[st], ((it >> 1) & 1) ^ 1);
        ptx::mbar_arrive_expect_tx(&kv_full[st], 2 * L::KV_BYTES);
        load_tile(smem + L::OFF_K + st * L::KV_BYTES, &tmK, &kv_full[st], col0, n * p.S + j * BKV, BKV, NB);
        load_tile(smem + L::OFF_V + st * L::KV_BYTES, &tmV, &kv_full[st], col0, n * p.S + j * BKV, BKV, NB);
        ++it;
      }
    }
  } else if (warp == 5) {
    if (ptx::elect_one()) {
      constexpr uint32_t idesc_qk = ptx::make_idesc_bf16(BQ, BKV, false, false);
      constexpr uint32_t idesc_pv = ptx::make_idesc_bf16(BQ, HD, false, true);
      const uint32_t aQ = ptx::smem_u32(smem + L::OFF_Q);
      const uint32_t aP = ptx::smem_u32(smem + L::OFF_P);
      ptx::mbar_wait(q_full, 0);
      int n_tiles = 0;
      for (int j = 0; j < num_kv; ++j) n_tiles += tile_skipped(p, q0, j) ? 0 : 1;
      // S = Q K^T of pipeline slot `it` (K/V stage it & 1) into S buffer it & 1
      auto issue_qk = [&](int it) {
        const int st = it & 1;
        const uint32_t aK = ptx::smem_u32(smem + L::OFF_K + st * L::KV_BYTES);
        ptx::mbar_wait(&kv_full[st], (it >> 1) & 1);
        ptx::tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < HD / 16; ++kk) {
          const int blk = kk >> 2, off = (kk & 3) * 32;
          ptx::mma_bf16_ss(t_s + st * 64, ptx::make_smem_desc(aQ + blk * (BQ * 128) + off, 16, 1024),
                           ptx::make_smem_desc(aK + blk * (BKV * 128) + off, 16, 1024), idesc_qk, kk > 0);
        }
        ptx::mma_commit(&s_full[st]);
      };
      if (n_tiles > 0) issue_qk(0);
      for (int it = 0; it < n_tiles; ++it) {
        // S buffer (it+1)&1 was last read by the softmax of slot it-1, whose p_full was awaited one iteration ago; the K/V
        // stage was released by P V of slot it-1 (the producer waits for that before it refills the stage)
        if (it + 1 < n_tiles) issue_qk(it + 1);
        const int st = it & 1;
        const uint32_t aV = ptx::smem_u32(smem + L::OFF_V + st * L::KV_BYTES);
        ptx::mbar_wait(p_full, it & 1);
        ptx::tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < BKV / 16; ++kk) {
          ptx::mma_bf16_ss(t_o, ptx::make_smem_desc(aP + kk * 32, 16, 1024),
                           ptx::make_smem_desc(aV + kk * 2048, BKV * 128, 1024), idesc_pv, (it > 0 || kk > 0));
        }
        ptx::mma_commit(&kv_empty[st]);
        ptx::mma_commit(pv_done);
        if (it == n_tiles - 1) ptx::mma_commit(o_full);
      }
    }
  } else {
    // ---------------------------------------------------------------- softmax: thread == query row
    const int tid = threadIdx.x;
    const int i = q0 + tid;
    const bool row_valid = i < p.S;
    const uint32_t lane_off = static_cast<uint32_t>(warp * 32) << 16;
    float m_ref = -INFINITY, l = 0.f;
    int it = 0;
    // the key-validity word of the next processed tile is fetched while this one is being worked on
    auto next_j = [&](int j) {
      ++j;
      while (j < num_kv && tile_skipped(p, q0, j)) ++j;
      return j;
    };
    int j = next_j(-1);
    uint64_t kw = j < num_kv ? load_kword(p, n, j) : 0ull;
    for (; j < num_kv;) {
      const int jn = next_j(j);
      const uint64_t kw_next = jn < num_kv ? load_kword(p, n, jn) : 0ull;
      const uint64_t bits = key_bits_w(p, kw, j, i, CAUSAL, row_valid);
      // interior tiles (no padding, no causal edge, no foreign segment in any row of this warp) skip the per-element masking
      const bool full = __all_sync(0xffffffffu, bits == ~0ull);
      ptx::mbar_wait(&s_full[it & 1], (it >> 1) & 1);
      ptx::tc_fence_after();
      uint32_t r0[32], r1[32];
      ptx::tmem_ld32(t_s + (it & 1) * 64 + lane_off, r0);
      ptx::tmem_ld32(t_s + (it & 1) * 64 + lane_off + 32, r1);
      ptx::tmem_ld_wait();
      float cm0 = -INFINITY, cm1 = -INFINITY;
      if (full) {
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          cm0 = fmaxf(cm0, __uint_as_float(r0[c]));
          cm1 = fmaxf(cm1, __uint_as_float(r1[c]));
        }
      } else {
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          const float x = ((bits >> c) & 1ull) ? __uint_as_float(r0[c]) : -INFINITY;
          const float y = ((bits >> (32 + c)) & 1ull) ? __uint_as_float(r1[c]) : -INFINITY;
          r0[c] = __float_as_uint(x);
          r1[c] = __float_as_uint(y);
          cm0 = fmaxf(cm0, x);
          cm1 = fmaxf(cm1, y);
        }
      }
      const float cm = fmaxf(cm0, cm1) * p.scale_log2;  // scale > 0: max(s) * scale == max(s * scale); -inf stays -inf
      // Lazy rescale: only move the reference max when it grows by more than 2^8 (exact result either way).
      const bool need = cm > m_ref + 8.f;
      float alpha = 1.f;
      if (need) {
        alpha = ptx::ex2_approx(m_ref - cm);
        m_ref = cm;
        l *= alpha;
      }
      if (it > 0) {  // P V of the previous tile must be done before O is rescaled or the P buffer is overwritten
        ptx::mbar_wait(pv_done, (it - 1) & 1);
        ptx::tc_fence_after();
      }
      if (it > 0 && __any_sync(0xffffffffu, need)) {
#pragma unroll 1
        for (int c = 0; c < HD / 32; ++c) {
          uint32_t t[32];
          ptx::tmem_ld32(t_o + lane_off + c * 32, t);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 32; ++e) t[e] = __float_as_uint(__uint_as_float(t[e]) * alpha);
          ptx::tmem_st32(t_o + lane_off + c * 32, t);
        }
        ptx::tmem_st_wait();
      }
      // p = 2^(s * scale_log2 - m): one FFMA and one MUFU per element; four independent partial sums
      const float nmu = (m_ref == -INFINITY) ? 0.f : -m_ref;
      const float sc2 = p.scale_log2;
      float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
      float pv[32];
#pragma unroll
      for (int c = 0; c < 32; c += 4) {
        pv[c] = ptx::ex2_approx(fmaf(__uint_as_float(r0[c]), sc2, nmu));
        pv[c + 1] = ptx::ex2_approx(fmaf(__uint_as_float(r0[c + 1]), sc2, nmu));
        pv[c + 2] = ptx::ex2_approx(fmaf(__uint_as_float(r0[c + 2]), sc2, nmu));
        pv[c + 3] = ptx::ex2_approx(fmaf(__uint_as_float(r0[c + 3]), sc2, nmu));
        s0 += pv[c];
        s1 += pv[c + 1];
        s2 += pv[c + 2];
        s3 += pv[c + 3];
      }
      store_half_row_sw128(smem + L::OFF_P, tid, 0, pv);
#pragma unroll
      for (int c = 0; c < 32; c += 4) {
        pv[c] = ptx::ex2_approx(fmaf(__uint_as_float(r1[c]), sc2, nmu));
        pv[c + 1] = ptx::ex2_approx(fmaf(__uint_as_float(r1[c + 1]), sc2, nmu));
        pv[c + 2] = ptx::ex2_approx(fmaf(__uint_as_float(r1[c + 2]), sc2, nmu));
        pv[c + 3] = ptx::ex2_approx(fmaf(__uint_as_float(r1[c + 3]), sc2, nmu));
        s0 += pv[c];
        s1 += pv[c + 1];
        s2 += pv[c + 2];
        s3 += pv[c + 3];
      }
      store_half_row_sw128(smem + L::OFF_P, tid, 1, pv);
      l += (s0 + s1) + (s2 + s3);
      ptx::fence_proxy_async();
      ptx::tc_fence_before();
      ptx::mbar_arrive(p_full);
      ++it;
      j = jn;
      kw = kw_next;
    }
    ptx::mbar_wait(o_full, 0);
    ptx::tc_fence_after();
    const float inv = l > 0.f ? 1.f / l : 0.f;
    __nv_bfloat16* orow = p.o + static_cast<int64_t>(row0 + tid) * p.ld_o + col0;
#pragma unroll 1
    for (int c = 0; c < HD / 32; ++c) {
      uint32_t t[32];
      ptx::tmem_ld32(t_o + lane_off + c * 32, t);
      ptx::tmem_ld_wait();
      if (row_valid) {
        uint4* dst = reinterpret_cast<uint4*>(orow + c * 32);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint4 u;
          u.x = ptx::pack_bf16(__uint_as_float(t[q * 8 + 0]) * inv, __uint_as_float(t[q * 8 + 1]) * inv);
          u.y = ptx::pack_bf16(__uint_as_float(t[q * 8 + 2]) * inv, __uint_as_float(t[q * 8 + 3]) * inv);
          u.z = ptx::pack_bf16(__uint_as_float(t[q * 8 + 4]) * inv, __uint_as_float(t[q * 8 + 5]) * inv);
          u.w = ptx::pack_bf16(__uint_as_float(t[q * 8 + 6]) * inv, __uint_as_float(t[q * 8 + 7]) * inv);
          dst[q] = u;
        }
      }
      __syncwarp();
    }
    if (row_valid && p.lse)
      p.lse[static_cast<int64_t>(row0 + tid) * p.H + h] = l > 0.f ? (m_ref + log2f(l)) * LN2 : INFINITY;
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem, L::TMEM_COLS);
  }
}

// =====================================================================================================
// Forward, P kept in tensor memory. Same tiling as attn_fwd_kernel (128 query rows per CTA, 64-key tiles, two CTAs per SM),
// but (1) the softmax threads write P as bf16 pairs straight back into the S columns they just read (tcgen05.st) and
// O += P V takes its A operand from TMEM, so there is no shared-memory P buffer, no generic->async proxy fence, and P is
// double-buffered for free: the exponentials of tile j+1 no longer wait for P V of tile j; (2) K and V travel in separate
// two-stage rings, a K stage being released as soon as Q K^T has read it, so the load of K(j+2) starts a whole tile earlier
// than when it shared a stage with V. O is only touched by the softmax threads when the running maximum moves (lazy
// rescale), which is the one place that still waits for the previous P V.
// =====================================================================================================
template <int HD>
struct FwdTsSmem {
  static constexpr int NB = HD / 64;
  static constexpr int Q_BYTES = NB * BQ * 128;
  static constexpr int KV_BYTES = NB * BKV * 128;
  static constexpr int OFF_Q = 0;
  static constexpr int OFF_K = OFF_Q + Q_BYTES;      // [2]
  static constexpr int OFF_V = OFF_K + 2 * KV_BYTES;  // [2]
  static constexpr int OFF_BAR = OFF_V + 2 * KV_BYTES;
  static constexpr int TOTAL = OFF_BAR + 256;
  static constexpr int TMEM_COLS = 256;  // S / P double-buffered (2 x 64) + O (HD)
};

template <int HD, bool CAUSAL>
__global__ void __launch_bounds__(NTHREADS, 2)
    attn_fwd_ts_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                       const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
  using L = FwdTsSmem<HD>;
  constexpr int NB = L::NB;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::OFF_BAR);
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;    // [2]
  uint64_t* k_empty = bars + 3;   // [2]
  uint64_t* v_full = bars + 5;    // [2]
  uint64_t* v_empty = bars + 7;   // [2]
  uint64_t* s_full = bars + 9;    // [2]
  uint64_t* p_full = bars + 11;   // [2]
  uint64_t* pv_done = bars + 13;
  uint64_t* o_full = bars + 14;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 15);

  const int warp = threadIdx.x >> 5;
  const int qt = CAUSAL ? (gridDim.x - 1 - blockIdx.x) : blockIdx.x;
  const int h = blockIdx.y, n = blockIdx.z;
  const int q0 = qt * BQ;
  const int row0 = n * p.S + q0;
  const int col0 = h * HD;
  int num_kv = (p.S + BKV - 1) / BKV;
  if (CAUSAL) num_kv = min(num_kv, (q0 + BQ + BKV - 1) / BKV);
  const TileSkip skip = make_tile_skip(p, q0);
  auto next_j = [&](int j) { return next_tile_after(skip, j); };  // next processed key tile after j (>= num_kv: none)

  if (warp == 4) {
    if (ptx::elect_one()) {
      ptx::prefetch_tmap(&tmQ);
      ptx::prefetch_tmap(&tmK);
      ptx::prefetch_tmap(&tmV);
    }
    __syncwarp();
    ptx::tmem_alloc(tmem_slot, L::TMEM_COLS);
    ptx::tmem_relinquish();
  } else if (warp == 5) {
    if (ptx::elect_one()) {
      if ((ptx::smem_u32(smem) & 1023u) != 0) __trap();
      ptx::mbar_init(q_full, 1);
      for (int i = 0; i < 2; ++i) {
        ptx::mbar_init(&k_full[i], 1);
        ptx::mbar_init(&k_empty[i], 1);
        ptx::mbar_init(&v_full[i], 1);
        ptx::mbar_init(&v_empty[i], 1);
        ptx::mbar_init(&s_full[i], 1);
        ptx::mbar_init(&p_full[i], 128);
      }
      ptx::mbar_init(pv_done, 1);
      ptx::mbar_init(o_full, 1);
      ptx::fence_barrier_init();
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t t_s = tmem, t_o = tmem + 128;  // S / P buffers at columns [0, 64) and [64, 128)

  if (warp == 4) {
    if (ptx::elect_one()) {
      ptx::mbar_arrive_expect_tx(q_full, L::Q_BYTES);
      load_tile(smem + L::OFF_Q, &tmQ, q_full, col0, row0, BQ, NB);
      int it = 0;
      for (int j = next_j(-1); j < num_kv; j = next_j(j), ++it) {
        const int st = it & 1;
        const uint32_t ph = ((it >> 1) & 1) ^ 1;
        ptx::mbar_wait(&k_empty[st], ph);
        ptx::mbar_arrive_expect_tx(&k_full[st], L::KV_BYTES);
        load_tile(smem + L::OFF_K + st * L::KV_BYTES, &tmK, &k_full[st], col0, n * p.S + j * BKV, BKV, NB);
        ptx::mbar_wait(&v_empty[st], ph);
        ptx::mbar_arrive_expect_tx(&v_full[st], L::KV_BYTES);
        load_tile(smem + L::OFF_V + st * L::KV_BYTES, &tmV, &v_full[st], col0, n * p.S + j * BKV, BKV, NB);
      }
    }
  } else if (warp == 5) {
    if (ptx::elect_one()) {
      constexpr uint32_t idesc_qk = ptx::make_idesc_bf16(BQ, BKV, false, false);
      constexpr uint32_t idesc_pv = ptx::make_idesc_bf16(BQ, HD, false, true);
      const uint32_t aQ = ptx::smem_u32(smem + L::OFF_Q);
      ptx::mbar_wait(q_full, 0);
      int n_tiles = 0;
      for (int j = next_j(-1); j < num_kv; j = next_j(j)) ++n_tiles;
      // S = Q K^T of pipeline slot `it` into S buffer it & 1; the K stage is free again as soon as these MMAs are done
      auto issue_qk = [&](int it) {
        const int st = it & 1;
        const uint32_t aK = ptx::smem_u32(smem + L::OFF_K + st * L::KV_BYTES);
        ptx::mbar_wait(&k_full[st], (it >> 1) & 1);
        ptx::tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < HD / 16; ++kk) {
          const int blk = kk >> 2, off = (kk & 3) * 32;
          ptx::mma_bf16_ss(t_s + st * 64, ptx::make_smem_desc(aQ + blk * (BQ * 128) + off, 16, 1024),
                           ptx::make_smem_desc(aK + blk * (BKV * 128) + off, 16, 1024), idesc_qk, kk > 0);
        }
        ptx::mma_commit(&s_full[st]);
        ptx::mma_commit(&k_empty[st]);
      };
      if (n_tiles > 0) issue_qk(0);
      for (int it = 0; it < n_tiles; ++it) {
        // S buffer (it+1)&1 still holds P of slot it-1: P V of that slot was issued one iteration ago, and the tensor pipe
        // executes in issue order, so Q K^T of slot it+1 overwrites it only after it has been consumed
        if (it + 1 < n_tiles) issue_qk(it + 1);
        const int b = it & 1;
        const uint32_t aV = ptx::smem_u32(smem + L::OFF_V + b * L::KV_BYTES);
        ptx::mbar_wait(&v_full[b], (it >> 1) & 1);
        ptx::mbar_wait(&p_full[b], (it >> 1) & 1);
        ptx::tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < BKV / 16; ++kk) {  // A = P from TMEM: 16 keys = 8 columns of bf16 pairs
          ptx::mma_bf16_ts(t_o, t_s + b * 64 + kk * 8, ptx::make_smem_desc(aV + kk * 2048, BKV * 128, 1024), idesc_pv,
                           (it > 0 || kk > 0));
        }
        ptx::mma_commit(&v_empty[b]);
        ptx::mma_commit(pv_done);
        if (it == n_tiles - 1) ptx::mma_commit(o_full);
      }
    }
  } else {
    // ---------------------------------------------------------------- softmax: thread == query row
    const int tid = threadIdx.x;
    const int i = q0 + tid;
    const bool row_valid = i < p.S;
    const uint32_t lane_off = static_cast<uint32_t>(warp * 32) << 16;
    const RowVis vis = make_row_vis(p, i, CAUSAL);
    float m_ref = -INFINITY, l = 0.f;
    int it = 0;
    int j = next_j(-1);
    uint64_t kw = j < num_kv ? load_kword(p, n, j) : 0ull;
    for (; j < num_kv; ++it) {
      const int jn = next_j(j);
      const uint64_t kw_next = jn < num_kv ? load_kword(p, n, jn) : 0ull;
      const uint64_t bits = key_bits_v(vis, kw, j);
      const bool full = __all_sync(0xffffffffu, bits == ~0ull);
      const int b = it & 1;
      ptx::mbar_wait(&s_full[b], (it >> 1) & 1);
      ptx::tc_fence_after();
      uint32_t r0[32], r1[32];
      ptx::tmem_ld32(t_s + b * 64 + lane_off, r0);
      ptx::tmem_ld32(t_s + b * 64 + lane_off + 32, r1);
      ptx::tmem_ld_wait();
      float cm0 = -INFINITY, cm1 = -INFINITY;
      if (full) {
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          cm0 = fmaxf(cm0, __uint_as_float(r0[c]));
          cm1 = fmaxf(cm1, __uint_as_float(r1[c]));
        }
      } else {
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          const float x = ((bits >> c) & 1ull) ? __uint_as_float(r0[c]) : -INFINITY;
          const float y = ((bits >> (32 + c)) & 1ull) ? __uint_as_float(r1[c]) : -INFINITY;
          r0[c] = __float_as_uint(x);
          r1[c] = __float_as_uint(y);
          cm0 = fmaxf(cm0, x);
          cm1 = fmaxf(cm1, y);
        }
      }
      const float cm = fmaxf(cm0, cm1) * p.scale_log2;
      const bool need = cm > m_ref + 8.f;  // lazy rescale: move the reference maximum only when it grows by more than 2^8
      float alpha = 1.f;
      if (need) {
        alpha = ptx::ex2_approx(m_ref - cm);
        m_ref = cm;
        l *= alpha;
      }
      if (it > 0 && __any_sync(0xffffffffu, need)) {
        // O may only be touched between P V of slot it-1 (complete: phase it-1 of pv_done; the barrier cannot run further
        // ahead than that before this thread has delivered P of slot it) and P V of slot it (waits for this thread's arrive)
        ptx::mbar_wait(pv_done, (it - 1) & 1);
        ptx::tc_fence_after();
#pragma unroll 1
        for (int c = 0; c < HD / 32; ++c) {
          uint32_t t[32];
          ptx::tmem_ld32(t_o + lane_off + c * 32, t);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 32; ++e) t[e] = __float_as_uint(__uint_as_float(t[e]) * alpha);
          ptx::tmem_st32(t_o + lane_off + c * 32, t);
        }
      }
      const float nmu = (m_ref == -INFINITY) ? 0.f : -m_ref;
      const float sc2 = p.scale_log2;
      float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
      uint32_t pk[32];  // column c of the P buffer = keys 2c, 2c+1 of this row as a bf16 pair
#pragma unroll
      for (int c = 0; c < 32; c += 4) {
        const float a0 = ptx::ex2_approx(fmaf(__uint_as_float(r0[c]), sc2, nmu));
        const float a1 = ptx::ex2_approx(fmaf(__uint_as_float(r0[c + 1]), sc2, nmu));
        const float a2 = ptx::ex2_approx(fmaf(__uint_as_float(r0[c + 2]), sc2, nmu));
        const float a3 = ptx::ex2_approx(fmaf(__uint_as_float(r0[c + 3]), sc2, nmu));
        s0 += a0;
        s1 += a1;
        s2 += a2;
        s3 += a3;
        pk[c / 2] = ptx::pack_bf16(a0, a1);
        pk[c / 2 + 1] = ptx::pack_bf16(a2, a3);
      }
#pragma unroll
      for (int c = 0; c < 32; c += 4) {
        const float a0 = ptx::ex2_approx(fmaf(__uint_as_float(r1[c]), sc2, nmu));
        const float a1 = ptx::ex2_approx(fmaf(__uint_as_float(r1[c + 1]), sc2, nmu));
        const float a2 = ptx::ex2_approx(fmaf(__uint_as_float(r1[c + 2]), sc2, nmu));
        const float a3 = ptx::ex2_approx(fmaf(__uint_as_float(r1[c + 3]), sc2, nmu));
        s0 += a0;
        s1 += a1;
        s2 += a2;
        s3 += a3;
        pk[16 + c / 2] = ptx::pack_bf16(a0, a1);
        pk[16 + c / 2 + 1] = ptx::pack_bf16(a2, a3);
      }
      l += (s0 + s1) + (s2 + s3);
      ptx::tmem_st32(t_s + b * 64 + lane_off, pk);
      ptx::tmem_st_wait();
      ptx::tc_fence_before();
      ptx::mbar_arrive(&p_full[b]);
      j = jn;
      kw = kw_next;
    }
    ptx::mbar_wait(o_full, 0);
    ptx::tc_fence_after();
    const float inv = l > 0.f ? 1.f / l : 0.f;
    __nv_bfloat16* orow = p.o + static_cast<int64_t>(row0 + tid) * p.ld_o + col0;
#pragma unroll 1
    for (int c = 0; c < HD / 32; ++c) {
      uint32_t t[32];
      ptx::tmem_ld32(t_o + lane_off + c * 32, t);
      ptx::tmem_ld_wait();
      if (row_valid) {
        uint4* dst = reinterpret_cast<uint4*>(orow + c * 32);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint4 u;
          u.x = ptx::pack_bf16(__uint_as_float(t[q * 8 + 0]) * inv, __uint_as_float(t[q * 8 + 1]) * inv);
          u.y = ptx::pack_bf16(__uint_as_float(t[q * 8 + 2]) * inv, __uint_as_float(t[q * 8 + 3]) * inv);
          u.z = ptx::pack_bf16(__uint_as_float(t[q * 8 + 4]) * inv, __uint_as_float(t[q * 8 + 5]) * inv);
          u.w = ptx::pack_bf16(__uint_as_float(t[q * 8 + 6]) * inv, __uint_as_float(t[q * 8 + 7]) * inv);
          dst[q] = u;
        }
      }
      __syncwarp();
    }
    if (row_valid && p.lse)
      p.lse[static_cast<int64_t>(row0 + tid) * p.H + h] = l > 0.f ? (m_ref + log2f(l)) * LN2 : INFINITY;
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem, L::TMEM_COLS);
  }
}

// =====================================================================================================
// Backward, shared score recomputation: p = exp2(s*scale_log2 - lse2), ds' = p*(dp - delta). The softmax scale that
// belongs on ds is applied once per output element in the dQ / dK epilogues instead of once per score.
// =====================================================================================================
template <bool WANT_P>
__device__ __forceinline__ void recompute_half(const uint32_t (&s)[32], const uint32_t (&dp)[32], uint64_t bits, bool full,
                                               int half, float scale_log2, float lse2, float delta, float (&pv)[32],
                                               float (&dsv)[32]) {
  const float nl = -lse2;  // invalid rows carry lse2 = +inf: every p is 2^-inf = 0
  if (full) {              // interior tile: no masking, one FFMA + MUFU + FADD + FMUL per score
#pragma unroll
    for (int c = 0; c < 32; ++c) {
      const float pr = ptx::ex2_approx(fmaf(__uint_as_float(s[c]), scale_log2, nl));
      if (WANT_P) pv[c] = pr;
      dsv[c] = pr * (__uint_as_float(dp[c]) - delta);
    }
  } else {
    const uint32_t hb = static_cast<uint32_t>(bits >> (half * 32));
#pragma unroll
    for (int c = 0; c < 32; ++c) {
      const float e = ptx::ex2_approx(fmaf(__uint_as_float(s[c]), scale_log2, nl));
      const float pr = ((hb >> c) & 1u) ? e : 0.f;
      if (WANT_P) pv[c] = pr;
      dsv[c] = pr * (__uint_as_float(dp[c]) - delta);
    }
  }
}

// ---------------------------------------------------------------- dQ: one CTA per query tile, loops over key tiles
template <int HD>
struct DqSmem {
  static constexpr int NB = HD / 64;
  static constexpr int Q_BYTES = NB * BQ * 128;
  static constexpr int KV_BYTES = NB * BKV * 128;
  static constexpr int OFF_Q = 0;
  static constexpr int OFF_DO = OFF_Q + Q_BYTES;
  // One CTA per SM, pipelined inside the CTA (as the dK/dV kernel): three K/V stages keep two tile loads in flight, S/dP
  // live in two TMEM buffers and dS in two smem buffers, so S,dP of key tile j+1 are computed while the softmax warps turn
  // tile j into dS and dQ += dS_j K_j follows as soon as that is written.
  // dS never touches shared memory: the softmax threads write it as bf16 pairs back into the S columns they read
  // (tcgen05.st) and dQ += dS K takes its A operand from TMEM, which also leaves room for a fourth K/V stage.
  static constexpr int KV_STAGES = 4;
  static constexpr int OFF_O = OFF_DO + Q_BYTES;  // the forward output tile: delta = rowsum(dO * O) is computed here, once per row
  static constexpr int OFF_K = OFF_O + Q_BYTES;
  static constexpr int OFF_V = OFF_K + KV_STAGES * KV_BYTES;
  static constexpr int OFF_DELTA = OFF_V + KV_STAGES * KV_BYTES;  // float [2][BQ]: per-row partial dot products of the two column halves
  static constexpr int OFF_BAR = OFF_DELTA + 2 * BQ * 4;
  static constexpr int TOTAL = OFF_BAR + 256;
  static constexpr int TMEM_COLS = 512;  // [S0 (dS0 aliased) | dP0 | S1 (dS1) | dP1] (4 x 64) + dQ (HD)
};

template <int HD, bool CAUSAL>
__global__ void __launch_bounds__(NTHREADS_BWD, 1)
    attn_bwd_dq_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmdO,
                       const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmK,
                       const __grid_constant__ CUtensorMap tmV, const AttnParams p, float* __restrict__ delta_out) {
  using L = DqSmem<HD>;
  constexpr int NB = L::NB;
  constexpr int NS = L::KV_STAGES;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::OFF_BAR);
  uint64_t* q_full = bars + 0;
  uint64_t* kv_full = bars + 1;    // [4]
  uint64_t* kv_empty = bars + 5;   // [4]
  uint64_t* sdp_full = bars + 9;   // [2]
  uint64_t* ds_full = bars + 11;   // [2]
  uint64_t* dq_done = bars + 13;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);
  uint64_t* o_full = bars + 15;    // the O tile (delta only): loaded behind the first K/V stage so it never delays the first MMAs

  const int warp = threadIdx.x >> 5;
  const int qt = CAUSAL ? (gridDim.x - 1 - blockIdx.x) : blockIdx.x;
  const int h = blockIdx.y, n = blockIdx.z;
  const int q0 = qt * BQ;
  const int row0 = n * p.S + q0;
  const int col0 = h * HD;
  int num_kv = (p.S + BKV - 1) / BKV;
  if (CAUSAL) num_kv = min(num_kv, (q0 + BQ + BKV - 1) / BKV);
  const TileSkip skip = make_tile_skip(p, q0);
  auto next_j = [&](int j) { return next_tile_after(skip, j); };  // next processed key tile after j (>= num_kv: none)

  if (warp == 8) {
    ptx::tmem_alloc(tmem_slot, L::TMEM_COLS);
    ptx::tmem_relinquish();
  } else if (warp == 9) {
    if (ptx::elect_one()) {
      if ((ptx::smem_u32(smem) & 1023u) != 0) __trap();
      ptx::mbar_init(q_full, 1);
      for (int i = 0; i < NS; ++i) {
        ptx::mbar_init(&kv_full[i], 1);
        ptx::mbar_init(&kv_empty[i], 1);
      }
      for (int i = 0; i < 2; ++i) {
        ptx::mbar_init(&sdp_full[i], 1);
        ptx::mbar_init(&ds_full[i], 256);
      }
      ptx::mbar_init(dq_done, 1);
      ptx::mbar_init(o_full, 1);
      ptx::fence_barrier_init();
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t t_dq = tmem + 256;

  if (warp == 8) {
    if (ptx::elect_one()) {
      ptx::mbar_arrive_expect_tx(q_full, 2 * L::Q_BYTES);
      load_tile(smem + L::OFF_Q, &tmQ, q_full, col0, row0, BQ, NB);
      load_tile(smem + L::OFF_DO, &tmdO, q_full, col0, row0, BQ, NB);
      auto load_o = [&]() {
        ptx::mbar_arrive_expect_tx(o_full, L::Q_BYTES);
        load_tile(smem + L::OFF_O, &tmO, o_full, col0, row0, BQ, NB);
      };
      int it = 0;
      for (int j = next_j(-1); j < num_kv; j = next_j(j), ++it) {
        const int st = it % NS;
        ptx::mbar_wait(&kv_empty[st], ((it / NS) & 1) ^ 1);
        ptx::mbar_arrive_expect_tx(&kv_full[st], 2 * L::KV_BYTES);
        load_tile(smem + L::OFF_K + st * L::KV_BYTES, &tmK, &kv_full[st], col0, n * p.S + j * BKV, BKV, NB);
        load_tile(smem + L::OFF_V + st * L::KV_BYTES, &tmV, &kv_full[st], col0, n * p.S + j * BKV, BKV, NB);
        if (it == 0) load_o();
      }
      if (it == 0) load_o();
    }
  } else if (warp == 9) {
    if (ptx::elect_one()) {
      constexpr uint32_t idesc_s = ptx::make_idesc_bf16(BQ, BKV, false, false);
      constexpr uint32_t idesc_dq = ptx::make_idesc_bf16(BQ, HD, false, true);
      const uint32_t aQ = ptx::smem_u32(smem + L::OFF_Q);
      const uint32_t adO = ptx::smem_u32(smem + L::OFF_DO);
      int n_tiles = 0;
      for (int j = next_j(-1); j < num_kv; j = next_j(j)) ++n_tiles;
      // S = Q K^T and dP = dO V^T of pipeline slot `it` (K/V stage it % 3) into TMEM buffer it & 1
      auto issue_sdp = [&](int it) {
        const int st = it % NS, b = it & 1;
        const uint32_t aK = ptx::smem_u32(smem + L::OFF_K + st * L::KV_BYTES);
        const uint32_t aV = ptx::smem_u32(smem + L::OFF_V + st * L::KV_BYTES);
        const uint32_t t_s = tmem + b * 128, t_dp = t_s + 64;
        ptx::mbar_wait(&kv_full[st], (it / NS) & 1);
        ptx::tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < HD / 16; ++kk) {
          const int blk = kk >> 2, off = (kk & 3) * 32;
          ptx::mma_bf16_ss(t_s, ptx::make_smem_desc(aQ + blk * (BQ * 128) + off, 16, 1024),
                           ptx::make_smem_desc(aK + blk * (BKV * 128) + off, 16, 1024), idesc_s, kk > 0);
        }
#pragma unroll
        for (int kk = 0; kk < HD / 16; ++kk) {
          const int blk = kk >> 2, off = (kk & 3) * 32;
          ptx::mma_bf16_ss(t_dp, ptx::make_smem_desc(adO + blk * (BQ * 128) + off, 16, 1024),
                           ptx::make_smem_desc(aV + blk * (BKV * 128) + off, 16, 1024), idesc_s, kk > 0);
        }
        ptx::mma_commit(&sdp_full[b]);
      };
      ptx::mbar_wait(q_full, 0);
      if (n_tiles > 0) issue_sdp(0);
      for (int it = 0; it < n_tiles; ++it) {
        // TMEM buffer (it+1)&1 holds dS of slot it-1: dQ += dS K of that slot was issued one iteration ago and the tensor
        // pipe runs in issue order, so S, dP of slot it+1 overwrite the buffer only after it has been consumed
        if (it + 1 < n_tiles) issue_sdp(it + 1);
        const int st = it % NS, b = it & 1;
        const uint32_t aK = ptx::smem_u32(smem + L::OFF_K + st * L::KV_BYTES);
        const uint32_t t_ds = tmem + b * 128;  // keys 0..31 as bf16 pairs in columns [0, 16), keys 32..63 in [32, 48)
        ptx::mbar_wait(&ds_full[b], (it >> 1) & 1);
        ptx::tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < BKV / 16; ++kk) {
          ptx::mma_bf16_ts(t_dq, t_ds + (kk >> 1) * 32 + (kk & 1) * 8, ptx::make_smem_desc(aK + kk * 2048, BKV * 128, 1024),
                           idesc_dq, (it > 0 || kk > 0));
        }
        ptx::mma_commit(&kv_empty[st]);  // the K/V stage may be refilled once these MMAs are done
        if (it == n_tiles - 1) ptx::mma_commit(dq_done);
      }
    }
  } else {
    // two threads per query row: warp & 3 selects the 32 TMEM lanes, warp >> 2 the 32-column half of the score tile
    const int half = warp >> 2;
    const int tid = (warp & 3) * 32 + (threadIdx.x & 31);
    const int i = q0 + tid;
    const bool row_valid = i < p.S;
    const uint32_t lane_off = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const float lse2 = row_valid ? p.lse[static_cast<int64_t>(row0 + tid) * p.H + h] * LOG2E : INFINITY;
    // delta[row] = sum_c dO[row, c] * O[row, c], from the tiles in shared memory: this thread takes the 64 columns of its half
    // (one 128 B row of block `half`; dO and O share the swizzle, so chunk k of one meets chunk k of the other), the two
    // halves meet through shared memory. Written out for the dK/dV kernel, which runs after this one.
    float* dl_s = reinterpret_cast<float*>(smem + L::OFF_DELTA);
    {
      static_assert(NB == 2, "one 64-column block per thread half");
      ptx::mbar_wait(q_full, 0);
      ptx::mbar_wait(o_full, 0);
      const uint4* ro = reinterpret_cast<const uint4*>(smem + L::OFF_O + half * (BQ * 128) + tid * 128);
      const uint4* rd = reinterpret_cast<const uint4*>(smem + L::OFF_DO + half * (BQ * 128) + tid * 128);
      float acc = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const uint4 a = ro[k], b = rd[k];
        const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          acc = fmaf(__uint_as_float(aw[e] << 16), __uint_as_float(bw[e] << 16), acc);
          acc = fmaf(__uint_as_float(aw[e] & 0xffff0000u), __uint_as_float(bw[e] & 0xffff0000u), acc);
        }
      }
      dl_s[half * BQ + tid] = acc;
      ptx::named_bar_sync(1, 256);
    }
    const float delta = row_valid ? dl_s[tid] + dl_s[BQ + tid] : 0.f;
    if (half == 0 && row_valid) delta_out[static_cast<int64_t>(row0 + tid) * p.H + h] = delta;
    const RowVis vis = make_row_vis(p, i, CAUSAL);
    int it = 0;
    int j = next_j(-1);
    uint64_t kw = j < num_kv ? load_kword(p, n, j) : 0ull;
    for (; j < num_kv; ++it) {
      const int jn = next_j(j);
      const uint64_t kw_next = jn < num_kv ? load_kword(p, n, jn) : 0ull;
      const uint64_t bits = key_bits_v(vis, kw, j);
      const bool full = __all_sync(0xffffffffu, (half ? (bits >> 32) : (bits & 0xffffffffull)) == 0xffffffffull);
      const int b = it & 1;
      const uint32_t t_s = tmem + b * 128, t_dp = t_s + 64;
      ptx::mbar_wait(&sdp_full[b], (it >> 1) & 1);  // also: dQ of two slots ago has consumed this buffer's dS (issue order)
      ptx::tc_fence_after();
      {
        uint32_t s[32], dp[32];
        ptx::tmem_ld32(t_s + lane_off + half * 32, s);
        ptx::tmem_ld32(t_dp + lane_off + half * 32, dp);
        ptx::tmem_ld_wait();
        float pv[32], dsv[32];
        recompute_half<false>(s, dp, bits, full, half, p.scale_log2, lse2, delta, pv, dsv);
        uint32_t pk[16];  // this thread's 32 keys as bf16 pairs, over the first half of the S columns it has just read
#pragma unroll
        for (int c = 0; c < 16; ++c) pk[c] = ptx::pack_bf16(dsv[2 * c], dsv[2 * c + 1]);
        ptx::tmem_st16(t_s + lane_off + half * 32, pk);
      }
      ptx::tmem_st_wait();
      ptx::tc_fence_before();
      ptx::mbar_arrive(&ds_full[b]);
      j = jn;
      kw = kw_next;
    }
    ptx::mbar_wait(dq_done, 0);
    ptx::tc_fence_after();
    // this thread's 64 columns of the row as two runs: [half*32, +32) and the rotary partners [64 + half*32, +32)
    static_assert(HD == 128, "rotary pairs are (j, j + 64)");
    __nv_bfloat16* drow = p.dq + static_cast<int64_t>(row0 + tid) * p.ld_dqkv + col0;
    {
      uint32_t ta[32], tb[32];
      ptx::tmem_ld32(t_dq + lane_off + half * 32, ta);
      ptx::tmem_ld32(t_dq + lane_off + 64 + half * 32, tb);
      const float* cs = nullptr;
      if (p.rope_cs) {  // every MMA has completed (dq_done): the K/V stages are free to hold the tile's table rows
        static_assert(L::KV_STAGES * 2 * L::KV_BYTES >= CS_STAGE_BYTES, "staging fits in the K/V ring");
        float* cs_s = reinterpret_cast<float*>(smem + L::OFF_K);
        stage_rope_rows(cs_s, p.rope_cs, HD / 2, warp, [&](int r) {
          const int ri = min(q0 + r, p.S - 1);  // rows beyond the sequence are never stored: any valid table row will do
          return p.rope_pos ? __ldg(p.rope_pos + n * p.S + ri) : ri;
        });
        ptx::named_bar_sync(1, 256);
        cs = cs_s + tid * CS_ROW_FLOATS + half * 64;
      }
      ptx::tmem_ld_wait();
      if (row_valid) store_rotary_pair_runs(drow, half * 32, ta, tb, p.scale, cs);
      __syncwarp();
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem, L::TMEM_COLS);
  }
}

// ---------------------------------------------------------------- dK, dV: one CTA per 128-key tile, loops over 64-row query tiles
// TRANSPOSED scores: S^T = K Q^T and dP^T = V dO^T put the KEYS on the 128 TMEM lanes and the query rows on the columns, so
//   * the resident operand is the big one (K, V: 128 rows) and the streamed one the small one (Q, dO: 64 rows = 32 KB per
//     pipeline slot for 4 GEMM-units; the [64 keys x 128 queries] form streamed 64 KB for the same work),
//   * P^T / dS^T have the keys on the lanes, i.e. they are K-major A operands of dV += P^T dO, dK += dS^T Q: the softmax
//     threads write them as bf16 pairs over the S^T / dP^T columns they have just read (tcgen05.st) and the MMAs take A from
//     TMEM — no shared-memory round trip, no proxy fence; dV, dK accumulate as [128 keys x HD] tiles whose rows are
//     contiguous in global memory (16 B epilogue stores, no transposed 2 B scatter),
//   * four Q/dO stages fit (K,V 64 KB + 4 x 32 KB), so the load of slot it+2 no longer waits for the accumulate MMAs of
//     slot it to release its stage.
// The per-query statistics (lse, delta) vary along the COLUMNS here: they are staged through shared memory once per slot
// and read back as broadcast float4 loads.
constexpr int BKT = 128;  // keys per CTA
constexpr int BQT = 64;   // query rows per pipeline slot

template <int HD>
struct DkvSmem {
  static constexpr int NB = HD / 64;
  static constexpr int KV_BYTES = NB * BKT * 128;
  static constexpr int Q_BYTES = NB * BQT * 128;
  static constexpr int NS = 4;
  static constexpr int OFF_K = 0;
  static constexpr int OFF_V = OFF_K + KV_BYTES;
  static constexpr int OFF_Q = OFF_V + KV_BYTES;       // [NS]
  static constexpr int OFF_DO = OFF_Q + NS * Q_BYTES;  // [NS]
  static constexpr int OFF_STAT = OFF_DO + NS * Q_BYTES;  // float [2][2][BQT]: -lse*log2(e) | delta of the slot's query rows
  static constexpr int OFF_BAR = OFF_STAT + 2 * 2 * BQT * 4;
  static constexpr int TOTAL = OFF_BAR + 128;
  static constexpr int TMEM_COLS = 512;  // S^T,dP^T double-buffered (4 x 64) + dV + dK (2 x HD)
};
static_assert(DkvSmem<128>::TOTAL <= 232448, "dK/dV kernel shared memory");

// p^T = exp2(s*scale_log2 - lse2[q]), ds' = p^T * (dp - delta[q]) for 32 query columns of one key row; nl / dl point at
// the 32 per-column statistics in shared memory (every lane reads the same addresses: broadcast).
__device__ __forceinline__ void recompute_half_t(const uint32_t (&s)[32], const uint32_t (&dp)[32], uint32_t m32, bool full,
                                                 float scale_log2, const float* nl, const float* dl, float (&pv)[32],
                                                 float (&dsv)[32]) {
#pragma unroll
  for (int c4 = 0; c4 < 32; c4 += 4) {
    const float4 a = *reinterpret_cast<const float4*>(nl + c4);
    const float4 d = *reinterpret_cast<const float4*>(dl + c4);
    const float na[4] = {a.x, a.y, a.z, a.w}, da[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int c = c4 + e;
      float pr = ptx::ex2_approx(fmaf(__uint_as_float(s[c]), scale_log2, na[e]));
      if (!full) pr = ((m32 >> c) & 1u) ? pr : 0.f;
      pv[c] = pr;
      dsv[c] = pr * (__uint_as_float(dp[c]) - da[e]);
    }
  }
}

template <int HD, bool CAUSAL>
__global__ void __launch_bounds__(NTHREADS_BWD, 1)
    attn_bwd_dkdv_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmdO,
                         const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                         const AttnParams p) {
  static_assert(HD == 128, "dV / dK accumulate as [128 keys x HD] with HD = 128 TMEM columns each");
  using L = DkvSmem<HD>;
  constexpr int NB = L::NB;
  constexpr int NS = L::NS;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::OFF_BAR);
  uint64_t* kv_full = bars + 0;
  uint64_t* qdo_full = bars + 1;    // [4]
  uint64_t* qdo_empty = bars + 5;   // [4]
  uint64_t* sdp_full = bars + 9;    // [2]
  uint64_t* pds_full = bars + 11;   // [2]
  uint64_t* done = bars + 13;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);
  float* stat = reinterpret_cast<float*>(smem + L::OFF_STAT);

  const int warp = threadIdx.x >> 5;
  const int j = blockIdx.x;
  const int h = blockIdx.y, n = blockIdx.z;
  const int k0 = j * BKT;
  const int col0 = h * HD;
  // query tiles that can see any key of this tile: from the causal diagonal to the end of the sequence, or, when every key
  // lies inside response segments (shared-prefix packing), to the end of the last key's segment
  const int first_qt = CAUSAL ? (k0 / BQT) : 0;
  int hi_max = p.S;
  if (p.segR > 0 && k0 >= p.S0) {
    const int k_last = min(k0 + BKT, p.S) - 1;
    hi_max = min(p.S, p.S0 + ((k_last - p.S0) / p.segR + 1) * p.segR);
  }
  const int last_qt = (hi_max - 1) / BQT;
  const int n_slots = last_qt - first_qt + 1;  // >= 1: the diagonal tile

  if (warp == 8) {
    if (ptx::elect_one()) {
      ptx::prefetch_tmap(&tmQ);
      ptx::prefetch_tmap(&tmdO);
      ptx::prefetch_tmap(&tmK);
      ptx::prefetch_tmap(&tmV);
    }
    __syncwarp();
    ptx::tmem_alloc(tmem_slot, L::TMEM_COLS);
    ptx::tmem_relinquish();
  } else if (warp == 9) {
    if (ptx::elect_one()) {
      if ((ptx::smem_u32(smem) & 1023u) != 0) __trap();
      ptx::mbar_init(kv_full, 1);
      for (int i = 0; i < NS; ++i) {
        ptx::mbar_init(&qdo_full[i], 1);
        ptx::mbar_init(&qdo_empty[i], 1);
      }
      for (int i = 0; i < 2; ++i) {
        ptx::mbar_init(&sdp_full[i], 1);
        ptx::mbar_init(&pds_full[i], 256);
      }
      ptx::mbar_init(done, 1);
      ptx::fence_barrier_init();
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  // TMEM columns: [S^T 0 | dP^T 0 | S^T 1 | dP^T 1] (64 each) | dV (HD) | dK (HD)
  const uint32_t t_dv = tmem + 256, t_dk = tmem + 256 + HD;

  if (warp == 8) {
    if (ptx::elect_one()) {
      ptx::mbar_arrive_expect_tx(kv_full, 2 * L::KV_BYTES);
      load_tile(smem + L::OFF_K, &tmK, kv_full, col0, n * p.S + k0, BKT, NB);
      load_tile(smem + L::OFF_V, &tmV, kv_full, col0, n * p.S + k0, BKT, NB);
      for (int it = 0; it < n_slots; ++it) {
        const int st = it % NS;
        ptx::mbar_wait(&qdo_empty[st], ((it / NS) & 1) ^ 1);
        ptx::mbar_arrive_expect_tx(&qdo_full[st], 2 * L::Q_BYTES);
        const int r0 = n * p.S + (first_qt + it) * BQT;
        load_tile(smem + L::OFF_Q + st * L::Q_BYTES, &tmQ, &qdo_full[st], col0, r0, BQT, NB);
        load_tile(smem + L::OFF_DO + st * L::Q_BYTES, &tmdO, &qdo_full[st], col0, r0, BQT, NB);
      }
    }
  } else if (warp == 9) {
    if (ptx::elect_one()) {
      constexpr uint32_t idesc_s = ptx::make_idesc_bf16(BKT, BQT, false, false);
      constexpr uint32_t idesc_acc = ptx::make_idesc_bf16(BKT, HD, false, true);
      const uint32_t aK = ptx::smem_u32(smem + L::OFF_K);
      const uint32_t aV = ptx::smem_u32(smem + L::OFF_V);
      // S^T = K Q^T and dP^T = V dO^T of pipeline slot `it` (Q/dO stage it % 3) into TMEM buffer it & 1
      auto issue_sdp = [&](int it) {
        const int st = it % NS, b = it & 1;
        const uint32_t aQ = ptx::smem_u32(smem + L::OFF_Q + st * L::Q_BYTES);
        const uint32_t adO = ptx::smem_u32(smem + L::OFF_DO + st * L::Q_BYTES);
        const uint32_t t_s = tmem + b * 128, t_dp = t_s + 64;
        ptx::mbar_wait(&qdo_full[st], (it / NS) & 1);
        ptx::tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < HD / 16; ++kk) {
          const int blk = kk >> 2, off = (kk & 3) * 32;
          ptx::mma_bf16_ss(t_s, ptx::make_smem_desc(aK + blk * (BKT * 128) + off, 16, 1024),
                           ptx::make_smem_desc(aQ + blk * (BQT * 128) + off, 16, 1024), idesc_s, kk > 0);
        }
#pragma unroll
        for (int kk = 0; kk < HD / 16; ++kk) {
          const int blk = kk >> 2, off = (kk & 3) * 32;
          ptx::mma_bf16_ss(t_dp, ptx::make_smem_desc(aV + blk * (BKT * 128) + off, 16, 1024),
                           ptx::make_smem_desc(adO + blk * (BQT * 128) + off, 16, 1024), idesc_s, kk > 0);
        }
        ptx::mma_commit(&sdp_full[b]);
      };
      ptx::mbar_wait(kv_full, 0);
      issue_sdp(0);
      for (int it = 0; it < n_slots; ++it) {
        // TMEM buffer (it+1)&1 holds P^T / dS^T of slot it-1: their accumulate MMAs were issued one iteration ago and the
        // tensor pipe runs in issue order, so S^T, dP^T of slot it+1 overwrite the buffer only after it has been consumed
        if (it + 1 < n_slots) issue_sdp(it + 1);
        const int st = it % NS, b = it & 1;
        const uint32_t aQ = ptx::smem_u32(smem + L::OFF_Q + st * L::Q_BYTES);
        const uint32_t adO = ptx::smem_u32(smem + L::OFF_DO + st * L::Q_BYTES);
        // query rows 0..31 of the slot as bf16 pairs in columns [0, 16) of the S^T (P^T) / dP^T (dS^T) buffer, rows 32..63 in [32, 48)
        const uint32_t t_p = tmem + b * 128, t_ds = t_p + 64;
        ptx::mbar_wait(&pds_full[b], (it >> 1) & 1);
        ptx::tc_fence_after();
        // dV[keys, hd] += P^T dO ; dK[keys, hd] += dS^T Q   (contraction over the slot's 64 query rows; A from TMEM, B MN-major)
#pragma unroll
        for (int kk = 0; kk < BQT / 16; ++kk) {
          ptx::mma_bf16_ts(t_dv, t_p + (kk >> 1) * 32 + (kk & 1) * 8, ptx::make_smem_desc(adO + kk * 2048, BQT * 128, 1024),
                           idesc_acc, (it > 0 || kk > 0));
        }
#pragma unroll
        for (int kk = 0; kk < BQT / 16; ++kk) {
          ptx::mma_bf16_ts(t_dk, t_ds + (kk >> 1) * 32 + (kk & 1) * 8, ptx::make_smem_desc(aQ + kk * 2048, BQT * 128, 1024),
                           idesc_acc, (it > 0 || kk > 0));
        }
        ptx::mma_commit(&qdo_empty[st]);  // the Q/dO stage may be refilled once these MMAs are done
        if (it == n_slots - 1) ptx::mma_commit(done);
      }
    }
  } else {
    // two threads per TMEM lane (= key): warp & 3 selects the 32 lanes, warp >> 2 the 32-column half of the 64 query rows
    const int half = warp >> 2;
    const int tid = (warp & 3) * 32 + (threadIdx.x & 31);  // key row within the tile
    const uint32_t lane_off = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const int k = k0 + tid;
    // queries that may see key k: [lo_k, hi_k) (causal from the diagonal; a response-segment key only inside its segment)
    const bool key_ok = k < p.S && ((load_kword(p, n, k >> 6) >> (k & 63)) & 1ull);
    const int lo_k = CAUSAL ? k : 0;
    int hi_k = p.S;
    if (p.segR > 0 && k >= p.S0) hi_k = min(p.S, p.S0 + ((k - p.S0) / p.segR + 1) * p.segR);
    if (!key_ok) hi_k = 0;
    // per-query statistics of the NEXT slot are fetched while this one is being worked on (threads 0..127 of the group:
    // 64 x -lse*log2(e), 64 x delta)
    const int sidx = threadIdx.x & 63;
    const bool loads_stat = threadIdx.x < 2 * BQT;
    auto load_stat = [&](int it) -> float {
      const int i = (first_qt + it) * BQT + sidx;
      if (!loads_stat || it >= n_slots || i >= p.S) return threadIdx.x < BQT ? -INFINITY : 0.f;
      const int64_t g = (static_cast<int64_t>(n) * p.S + i) * p.H + h;
      return threadIdx.x < BQT ? -__ldg(p.lse + g) * LOG2E : __ldg(p.delta + g);
    };
    float stat_next = load_stat(0);
    for (int it = 0; it < n_slots; ++it) {
      const int b = it & 1;
      float* st_buf = stat + b * 2 * BQT;
      if (loads_stat) st_buf[threadIdx.x] = stat_next;
      stat_next = load_stat(it + 1);
      const int qbase = (first_qt + it) * BQT + half * 32;
      const int lo = min(max(lo_k - qbase, 0), 32), hi = min(max(hi_k - qbase, 0), 32);
      const uint32_t m32 = hi > lo ? ((hi >= 32 ? ~0u : ((1u << hi) - 1u)) & ~((1u << lo) - 1u)) : 0u;
      const bool full = __all_sync(0xffffffffu, m32 == ~0u);
      const uint32_t t_s = tmem + b * 128, t_dp = t_s + 64;
      ptx::named_bar_sync(1, 256);  // this slot's statistics are in shared memory (two buffers: the previous slot's are still being read)
      ptx::mbar_wait(&sdp_full[b], (it >> 1) & 1);  // also: the MMAs that read this buffer's P^T / dS^T two slots ago are done
      ptx::tc_fence_after();
      {
        uint32_t s[32], dp[32];
        ptx::tmem_ld32(t_s + lane_off + half * 32, s);
        ptx::tmem_ld32(t_dp + lane_off + half * 32, dp);
        ptx::tmem_ld_wait();
        float pv[32], dsv[32];
        recompute_half_t(s, dp, m32, full, p.scale_log2, st_buf + half * 32, st_buf + BQT + half * 32, pv, dsv);
        uint32_t pk[16], dk[16];  // 32 query columns as bf16 pairs, over the first half of the columns this thread has just read
#pragma unroll
        for (int c = 0; c < 16; ++c) {
          pk[c] = ptx::pack_bf16(pv[2 * c], pv[2 * c + 1]);
          dk[c] = ptx::pack_bf16(dsv[2 * c], dsv[2 * c + 1]);
        }
        ptx::tmem_st16(t_s + lane_off + half * 32, pk);
        ptx::tmem_st16(t_dp + lane_off + half * 32, dk);
      }
      ptx::tmem_st_wait();
      ptx::tc_fence_before();
      ptx::mbar_arrive(&pds_full[b]);
    }
    ptx::mbar_wait(done, 0);
    ptx::tc_fence_after();
    // thread == key row. dV: this thread's half of the head dimension, 64 contiguous bf16. dK: two runs of 32 columns,
    // [half*32, +32) and the rotary partners [64 + half*32, +32), so the inverse rotation needs no second thread.
    const bool row_ok = k < p.S;
    __nv_bfloat16* vrow = p.dv + (static_cast<int64_t>(n) * p.S + k) * p.ld_dqkv + col0 + half * 64;
    __nv_bfloat16* krow = p.dk + (static_cast<int64_t>(n) * p.S + k) * p.ld_dqkv + col0;
#pragma unroll 1
    for (int c = 0; c < 2; ++c) {
      uint32_t tv[32];
      ptx::tmem_ld32(t_dv + lane_off + half * 64 + c * 32, tv);
      ptx::tmem_ld_wait();
      if (row_ok) {
        uint4* dv4 = reinterpret_cast<uint4*>(vrow + c * 32);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint4 u;
          u.x = ptx::pack_bf16(__uint_as_float(tv[q * 8 + 0]), __uint_as_float(tv[q * 8 + 1]));
          u.y = ptx::pack_bf16(__uint_as_float(tv[q * 8 + 2]), __uint_as_float(tv[q * 8 + 3]));
          u.z = ptx::pack_bf16(__uint_as_float(tv[q * 8 + 4]), __uint_as_float(tv[q * 8 + 5]));
          u.w = ptx::pack_bf16(__uint_as_float(tv[q * 8 + 6]), __uint_as_float(tv[q * 8 + 7]));
          dv4[q] = u;
        }
      }
      __syncwarp();
    }
    {
      uint32_t ta[32], tb[32];
      ptx::tmem_ld32(t_dk + lane_off + half * 32, ta);
      ptx::tmem_ld32(t_dk + lane_off + 64 + half * 32, tb);
      const float* cs = nullptr;
      if (p.rope_cs) {  // every MMA has completed (done): the Q/dO stages are free to hold the tile's table rows
        static_assert(2 * L::NS * L::Q_BYTES >= CS_STAGE_BYTES, "staging fits in the Q/dO ring");
        float* cs_s = reinterpret_cast<float*>(smem + L::OFF_Q);
        stage_rope_rows(cs_s, p.rope_cs, HD / 2, warp, [&](int r) {
          const int ki = min(k0 + r, p.S - 1);
          return p.rope_pos ? __ldg(p.rope_pos + n * p.S + ki) : ki;
        });
        ptx::named_bar_sync(1, 256);
        cs = cs_s + tid * CS_ROW_FLOATS + half * 64;
      }
      ptx::tmem_ld_wait();
      if (row_ok) store_rotary_pair_runs(krow, half * 32, ta, tb, p.scale, cs);
      __syncwarp();
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem, L::TMEM_COLS);
  }
}

__global__ void pack_mask_kernel(const uint8_t* __restrict__ mask, uint64_t* __restrict__ bits, int N, int S, int nwords) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= N * nwords) return;
  const int n = idx / nwords, w = idx % nwords;
  uint64_t b = 0;
  for (int c = 0; c < 64; ++c) {
    const int k = w * 64 + c;
    if (k < S && mask[static_cast<int64_t>(n) * S + k]) b |= (1ull << c);
  }
  bits[idx] = b;
}

thread_local char g_err[512] = {0};

int make_maps(const AttnArgs& a, const __nv_bfloat16* base, int64_t ld, int box_rows, CUtensorMap* m) {
  const uint64_t T = static_cast<uint64_t>(a.N) * a.S;
  int rc = make_tmap_bf16_2d(m, base, T, static_cast<uint64_t>(a.H) * a.HD, ld, box_rows, 64);
  if (rc != 0) snprintf(g_err, sizeof(g_err), "attention tensor map: %s", tmap_last_error());
  return rc;
}

AttnParams make_params(const AttnArgs& a) {
  AttnParams p{};
  p.N = a.N;
  p.S = a.S;
  p.H = a.H;
  p.nwords = (a.S + 63) / 64;
  p.S0 = a.prefix_len;
  p.segR = a.segment_len;
  p.scale = a.scale;
  p.scale_log2 = a.scale * LOG2E;
  p.kbits = a.kbits;
  p.o = a.o;
  p.ld_o = a.ld_o;
  p.lse = a.lse;
  p.delta = a.delta;
  p.dq = a.dq;
  p.dk = a.dk;
  p.dv = a.dv;
  p.ld_dqkv = a.ld_dqkv;
  p.rope_cs = a.rope_cs;
  p.rope_pos = a.rope_pos;
  return p;
}

template <typename K>
int set_smem(K kern, int bytes) {
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e != cudaSuccess) {
    snprintf(g_err, sizeof(g_err), "cudaFuncSetAttribute(%d): %s", bytes, cudaGetErrorString(e));
    return -10;
  }
  return 0;
}

template <int HD, bool CAUSAL>
int launch_fwd(const AttnArgs& a, cudaStream_t stream) {
  CUtensorMap mq, mk, mv;
  if (make_maps(a, a.q, a.ld_qkv, BQ, &mq) || make_maps(a, a.k, a.ld_qkv, BKV, &mk) || make_maps(a, a.v, a.ld_qkv, BKV, &mv))
    return -6;
  // OPADPO_ATTN_FWD=ss selects the first-generation kernel (P staged through shared memory); default: P in tensor memory
  static const bool use_ts = []() {
    const char* e = getenv("OPADPO_ATTN_FWD");
    return !(e && e[0] == 's' && e[1] == 's');
  }();
  dim3 grid((a.S + BQ - 1) / BQ, a.H, a.N);
  if (use_ts) {
    auto kern = attn_fwd_ts_kernel<HD, CAUSAL>;
    static bool once_ts = false;
    if (!once_ts) {
      if (int rc = set_smem(kern, FwdTsSmem<HD>::TOTAL)) return rc;
      once_ts = true;
    }
    kern<<<grid, NTHREADS, FwdTsSmem<HD>::TOTAL, stream>>>(mq, mk, mv, make_params(a));
  } else {
    auto kern = attn_fwd_kernel<HD, CAUSAL>;
    static bool once = false;
    if (!once) {
      if (int rc = set_smem(kern, FwdSmem<HD>::TOTAL)) return rc;
      once = true;
    }
    kern<<<grid, NTHREADS, FwdSmem<HD>::TOTAL, stream>>>(mq, mk, mv, make_params(a));
  }
  count_launch();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    snprintf(g_err, sizeof(g_err), "attn fwd launch: %s", cudaGetErrorString(e));
    return -11;
  }
  return 0;
}

template <bool CAUSAL>
int launch_bwd(const AttnArgs& a, cudaStream_t stream) {
  constexpr int HD = 128;
  CUtensorMap mq, mdo, mk, mv;
  if (make_maps(a, a.q, a.ld_qkv, BQ, &mq) || make_maps(a, a.dO, a.ld_do, BQ, &mdo) ||
      make_maps(a, a.k, a.ld_qkv, BKV, &mk) || make_maps(a, a.v, a.ld_qkv, BKV, &mv))
    return -6;
  // the dK/dV kernel keeps 128-key tiles resident and streams 64-row query tiles
  CUtensorMap mq64, mdo64, mk128, mv128;
  if (make_maps(a, a.q, a.ld_qkv, BQT, &mq64) || make_maps(a, a.dO, a.ld_do, BQT, &mdo64) ||
      make_maps(a, a.k, a.ld_qkv, BKT, &mk128) || make_maps(a, a.v, a.ld_qkv, BKT, &mv128))
    return -6;
  CUtensorMap mo;  // the dQ kernel also computes delta = rowsum(dO * O) from the O tile (and writes it for the dK/dV kernel)
  if (make_maps(a, a.o, a.ld_o, BQ, &mo)) return -6;
  AttnParams p = make_params(a);
  auto kdq = attn_bwd_dq_kernel<HD, CAUSAL>;
  auto kdkv = attn_bwd_dkdv_kernel<HD, CAUSAL>;
  static bool once = false;
  if (!once) {
    if (int rc = set_smem(kdq, DqSmem<HD>::TOTAL)) return rc;
    if (int rc = set_smem(kdkv, DkvSmem<HD>::TOTAL)) return rc;
    once = true;
  }
  dim3 gq((a.S + BQ - 1) / BQ, a.H, a.N);
  kdq<<<gq, NTHREADS_BWD, DqSmem<HD>::TOTAL, stream>>>(mq, mdo, mo, mk, mv, p, a.delta);
  dim3 gk((a.S + BKT - 1) / BKT, a.H, a.N);
  kdkv<<<gk, NTHREADS_BWD, DkvSmem<HD>::TOTAL, stream>>>(mq64, mdo64, mk128, mv128, p);
  count_launch(2);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    snprintf(g_err, sizeof(g_err), "attn bwd launch: %s", cudaGetErrorString(e));
    return -11;
  }
  return 0;
}

}  // namespace

const char* attn_last_error() { return g_err; }

int attn_forward(const AttnArgs& a, cudaStream_t stream) {
  if (a.HD == 128) return a.causal ? launch_fwd<128, true>(a, stream) : launch_fwd<128, false>(a, stream);
  if (a.HD == 64) return a.causal ? launch_fwd<64, true>(a, stream) : launch_fwd<64, false>(a, stream);
  snprintf(g_err, sizeof(g_err), "attention: head_dim %d not supported (64 or 128)", a.HD);
  return -1;
}

int attn_backward(const AttnArgs& a, cudaStream_t stream) {
  if (a.HD != 128) {
    snprintf(g_err, sizeof(g_err), "attention backward: head_dim %d not supported (128)", a.HD);
    return -1;
  }
  return a.causal ? launch_bwd<true>(a, stream) : launch_bwd<false>(a, stream);
}

int attn_pack_mask(const uint8_t* mask, uint64_t* bits, int N, int S, cudaStream_t stream) {
  const int nwords = (S + 63) / 64;
  const int total = N * nwords;
  pack_mask_kernel<<<(total + 127) / 128, 128, 0, stream>>>(mask, bits, N, S, nwords);
  count_launch();
  return cudaGetLastError() == cudaSuccess ? 0 : -11;
}

}  // namespace opadpo
