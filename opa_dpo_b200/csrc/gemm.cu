// Persistent, warp-specialised tcgen05 GEMM (sm_100a). See gemm.h for the math and DESIGN.md for the layout.
//
//   warp 0 : TMA producer (one elected lane)          global -> 128B-swizzled smem ring (4 stages x 48 KB)
//   warp 1 : MMA issuer   (one elected lane)          tcgen05.mma 128x256x16, fp32 accumulators in TMEM
//   warp 2 : TMEM allocator (512 columns = 2 accumulator buffers, so epilogue(i) overlaps mainloop(i+1))
//   warps 4-7 : epilogue, one thread per accumulator row (TMEM lane), tcgen05.ld 32 columns at a time
#include "gemm.h"

#include <cstdio>
#include <cstdlib>
#include <vector>

#include "common.h"
#include "ptx.cuh"
#include "tmap.h"

namespace opadpo {
namespace {

constexpr int BM = 128, BN = 256, BK = 64, STAGES = 4;
constexpr int A_BYTES = BM * BK * 2;   // 16 KB
constexpr int B_BYTES = BN * BK * 2;   // 32 KB
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int NUM_THREADS = 256;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;
constexpr float LOG2E = 1.4426950408889634f;

struct KParams {
  int M, N;
  int kb1, kb2;
  int a2_group_n, a2_group_stride;
  int splits, kb_per_split;
  int num_m, num_n, num_tiles;
  int group_m;  // m blocks that share each streamed B panel while their A tiles stay L2-resident
  float alpha;
  void* C;
  int64_t ldc;
  const __nv_bfloat16* bias;
  const __nv_bfloat16* residual;
  int64_t ldr;
  int act;
  const int32_t* labels;
  float4* lse_partials;
  float* label_logit;
  const float* lse;
  const float* dlp;
  const float* dent;  // EPI_CEGRAD: d(loss)/d(entropy) per row, or null
  const float* ent;   // EPI_CEGRAD: the row's entropy (needed with dent)
  const int32_t* rope_pos;  // EPI_BF16_ROPE
  const float2* rope_cs;
  int rope_cols, rope_S;
};

struct Tile {
  int m_blk, n_blk, kb_begin, kb_end, split;
};

// CL = 1: `tile` enumerates (m block, n block, split). CL = 2: `tile` enumerates (PAIR of m blocks, n block); the two
// CTAs of a cluster take m blocks 2*mp + rank of the same n block, so they can share (multicast) the B tile.
template <int CL>
__device__ __forceinline__ Tile decode_tile(const KParams& p, int tile, int rank) {
  constexpr int PS = CL == 1 ? 1 : 2;  // CTAs per cluster
  const int num_m = CL == 1 ? p.num_m : (p.num_m + 1) / 2;
  const int tiles_mn = num_m * p.num_n;
  const int split = tile / tiles_mn;
  const int t = tile - split * tiles_mn;
  const int GM = max(1, p.group_m / PS);
  const int per_group = GM * p.num_n;
  const int group = t / per_group;
  const int first_m = group * GM;
  const int gsize = min(GM, num_m - first_m);
  const int r = t - group * per_group;
  Tile o;
  o.split = split;
  o.m_blk = (first_m + r % gsize) * PS + rank;
  o.n_blk = r / gsize;
  if (p.splits == 1) {
    o.kb_begin = 0;
    o.kb_end = p.kb1 + p.kb2;
  } else {  // split-K over the concatenated k-blocks of both segments
    o.kb_begin = split * p.kb_per_split;
    o.kb_end = min(p.kb1 + p.kb2, o.kb_begin + p.kb_per_split);
  }
  return o;
}

__device__ __forceinline__ float act_fn(float x, int act) {
  if (act == ACT_GELU) return 0.5f * x * (1.f + erff(x * 0.70710678118654752f));
  if (act == ACT_QUICK_GELU) return x / (1.f + __expf(-1.702f * x));
  return x;
}

__device__ __forceinline__ void red_add_v4(float* dst, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

template <bool A_MN, bool B_MN, int EPI, int CL>
__global__ void __launch_bounds__(NUM_THREADS, 1)
    gemm_kernel(const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmB1,
                const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmB2, const KParams p) {
  // CL == 2: launched as clusters of two CTAs that work on vertically adjacent output tiles; each loads HALF of the
  // shared B tile and multicasts it into both CTAs' smem, so per-SM L2->smem traffic per k-block drops from 48 to 32 KB.
  // CL == 3: CTA PAIR (tcgen05 cta_group::2): ONE 256x256x16 MMA spans both SMs; each CTA keeps its 128 A rows and only
  // ITS half of the B tile in smem (32 KB per stage, 6 stages), the leader CTA issues every MMA, accumulators live in
  // both CTAs' TMEM, and smem operand reads per SM drop by a third.
  constexpr bool PAIR = CL == 3;
  constexpr int PS = CL == 1 ? 1 : 2;
  constexpr int NSTAGE = PAIR ? 6 : STAGES;
  constexpr int BB = PAIR ? B_BYTES / 2 : B_BYTES;  // B bytes per stage held by this CTA
  constexpr int SBYTES = A_BYTES + BB;
  static_assert(NSTAGE * SBYTES <= STAGES * STAGE_BYTES, "smem budget");
  const int rank = CL == 1 ? 0 : static_cast<int>(ptx::cluster_ctarank());
  const int tile_first = CL == 1 ? blockIdx.x : blockIdx.x / PS;
  const int tile_stride = CL == 1 ? gridDim.x : gridDim.x / PS;
  const int num_tiles = CL == 1 ? p.num_tiles : ((p.num_m + 1) / 2) * p.num_n * p.splits;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + 8;
  uint64_t* tfull_bar = empty_bar + 8;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0) {
    if (ptx::elect_one()) {
      ptx::prefetch_tmap(&tmA1);
      ptx::prefetch_tmap(&tmB1);
      ptx::prefetch_tmap(&tmA2);
      ptx::prefetch_tmap(&tmB2);
    }
  } else if (warp == 1) {
    if (ptx::elect_one()) {
      for (int i = 0; i < NSTAGE; ++i) {
        ptx::mbar_init(&full_bar[i], 1);
        ptx::mbar_init(&empty_bar[i], CL == 2 ? 2 : 1);  // CL == 2: the peer's MMA must also be done with the slot
      }
      for (int i = 0; i < 2; ++i) {
        ptx::mbar_init(&tfull_bar[i], 1);
        ptx::mbar_init(&tempty_bar[i], PAIR ? 8 : 4);  // PAIR: the leader waits for both CTAs' epilogue warps
      }
      ptx::fence_barrier_init();
    }
  } else if (warp == 2) {
    if (PAIR) {  // collective over the CTA pair: one warp of each CTA
      ptx::tmem_alloc2(tmem_slot, 512);
      ptx::tmem_relinquish2();
    } else {
      ptx::tmem_alloc(tmem_slot, 512);
      ptx::tmem_relinquish();
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (CL > 1) ptx::cluster_sync();  // the peer's barriers must be initialised before anything is multicast into them
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (ptx::elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = tile_first; tile < num_tiles; tile += tile_stride) {
        const Tile t = decode_tile<CL>(p, tile, rank);
        const int m0 = t.m_blk * BM, n0 = t.n_blk * BN;
        const int a2off = p.a2_group_n > 0 ? (n0 / p.a2_group_n) * p.a2_group_stride : 0;
        for (int kb = t.kb_begin; kb < t.kb_end; ++kb) {
          ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
          const bool seg2 = kb >= p.kb1;
          const CUtensorMap* ma = seg2 ? &tmA2 : &tmA1;
          const CUtensorMap* mb = seg2 ? &tmB2 : &tmB1;
          const int k0 = (seg2 ? kb - p.kb1 : kb) * BK;
          uint8_t* sA = smem + stage * SBYTES;
          uint8_t* sB = sA + A_BYTES;
          if constexpr (PAIR) {
            // both CTAs' tiles complete on the LEADER's full barrier (it alone waits on it and issues the MMAs)
            if (rank == 0) ptx::mbar_arrive_expect_tx(&full_bar[stage], 2 * SBYTES);
            const uint32_t fb = ptx::mapa(&full_bar[stage], 0);
            ptx::tma_load_2d_pair(sA, ma, fb, k0 + (seg2 ? a2off : 0), m0);
            if (!B_MN) {
              ptx::tma_load_2d_pair(sB, mb, fb, k0, n0 + rank * (BN / 2));
            } else {
#pragma unroll
              for (int g = 0; g < BN / 128; ++g)
                ptx::tma_load_2d_pair(sB + g * 8192, mb, fb, n0 + (rank * (BN / 128) + g) * 64, k0);
            }
            if (++stage == NSTAGE) {
              stage = 0;
              phase ^= 1;
            }
            continue;
          }
          ptx::mbar_arrive_expect_tx(&full_bar[stage], STAGE_BYTES);
          if (!A_MN) {
            ptx::tma_load_2d(sA, ma, &full_bar[stage], k0 + (seg2 ? a2off : 0), m0);
          } else {
#pragma unroll
            for (int g = 0; g < BM / 64; ++g) ptx::tma_load_2d(sA + g * 8192, ma, &full_bar[stage], m0 + g * 64, k0);
          }
          if (CL == 1) {
            if (!B_MN) {
              ptx::tma_load_2d(sB, mb, &full_bar[stage], k0, n0);
            } else {
#pragma unroll
              for (int g = 0; g < BN / 64; ++g) ptx::tma_load_2d(sB + g * 8192, mb, &full_bar[stage], n0 + g * 64, k0);
            }
          } else {  // my half of the B tile, multicast into both CTAs (the tensor map's box is BN/2 rows when K-major)
            if (!B_MN) {
              ptx::tma_load_2d_mc(sB + rank * (B_BYTES / 2), mb, &full_bar[stage], k0, n0 + rank * (BN / 2), 0x3);
            } else {
#pragma unroll
              for (int g = 0; g < BN / 128; ++g) {
                const int gg = rank * (BN / 128) + g;
                ptx::tma_load_2d_mc(sB + gg * 8192, mb, &full_bar[stage], n0 + gg * 64, k0, 0x3);
              }
            }
          }
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1 && (!PAIR || rank == 0)) {
    // ------------------------------------------------------------ MMA issuer (PAIR: the leader CTA only)
    if (ptx::elect_one()) {
      constexpr uint32_t idesc = ptx::make_idesc_bf16(PAIR ? 2 * BM : BM, BN, A_MN, B_MN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = tile_first; tile < num_tiles; tile += tile_stride) {
        const Tile t = decode_tile<CL>(p, tile, rank);
        ptx::mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        uint32_t accumulate = 0;
        for (int kb = t.kb_begin; kb < t.kb_end; ++kb) {
          ptx::mbar_wait(&full_bar[stage], phase);
          ptx::tc_fence_after();
          const uint32_t a_addr = ptx::smem_u32(smem + stage * SBYTES);
          const uint32_t b_addr = a_addr + A_BYTES;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t adesc = A_MN ? ptx::make_smem_desc(a_addr + k * 2048, 8192, 1024)
                                        : ptx::make_smem_desc(a_addr + k * 32, 16, 1024);
            const uint64_t bdesc = B_MN ? ptx::make_smem_desc(b_addr + k * 2048, 8192, 1024)
                                        : ptx::make_smem_desc(b_addr + k * 32, 16, 1024);
            if (PAIR)
              ptx::mma_bf16_ss2(d_tmem, adesc, bdesc, idesc, accumulate);
            else
              ptx::mma_bf16_ss(d_tmem, adesc, bdesc, idesc, accumulate);
            accumulate = 1;
          }
          if (CL == 1)
            ptx::mma_commit(&empty_bar[stage]);
          else if (CL == 2)
            ptx::mma_commit_mc(&empty_bar[stage], 0x3);  // frees the slot in both CTAs' rings
          else
            ptx::mma_commit2_mc(&empty_bar[stage], 0x3);
          if (++stage == NSTAGE) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (PAIR)
          ptx::mma_commit2_mc(&tfull_bar[acc], 0x3);  // both CTAs' epilogue warps drain their half of the accumulator
        else
          ptx::mma_commit(&tfull_bar[acc]);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------ epilogue (thread == accumulator row)
    const int quad = warp - 4;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = tile_first; tile < num_tiles; tile += tile_stride) {
      const Tile t = decode_tile<CL>(p, tile, rank);
      const int m0 = t.m_blk * BM, n0 = t.n_blk * BN;
      ptx::mbar_wait(&tfull_bar[acc], acc_phase);
      ptx::tc_fence_after();
      const int row = m0 + quad * 32 + lane;
      const bool row_ok = row < p.M;
      const uint32_t taddr = tmem_base + acc * BN + (static_cast<uint32_t>(quad * 32) << 16);

      if constexpr (EPI == EPI_BF16) {
        __nv_bfloat16* crow = reinterpret_cast<__nv_bfloat16*>(p.C) + static_cast<int64_t>(row) * p.ldc;
        const __nv_bfloat16* rrow = p.residual ? p.residual + static_cast<int64_t>(row) * p.ldr : nullptr;
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
          uint32_t r[32];
          ptx::tmem_ld32(taddr + c * 32, r);
          ptx::tmem_ld_wait();
          const int col0 = n0 + c * 32;
          if (row_ok && col0 < p.N) {
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) * p.alpha;
          if (col0 + 32 <= p.N) {
            if (p.bias) {
              const uint4* b4 = reinterpret_cast<const uint4*>(p.bias + col0);
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                uint4 u = __ldg(b4 + q);
                const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  float2 f = __bfloat1622float2(h[e]);
                  v[q * 8 + e * 2] += f.x;
                  v[q * 8 + e * 2 + 1] += f.y;
                }
              }
            }
            if (p.act != ACT_NONE) {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = act_fn(v[j], p.act);
            }
            if (rrow) {
              const uint4* r4 = reinterpret_cast<const uint4*>(rrow + col0);
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                uint4 u = __ldg(r4 + q);
                const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  float2 f = __bfloat1622float2(h[e]);
                  v[q * 8 + e * 2] += f.x;
                  v[q * 8 + e * 2 + 1] += f.y;
                }
              }
            }
            uint4* dst = reinterpret_cast<uint4*>(crow + col0);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              uint4 u;
              u.x = ptx::pack_bf16(v[q * 8 + 0], v[q * 8 + 1]);
              u.y = ptx::pack_bf16(v[q * 8 + 2], v[q * 8 + 3]);
              u.z = ptx::pack_bf16(v[q * 8 + 4], v[q * 8 + 5]);
              u.w = ptx::pack_bf16(v[q * 8 + 6], v[q * 8 + 7]);
              dst[q] = u;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int col = col0 + j;
              if (col < p.N) {
                float x = v[j];
                if (p.bias) x += __bfloat162float(p.bias[col]);
                x = act_fn(x, p.act);
                if (rrow) x += __bfloat162float(rrow[col]);
                crow[col] = __float2bfloat16(x);
              }
            }
          }
          }
          __syncwarp();
        }
      } else if constexpr (EPI == EPI_SWIGLU_BF16) {
        // gate|up projection with SwiGLU in the epilogue: the weight rows are interleaved so that a 256-column tile holds
        // 128 gate columns followed by their 128 up columns; the result has N/2 columns. The roundings of the unfused path
        // (gate and up stored as bf16, silu rounded to bf16 before the product: swiglu_fwd_kernel) are reproduced, so the
        // output is bit-identical to projection + swiglu_fwd on the non-interleaved weight.
        __nv_bfloat16* crow = reinterpret_cast<__nv_bfloat16*>(p.C) + static_cast<int64_t>(row) * p.ldc + (n0 >> 1);
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          uint32_t gt[32], up[32];
          ptx::tmem_ld32(taddr + c * 32, gt);
          ptx::tmem_ld32(taddr + 128 + c * 32, up);
          ptx::tmem_ld_wait();
          if (row_ok && n0 < p.N) {
            float v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float gg = __bfloat162float(__float2bfloat16(__uint_as_float(gt[j]) * p.alpha));
              const float uu = __bfloat162float(__float2bfloat16(__uint_as_float(up[j]) * p.alpha));
              v[j] = __bfloat162float(__float2bfloat16(gg / (1.f + __expf(-gg)))) * uu;
            }
            uint4* dst = reinterpret_cast<uint4*>(crow + c * 32);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              uint4 u;
              u.x = ptx::pack_bf16(v[q * 8 + 0], v[q * 8 + 1]);
              u.y = ptx::pack_bf16(v[q * 8 + 2], v[q * 8 + 3]);
              u.z = ptx::pack_bf16(v[q * 8 + 4], v[q * 8 + 5]);
              u.w = ptx::pack_bf16(v[q * 8 + 6], v[q * 8 + 7]);
              dst[q] = u;
            }
          }
          __syncwarp();
        }
      } else if constexpr (EPI == EPI_SWIGLU_BWD) {
        // d(act) never goes to memory: the tile is turned into d(gate), d(up) against the saved gate|up activations right here
        // (swiglu_bwd_kernel's arithmetic on d(act) rounded to bf16 as the unfused path stores it: bit-identical results). The
        // GEMM is tensor-bound, so the 4 F bytes per token this epilogue moves ride along for free.
        const __nv_bfloat16* grow = p.residual + static_cast<int64_t>(row) * p.ldr;
        __nv_bfloat16* orow = reinterpret_cast<__nv_bfloat16*>(p.C) + static_cast<int64_t>(row) * p.ldc;
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
          uint32_t r[32];
          ptx::tmem_ld32(taddr + c * 32, r);
          ptx::tmem_ld_wait();
          const int col0 = n0 + c * 32;
          if (row_ok && col0 + 32 <= p.N) {
            const uint4* g4 = reinterpret_cast<const uint4*>(grow + col0);
            const uint4* u4 = reinterpret_cast<const uint4*>(grow + p.N + col0);
            uint4* dg4 = reinterpret_cast<uint4*>(orow + col0);
            uint4* du4 = reinterpret_cast<uint4*>(orow + p.N + col0);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const uint4 gv = __ldg(g4 + q), uv = __ldg(u4 + q);
              const __nv_bfloat162* gh = reinterpret_cast<const __nv_bfloat162*>(&gv);
              const __nv_bfloat162* uh = reinterpret_cast<const __nv_bfloat162*>(&uv);
              float dg[8], du[8];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float2 gf = __bfloat1622float2(gh[e]), uf = __bfloat1622float2(uh[e]);
                const float gg[2] = {gf.x, gf.y}, uu[2] = {uf.x, uf.y};
#pragma unroll
                for (int t2 = 0; t2 < 2; ++t2) {
                  const float da = __bfloat162float(__float2bfloat16(__uint_as_float(r[q * 8 + e * 2 + t2]) * p.alpha));
                  const float sg = 1.f / (1.f + __expf(-gg[t2]));
                  const float silu = gg[t2] * sg;
                  du[e * 2 + t2] = da * silu;
                  dg[e * 2 + t2] = da * uu[t2] * sg * (1.f + gg[t2] * (1.f - sg));
                }
              }
              uint4 a, b;
              a.x = ptx::pack_bf16(dg[0], dg[1]);
              a.y = ptx::pack_bf16(dg[2], dg[3]);
              a.z = ptx::pack_bf16(dg[4], dg[5]);
              a.w = ptx::pack_bf16(dg[6], dg[7]);
              b.x = ptx::pack_bf16(du[0], du[1]);
              b.y = ptx::pack_bf16(du[2], du[3]);
              b.z = ptx::pack_bf16(du[4], du[5]);
              b.w = ptx::pack_bf16(du[6], du[7]);
              dg4[q] = a;
              du4[q] = b;
            }
          }
          __syncwarp();
        }
      } else if constexpr (EPI == EPI_BF16_ROPE) {
        // q|k|v projection with the rotary embedding applied to the fp32 accumulator (one bf16 rounding instead of two and
        // no separate pass over q and k): heads are 128 columns, HF's rotate_half pairs column j with j + 64, the table
        // holds (cos, sin) for j < 64 of every position; columns from rope_cols on (v) are stored as they are.
        __nv_bfloat16* crow = reinterpret_cast<__nv_bfloat16*>(p.C) + static_cast<int64_t>(row) * p.ldc;
        const int pos = row_ok ? (p.rope_pos ? p.rope_pos[row] : row % p.rope_S) : 0;
        const float2* cs = p.rope_cs + static_cast<int64_t>(pos) * 64;
#pragma unroll 1
        for (int hh = 0; hh < BN / 128; ++hh) {
          const int colh = n0 + hh * 128;
          const bool rot = colh < p.rope_cols;
#pragma unroll 1
          for (int half = 0; half < 2; ++half) {
            uint32_t lo[32], hi[32];
            ptx::tmem_ld32(taddr + hh * 128 + half * 32, lo);
            ptx::tmem_ld32(taddr + hh * 128 + 64 + half * 32, hi);
            ptx::tmem_ld_wait();
            if (row_ok && colh < p.N) {
              float a[32], b[32];
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                a[j] = __uint_as_float(lo[j]) * p.alpha;
                b[j] = __uint_as_float(hi[j]) * p.alpha;
              }
              if (rot) {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                  const float2 k = __ldg(cs + half * 32 + j);
                  const float na = a[j] * k.x - b[j] * k.y;
                  const float nb = b[j] * k.x + a[j] * k.y;
                  a[j] = na;
                  b[j] = nb;
                }
              }
              uint4* dlo = reinterpret_cast<uint4*>(crow + colh + half * 32);
              uint4* dhi = reinterpret_cast<uint4*>(crow + colh + 64 + half * 32);
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
                dlo[q] = u;
                dhi[q] = w;
              }
            }
            __syncwarp();
          }
        }
      } else if constexpr (EPI == EPI_F32_ATOMIC) {
        float* crow = reinterpret_cast<float*>(p.C) + static_cast<int64_t>(row) * p.ldc;
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
          uint32_t r[32];
          ptx::tmem_ld32(taddr + c * 32, r);
          ptx::tmem_ld_wait();
          const int col0 = n0 + c * 32;
          if (row_ok && col0 < p.N) {
          if (col0 + 32 <= p.N) {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              red_add_v4(crow + col0 + j, __uint_as_float(r[j]) * p.alpha, __uint_as_float(r[j + 1]) * p.alpha,
                         __uint_as_float(r[j + 2]) * p.alpha, __uint_as_float(r[j + 3]) * p.alpha);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (col0 + j < p.N) atomicAdd(crow + col0 + j, __uint_as_float(r[j]) * p.alpha);
          }
          }
          __syncwarp();
        }
      } else if constexpr (EPI == EPI_F32_PARTIAL) {
        // split-K partial sums, plain stores into slice `split` of C [splits, M, ldc] (deterministic reduction later)
        float* crow = reinterpret_cast<float*>(p.C) + (static_cast<int64_t>(t.split) * p.M + row) * p.ldc;
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
          uint32_t r[32];
          ptx::tmem_ld32(taddr + c * 32, r);
          ptx::tmem_ld_wait();
          const int col0 = n0 + c * 32;
          if (row_ok && col0 < p.N) {
            if (col0 + 32 <= p.N) {
#pragma unroll
              for (int j = 0; j < 32; j += 4)
                *reinterpret_cast<float4*>(crow + col0 + j) =
                    make_float4(__uint_as_float(r[j]) * p.alpha, __uint_as_float(r[j + 1]) * p.alpha,
                                __uint_as_float(r[j + 2]) * p.alpha, __uint_as_float(r[j + 3]) * p.alpha);
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (col0 + j < p.N) crow[col0 + j] = __uint_as_float(r[j]) * p.alpha;
            }
          }
          __syncwarp();
        }
      } else if constexpr (EPI == EPI_LSE) {
        // Online log-sum-exp over this tile's columns, in base 2 (z2 = z*log2e).
        const int label = row_ok ? p.labels[row] : -1;
        float mx = -INFINITY, s = 0.f, zs = 0.f;
        const float a2 = p.alpha * LOG2E;
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
          uint32_t r[32];
          ptx::tmem_ld32(taddr + c * 32, r);
          ptx::tmem_ld_wait();
          const int col0 = n0 + c * 32;
          if (col0 < p.N) {
          const int nvalid = min(32, p.N - col0);
          float cm = -INFINITY;
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (j < nvalid) cm = fmaxf(cm, __uint_as_float(r[j]) * a2);
          const float nm = fmaxf(mx, cm);
          const float sc = exp2f(mx - nm);  // mx=-inf -> 0
          s *= sc;
          zs *= sc;
          mx = nm;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            if (j < nvalid) {
              const float z2 = __uint_as_float(r[j]) * a2;
              const float e = exp2f(z2 - mx);
              s += e;
              zs += z2 * e;
              if (col0 + j == label) p.label_logit[row] = __uint_as_float(r[j]) * p.alpha;
            }
          }
          }
          __syncwarp();
        }
        if (row_ok) p.lse_partials[static_cast<int64_t>(row) * p.num_n + t.n_blk] = make_float4(mx, s, zs, 0.f);
      } else if constexpr (EPI == EPI_CEGRAD) {
        __nv_bfloat16* crow = reinterpret_cast<__nv_bfloat16*>(p.C) + static_cast<int64_t>(row) * p.ldc;
        const int label = row_ok ? p.labels[row] : -1;
        // ignore_index rows carry no gradient (compute_logprobs, utils/common_utils.py:112-118)
        const bool live = row_ok && label >= 0;
        const float g = (live && p.dlp) ? p.dlp[row] * p.alpha : 0.f;
        // entropy head (OPA entropy regulariser, opa_models/opa_trainer.py:63,82-85): with H = -sum p ln p,
        // dH/dz_v = -p_v (ln p_v + H); ge carries d(loss)/dH * alpha, hrow the row's H
        const float ge = (live && p.dent) ? p.dent[row] * p.alpha : 0.f;
        const float hrow = (ge != 0.f) ? p.ent[row] : 0.f;
        const bool any = g != 0.f || ge != 0.f;
        const float lse2 = row_ok ? p.lse[row] * LOG2E : 0.f;
        const float a2 = p.alpha * LOG2E;
        constexpr float LN2 = 0.6931471805599453f;
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
          uint32_t r[32];
          ptx::tmem_ld32(taddr + c * 32, r);
          ptx::tmem_ld_wait();
          const int col0 = n0 + c * 32;
          if (row_ok && col0 < p.N) {
          float v[32];
          if (ge == 0.f) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float pr = any ? exp2f(__uint_as_float(r[j]) * a2 - lse2) : 0.f;
              v[j] = g * ((col0 + j == label ? 1.f : 0.f) - pr);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float lp2 = __uint_as_float(r[j]) * a2 - lse2;  // log2 p_v
              const float pr = exp2f(lp2);
              v[j] = g * ((col0 + j == label ? 1.f : 0.f) - pr) - ge * pr * (lp2 * LN2 + hrow);
            }
          }
          if (col0 + 32 <= p.N) {
            uint4* dst = reinterpret_cast<uint4*>(crow + col0);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              uint4 u;
              u.x = ptx::pack_bf16(v[q * 8 + 0], v[q * 8 + 1]);
              u.y = ptx::pack_bf16(v[q * 8 + 2], v[q * 8 + 3]);
              u.z = ptx::pack_bf16(v[q * 8 + 4], v[q * 8 + 5]);
              u.w = ptx::pack_bf16(v[q * 8 + 6], v[q * 8 + 7]);
              dst[q] = u;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (col0 + j < p.N) crow[col0 + j] = __float2bfloat16(v[j]);
          }
          }
          __syncwarp();
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (PAIR && rank != 0)
          ptx::mbar_arrive_remote(ptx::mapa(&tempty_bar[acc], 0));  // the leader's MMA thread owns the accumulator ring
        else
          ptx::mbar_arrive(&tempty_bar[acc]);
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (CL > 1) ptx::cluster_sync();  // nobody leaves while the peer may still multicast into / arrive on this CTA's smem
  if (warp == 2) {
    ptx::tc_fence_after();
    if (PAIR)
      ptx::tmem_dealloc2(tmem_base, 512);
    else
      ptx::tmem_dealloc(tmem_base, 512);
  }
}

thread_local char g_err[512] = {0};

// Optional per-launch timing (bench.py's roofline): CUDA events on the launching stream around every GEMM launch.
struct ProfRec {
  cudaEvent_t a, b;
  double flops;
};
bool g_prof_on = false;
std::vector<ProfRec> g_prof;

template <bool A_MN, bool B_MN, int EPI, int CL>
int launch_impl(const CUtensorMap* maps, const KParams& kp, int grid, cudaStream_t stream) {
  auto kern = gemm_kernel<A_MN, B_MN, EPI, CL>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) {
      snprintf(g_err, sizeof(g_err), "cudaFuncSetAttribute(smem=%d): %s", SMEM_BYTES, cudaGetErrorString(e));
      return -10;
    }
    attr_set = true;
  }
  ProfRec rec{};
  if (g_prof_on) {
    cudaEventCreate(&rec.a);
    cudaEventCreate(&rec.b);
    rec.flops = 2.0 * kp.M * kp.N * 64.0 * (kp.kb1 + kp.kb2);
    cudaEventRecord(rec.a, stream);
  }
  if (CL == 1) {
    kern<<<grid, NUM_THREADS, SMEM_BYTES, stream>>>(maps[0], maps[1], maps[2], maps[3], kp);
  } else {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid, 1, 1);
    cfg.blockDim = dim3(NUM_THREADS, 1, 1);
    cfg.dynamicSmemBytes = SMEM_BYTES;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t le = cudaLaunchKernelEx(&cfg, kern, maps[0], maps[1], maps[2], maps[3], kp);
    if (le != cudaSuccess) {
      snprintf(g_err, sizeof(g_err), "gemm cluster launch: %s", cudaGetErrorString(le));
      return -11;
    }
  }
  if (g_prof_on) {
    cudaEventRecord(rec.b, stream);
    g_prof.push_back(rec);
  }
  count_launch();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    snprintf(g_err, sizeof(g_err), "gemm launch: %s", cudaGetErrorString(e));
    return -11;
  }
  return 0;
}

int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

}  // namespace

const char* gemm_last_error() { return g_err; }

void gemm_prof_begin() {
  g_prof.clear();
  g_prof_on = true;
}
// Sums the event-timed durations (ms) and executed (tile-padded K) FLOPs of every launch since gemm_prof_begin().
int gemm_prof_end(double* total_ms, double* total_flops, int64_t* launches) {
  g_prof_on = false;
  cudaDeviceSynchronize();
  double ms = 0, fl = 0;
  for (auto& r : g_prof) {
    float t = 0.f;
    cudaEventElapsedTime(&t, r.a, r.b);
    ms += t;
    fl += r.flops;
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
  }
  *total_ms = ms;
  *total_flops = fl;
  *launches = static_cast<int64_t>(g_prof.size());
  g_prof.clear();
  return 0;
}
int gemm_num_n_tiles(int N) { return (N + BN - 1) / BN; }
int gemm_effective_splits(int K1, int K2, int splits) {
  const int kb_all = (K1 + BK - 1) / BK + (K2 + BK - 1) / BK;
  if (splits < 1) splits = 1;
  if (splits > kb_all) splits = kb_all;
  const int per = (kb_all + splits - 1) / splits;
  return (kb_all + per - 1) / per;
}

int gemm_launch(const GemmArgs& a, cudaStream_t stream) {
  if (a.M <= 0 || a.N <= 0 || a.K1 <= 0) {
    snprintf(g_err, sizeof(g_err), "gemm: bad dims M=%d N=%d K1=%d", a.M, a.N, a.K1);
    return -1;
  }
  if (a.a_mn_major && !a.b_mn_major) {
    snprintf(g_err, sizeof(g_err), "gemm: (A MN-major, B K-major) is not instantiated");
    return -2;
  }
  if (a.splits > 1 && a.epilogue != EPI_F32_ATOMIC && a.epilogue != EPI_F32_PARTIAL) {
    snprintf(g_err, sizeof(g_err), "gemm: split-K needs EPI_F32_ATOMIC or EPI_F32_PARTIAL");
    return -3;
  }
  if (a.a2_group_n > 0 && (a.a2_group_n % BN != 0 || a.a_mn_major)) {
    snprintf(g_err, sizeof(g_err), "gemm: a2_group_n must be a multiple of %d with K-major A", BN);
    return -4;
  }
  if (a.epilogue == EPI_BF16_ROPE &&
      (a.N % 128 != 0 || a.rope_cols % 128 != 0 || !a.rope_cs || a.bias || a.residual || a.act != ACT_NONE ||
       (!a.rope_pos && a.rope_S <= 0) || a.splits > 1)) {
    snprintf(g_err, sizeof(g_err), "gemm: EPI_BF16_ROPE needs N and rope_cols multiples of 128, a table, positions or rope_S, and no bias / act / residual");
    return -8;
  }
  if (a.epilogue == EPI_SWIGLU_BF16 && (a.N % 256 != 0 || a.bias || a.residual || a.act != ACT_NONE || a.splits > 1)) {
    snprintf(g_err, sizeof(g_err), "gemm: EPI_SWIGLU_BF16 needs N %% 256 == 0 (gate / up rows interleaved in blocks of 128) and no bias / act / residual");
    return -9;
  }
  if (a.epilogue == EPI_SWIGLU_BWD && (a.N % 32 != 0 || !a.residual || a.ldr % 8 != 0 || (reinterpret_cast<uintptr_t>(a.residual) & 15) ||
                                        a.bias || a.act != ACT_NONE || a.splits > 1 || a.ldc < 2 * static_cast<int64_t>(a.N))) {
    snprintf(g_err, sizeof(g_err), "gemm: EPI_SWIGLU_BWD needs N %% 32 == 0, the saved gate|up activations in `residual` (16 B-aligned, ldr %% 8 == 0), ldc >= 2 N, no bias / act / split-K");
    return -9;
  }
  if ((a.epilogue == EPI_BF16 || a.epilogue == EPI_CEGRAD || a.epilogue == EPI_BF16_ROPE || a.epilogue == EPI_SWIGLU_BF16 || a.epilogue == EPI_SWIGLU_BWD) &&
      (a.ldc % 8 != 0 || (reinterpret_cast<uintptr_t>(a.C) & 15))) {
    snprintf(g_err, sizeof(g_err), "gemm: bf16 output needs ldc %% 8 == 0 and 16B-aligned C");
    return -5;
  }
  if ((a.epilogue == EPI_F32_ATOMIC || a.epilogue == EPI_F32_PARTIAL) && (a.ldc % 4 != 0 || (reinterpret_cast<uintptr_t>(a.C) & 15))) {
    snprintf(g_err, sizeof(g_err), "gemm: f32 output needs ldc %% 4 == 0 and 16B-aligned C");
    return -5;
  }

  // Cluster (B-multicast) variant: two CTAs on vertically adjacent tiles share each B tile. Worth it when there are many
  // m blocks; not combined with split-K. a.cluster: -1 auto, 0 off, 1 on. OPADPO_GEMM_CLUSTER=0 disables it globally.
  // Two-CTA variants (a.cluster: -1 auto, 0 off, 1 = cluster with the B tile multicast, 2 = CTA pair / cta_group::2).
  // Measured (profiles/r01_gemm_cluster_ab.txt): the CTA pair is +8..16 % on the step's GEMM shapes (above cuBLAS on
  // most) and -33 ms on the 680 ms training step; the multicast cluster is +-0 in the step. Auto = CTA pair whenever
  // there are >= 8 m blocks, no split-K and K-major A. OPADPO_GEMM_CLUSTER=0/1/2 overrides auto.
  static const int cluster_env_mode = []() {
    const char* e = getenv("OPADPO_GEMM_CLUSTER");
    return e ? atoi(e) : 2;
  }();
  const int num_m_blocks = (a.M + BM - 1) / BM;
  const bool single_split = a.splits <= 1;
  int cl_mode = a.cluster < 0 ? (num_m_blocks >= 8 ? cluster_env_mode : 0) : a.cluster;
  if (!single_split || cl_mode < 0 || cl_mode > 2) cl_mode = 0;
  if (cl_mode == 2 && a.a_mn_major) cl_mode = 0;  // the CTA-pair kernel is instantiated for K-major A only
  bool use_cluster = cl_mode > 0;
  const int b_box_rows = use_cluster ? BN / 2 : BN;

  CUtensorMap maps[4];
  auto mk = [&](CUtensorMap* m, const GemmOperand& op, bool mn_major, int rows_out, int K, int box_rows_kmajor,
                int extra_cols) -> int {
    if (!mn_major) return make_tmap_bf16_2d(m, op.ptr, rows_out, (uint64_t)K + extra_cols, op.ld, box_rows_kmajor, 64);
    return make_tmap_bf16_2d(m, op.ptr, K, rows_out, op.ld, 64, 64);
  };
  int groups = a.a2_group_n > 0 ? (a.N + a.a2_group_n - 1) / a.a2_group_n : 1;
  int rc = 0;
  rc |= mk(&maps[0], a.A1, a.a_mn_major, a.M, a.K1, BM, 0);
  rc |= mk(&maps[1], a.B1, a.b_mn_major, a.N, a.K1, b_box_rows, 0);
  if (a.K2 > 0) {
    rc |= mk(&maps[2], a.A2, a.a_mn_major, a.M, a.K2, BM, (groups - 1) * a.a2_group_stride);
    rc |= mk(&maps[3], a.B2, a.b_mn_major, a.N, a.K2, b_box_rows, 0);
  } else {
    maps[2] = maps[0];
    maps[3] = maps[1];
  }
  if (rc != 0) {
    snprintf(g_err, sizeof(g_err), "gemm: tensor map: %s", tmap_last_error());
    return -6;
  }

  KParams kp{};
  kp.M = a.M;
  kp.N = a.N;
  kp.kb1 = (a.K1 + BK - 1) / BK;
  kp.kb2 = (a.K2 + BK - 1) / BK;
  kp.a2_group_n = a.a2_group_n;
  kp.a2_group_stride = a.a2_group_stride;
  int splits = a.splits < 1 ? 1 : a.splits;
  const int kb_all = kp.kb1 + kp.kb2;
  if (splits > kb_all) splits = kb_all;
  kp.kb_per_split = (kb_all + splits - 1) / splits;
  kp.splits = (kb_all + kp.kb_per_split - 1) / kp.kb_per_split;
  kp.num_m = (a.M + BM - 1) / BM;
  kp.num_n = (a.N + BN - 1) / BN;
  kp.num_tiles = kp.num_m * kp.num_n * kp.splits;
  {
    // Rasterisation: keep about 32 MB of A tiles L2-resident per group so each B panel is streamed from HBM
    // num_m/group_m times (OPADPO_GEMM_GROUP_M overrides).
    static const int env_gm = []() {
      const char* e = getenv("OPADPO_GEMM_GROUP_M");
      return e ? atoi(e) : 0;
    }();
    const double tile_bytes = 2.0 * BM * 64.0 * (kp.kb1 + kp.kb2);
    int gm = static_cast<int>(32.0 * 1024 * 1024 / tile_bytes);
    gm = gm < 8 ? 8 : (gm > 64 ? 64 : gm);
    gm &= ~1;
    kp.group_m = env_gm > 0 ? env_gm : gm;
  }
  kp.alpha = a.alpha;
  kp.C = a.C;
  kp.ldc = a.ldc;
  kp.bias = a.bias;
  kp.residual = a.residual;
  kp.ldr = a.ldr;
  kp.act = a.act;
  kp.labels = a.labels;
  kp.lse_partials = a.lse_partials;
  kp.label_logit = a.label_logit;
  kp.lse = a.lse;
  kp.dlp = a.dlp;
  kp.dent = a.dent;
  kp.ent = a.ent;
  kp.rope_pos = a.rope_pos;
  kp.rope_cs = a.rope_cs;
  kp.rope_cols = a.rope_cols;
  kp.rope_S = a.rope_S;

  int grid = num_sms();
  if (a.max_ctas > 0 && a.max_ctas < grid) grid = a.max_ctas;
  if (use_cluster) {
    const int pair_tiles = ((kp.num_m + 1) / 2) * kp.num_n;
    grid &= ~1;
    if (grid > 2 * pair_tiles) grid = 2 * pair_tiles;
    if (grid < 2) use_cluster = false;
  }
  if (!use_cluster && grid > kp.num_tiles) grid = kp.num_tiles;

#define OPADPO_DISPATCH_CL(AMN, BMN, CL)                                                          \
  switch (a.epilogue) {                                                                           \
    case EPI_BF16: return launch_impl<AMN, BMN, EPI_BF16, CL>(maps, kp, grid, stream);            \
    case EPI_F32_ATOMIC: return launch_impl<AMN, BMN, EPI_F32_ATOMIC, CL>(maps, kp, grid, stream); \
    case EPI_LSE: return launch_impl<AMN, BMN, EPI_LSE, CL>(maps, kp, grid, stream);              \
    case EPI_CEGRAD: return launch_impl<AMN, BMN, EPI_CEGRAD, CL>(maps, kp, grid, stream);        \
    case EPI_F32_PARTIAL: return launch_impl<AMN, BMN, EPI_F32_PARTIAL, CL>(maps, kp, grid, stream); \
    case EPI_BF16_ROPE: return launch_impl<AMN, BMN, EPI_BF16_ROPE, CL>(maps, kp, grid, stream);   \
    case EPI_SWIGLU_BF16: return launch_impl<AMN, BMN, EPI_SWIGLU_BF16, CL>(maps, kp, grid, stream); \
    case EPI_SWIGLU_BWD: return launch_impl<AMN, BMN, EPI_SWIGLU_BWD, CL>(maps, kp, grid, stream); \
    default: break;                                                                               \
  }
#define OPADPO_DISPATCH(AMN, BMN)          \
  if (use_cluster) {                       \
    OPADPO_DISPATCH_CL(AMN, BMN, 2)        \
  } else {                                 \
    OPADPO_DISPATCH_CL(AMN, BMN, 1)        \
  }
#define OPADPO_DISPATCH_PAIR(BMN)          \
  if (use_cluster && cl_mode == 2) {       \
    OPADPO_DISPATCH_CL(false, BMN, 3)      \
  }
  if (!a.a_mn_major && !a.b_mn_major) {
    OPADPO_DISPATCH_PAIR(false)
    OPADPO_DISPATCH(false, false)
  } else if (!a.a_mn_major && a.b_mn_major) {
    OPADPO_DISPATCH_PAIR(true)
    OPADPO_DISPATCH(false, true)
  } else {
    OPADPO_DISPATCH(true, true)
  }
#undef OPADPO_DISPATCH_PAIR
#undef OPADPO_DISPATCH
#undef OPADPO_DISPATCH_CL
  snprintf(g_err, sizeof(g_err), "gemm: unknown epilogue %d", a.epilogue);
  return -7;
}

}  // namespace opadpo
