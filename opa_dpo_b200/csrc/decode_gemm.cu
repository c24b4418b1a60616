// Weight-streaming tcgen05 GEMM for the rollout's decode step (a handful of rows against a whole weight matrix):
//
//   part[split, m, n] = alpha * sum_{k in split} x[m, k] * W[n, k]        m < Mb <= 128,  W [N, K] bf16 K-major
//
// The roles of the operands are swapped against gemm.cu: the WEIGHT rows sit on the 128 TMEM lanes (MMA M = 128) and the
// few activation rows are the MMA's N (32 / 64 / 128), so a pipeline stage is 16 KB of weights + 4..16 KB of activations
// instead of 16 KB of mostly padding + 32 KB of weights. That buys 8 stages (128 KB of weight bytes in flight per SM, what
// a CTA needs to pull its share of HBM bandwidth at ~2.5 us of load latency) and 128-row weight tiles, i.e. twice as many
// tiles to spread over the SMs before any split-K is needed. Split-K slices are plain fp32 stores, reduced in a fixed order by
// the consumer kernels (elementwise.cu: partial_*), so the rollout stays run-to-run deterministic.
//
//   warp 0 : TMA producer      warp 1 : MMA issuer      warp 2 : TMEM allocator      warps 4-7 : epilogue (thread = weight row)
#include "decode_gemm.h"

#include <cstdio>

#include "common.h"
#include "ptx.cuh"
#include "tmap.h"

namespace opadpo {
namespace {

constexpr int DM = 128;  // weight rows per tile (MMA M)
constexpr int DK = 64;   // k-block
constexpr int W_BYTES = DM * DK * 2;
constexpr int D_THREADS = 256;

template <int BNB>
struct DecodeCfg {
  static constexpr int X_BYTES = BNB * DK * 2;
  static constexpr int STAGE = W_BYTES + X_BYTES;
  static constexpr int NSTAGE = BNB <= 64 ? 8 : 6;
  static constexpr int SMEM = NSTAGE * STAGE + 1024 + 256;
  static constexpr int TMEM_COLS = 2 * BNB < 32 ? 32 : 2 * BNB;  // two accumulator buffers (power of two: 64 / 128 / 256)
};

struct DParams {
  int Mb, N;
  int kb_all, kb_per_split, splits;
  int num_n, items;
  float alpha;
  float* part;
  int64_t ldc;
};

template <int BNB>
__global__ void __launch_bounds__(D_THREADS, 1)
    decode_gemm_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmX, const DParams p) {
  using C = DecodeCfg<BNB>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + C::NSTAGE * C::STAGE);
  uint64_t* empty_bar = full_bar + 8;
  uint64_t* tfull_bar = empty_bar + 8;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0) {
    if (ptx::elect_one()) {
      ptx::prefetch_tmap(&tmW);
      ptx::prefetch_tmap(&tmX);
    }
  } else if (warp == 1) {
    if (ptx::elect_one()) {
      for (int i = 0; i < C::NSTAGE; ++i) {
        ptx::mbar_init(&full_bar[i], 1);
        ptx::mbar_init(&empty_bar[i], 1);
      }
      for (int i = 0; i < 2; ++i) {
        ptx::mbar_init(&tfull_bar[i], 1);
        ptx::mbar_init(&tempty_bar[i], 4);
      }
      ptx::fence_barrier_init();
    }
  } else if (warp == 2) {
    ptx::tmem_alloc(tmem_slot, C::TMEM_COLS);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // work item -> (weight tile, K split): consecutive CTAs take consecutive weight tiles of the same split, so the CTAs that
  // run side by side read the same activation k-range out of L2
  auto kb_range = [&](int item, int& n_blk, int& split, int& kb0, int& kb1) {
    split = item / p.num_n;
    n_blk = item - split * p.num_n;
    kb0 = split * p.kb_per_split;
    kb1 = min(p.kb_all, kb0 + p.kb_per_split);
  };

  if (warp == 0) {
    if (ptx::elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
        int n_blk, split, kb0, kb1;
        kb_range(item, n_blk, split, kb0, kb1);
        for (int kb = kb0; kb < kb1; ++kb) {
          ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sW = smem + stage * C::STAGE;
          ptx::mbar_arrive_expect_tx(&full_bar[stage], C::STAGE);
          ptx::tma_load_2d(sW, &tmW, &full_bar[stage], kb * DK, n_blk * DM);
          ptx::tma_load_2d(sW + W_BYTES, &tmX, &full_bar[stage], kb * DK, 0);
          if (++stage == C::NSTAGE) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (ptx::elect_one()) {
      constexpr uint32_t idesc = ptx::make_idesc_bf16(DM, BNB, false, false);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
        int n_blk, split, kb0, kb1;
        kb_range(item, n_blk, split, kb0, kb1);
        ptx::mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BNB;
        uint32_t accumulate = 0;
        for (int kb = kb0; kb < kb1; ++kb) {
          ptx::mbar_wait(&full_bar[stage], phase);
          ptx::tc_fence_after();
          const uint32_t w_addr = ptx::smem_u32(smem + stage * C::STAGE);
          const uint32_t x_addr = w_addr + W_BYTES;
#pragma unroll
          for (int k = 0; k < DK / 16; ++k) {
            ptx::mma_bf16_ss(d_tmem, ptx::make_smem_desc(w_addr + k * 32, 16, 1024), ptx::make_smem_desc(x_addr + k * 32, 16, 1024),
                             idesc, accumulate);
            accumulate = 1;
          }
          ptx::mma_commit(&empty_bar[stage]);
          if (++stage == C::NSTAGE) {
            stage = 0;
            phase ^= 1;
          }
        }
        ptx::mma_commit(&tfull_bar[acc]);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else if (warp >= 4) {
    // epilogue: thread == weight row n; its BNB accumulator columns are the activation rows m. For every m the 32 lanes of
    // a warp store 32 consecutive n: one 128 B line per (warp, m).
    const int quad = warp - 4;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
      int n_blk, split, kb0, kb1;
      kb_range(item, n_blk, split, kb0, kb1);
      ptx::mbar_wait(&tfull_bar[acc], acc_phase);
      ptx::tc_fence_after();
      const int n = n_blk * DM + quad * 32 + lane;
      const uint32_t taddr = tmem_base + acc * BNB + (static_cast<uint32_t>(quad * 32) << 16);
      float* out = p.part + static_cast<int64_t>(split) * p.Mb * p.ldc + n;
#pragma unroll 1
      for (int c = 0; c < BNB / 32; ++c) {
        uint32_t r[32];
        ptx::tmem_ld32(taddr + c * 32, r);
        ptx::tmem_ld_wait();
        if (n < p.N) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int m = c * 32 + j;
            if (m < p.Mb) out[static_cast<int64_t>(m) * p.ldc] = __uint_as_float(r[j]) * p.alpha;
          }
        }
        __syncwarp();
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&tempty_bar[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

thread_local char g_err[512] = {0};

int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

template <int BNB>
int launch(const DecodeGemmArgs& a, const DParams& p, cudaStream_t s) {
  using C = DecodeCfg<BNB>;
  CUtensorMap mw, mx;
  if (make_tmap_bf16_2d(&mw, a.W, a.N, a.K, a.ldw, DM, DK) || make_tmap_bf16_2d(&mx, a.x, a.Mb, a.K, a.ldx, BNB, DK)) {
    snprintf(g_err, sizeof(g_err), "decode gemm: tensor map: %s", tmap_last_error());
    return -6;
  }
  auto kern = decode_gemm_kernel<BNB>;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
    if (e != cudaSuccess) {
      snprintf(g_err, sizeof(g_err), "decode gemm: cudaFuncSetAttribute(%d): %s", C::SMEM, cudaGetErrorString(e));
      return -10;
    }
    attr = true;
  }
  const int grid = p.items < sm_count() ? p.items : sm_count();
  kern<<<grid, D_THREADS, C::SMEM, s>>>(mw, mx, p);
  count_launch();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    snprintf(g_err, sizeof(g_err), "decode gemm launch: %s", cudaGetErrorString(e));
    return -11;
  }
  return 0;
}

}  // namespace

const char* decode_gemm_last_error() { return g_err; }

// Split count: the persistent grid takes ceil(items / SMs) items per CTA, each kb_per_split k-blocks long; every extra split
// writes and re-reads one more fp32 slice [Mb, N] (priced in k-block equivalents of one CTA's 16 KB weight tile).
int decode_gemm_pick_splits(int Mb, int N, int K, int max_splits) {
  const int kb = (K + DK - 1) / DK;
  const int num_n = (N + DM - 1) / DM;
  const int G = sm_count();
  const double slice_cost = static_cast<double>(Mb) * N * 8.0 / (static_cast<double>(G) * W_BYTES);
  int best = 1;
  double best_cost = 1e30;
  const int cap = max_splits > 0 ? max_splits : 16;
  for (int sp = 1; sp <= cap && sp <= kb; ++sp) {
    const int per = (kb + sp - 1) / sp;
    const int eff = (kb + per - 1) / per;
    if (eff != sp) continue;  // same slicing as a smaller count
    const int items = num_n * sp;
    const double cost = static_cast<double>((items + G - 1) / G) * per + slice_cost * sp;
    if (cost < best_cost - 1e-9) {
      best_cost = cost;
      best = sp;
    }
  }
  return best;
}

int decode_gemm_launch(const DecodeGemmArgs& a, cudaStream_t s) {
  if (a.Mb <= 0 || a.Mb > 128 || a.N <= 0 || a.K <= 0 || a.splits < 1) {
    snprintf(g_err, sizeof(g_err), "decode gemm: bad dims Mb=%d (1..128) N=%d K=%d splits=%d", a.Mb, a.N, a.K, a.splits);
    return -1;
  }
  if (a.ldc % 1 != 0 || (reinterpret_cast<uintptr_t>(a.part) & 3)) {
    snprintf(g_err, sizeof(g_err), "decode gemm: misaligned output");
    return -5;
  }
  DParams p{};
  p.Mb = a.Mb;
  p.N = a.N;
  p.kb_all = (a.K + DK - 1) / DK;
  int sp = a.splits > p.kb_all ? p.kb_all : a.splits;
  p.kb_per_split = (p.kb_all + sp - 1) / sp;
  p.splits = (p.kb_all + p.kb_per_split - 1) / p.kb_per_split;
  if (p.splits != a.splits) {
    snprintf(g_err, sizeof(g_err), "decode gemm: %d splits do not slice %d k-blocks evenly (use decode_gemm_pick_splits)", a.splits,
             p.kb_all);
    return -3;
  }
  p.num_n = (a.N + DM - 1) / DM;
  p.items = p.num_n * p.splits;
  p.alpha = a.alpha;
  p.part = a.part;
  p.ldc = a.ldc;
  if (a.Mb <= 32) return launch<32>(a, p, s);
  if (a.Mb <= 64) return launch<64>(a, p, s);
  return launch<128>(a, p, s);
}

}  // namespace opadpo
