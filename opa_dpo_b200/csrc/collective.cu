// Gradient all-reduce of the data-parallel OPA-DPO step as ONE kernel over NVLink 5 / NVSwitch peer memory:
//   fp32 LoRA-grad slice --pack--> bf16 staging (symmetric memory, mapped by every rank)
//   --[multimem.ld_reduce across ranks, fp32 accumulate in the switch]--> --[multimem.st to every rank]-->
//   --unpack--> the same fp32 slice, now holding the sum over ranks (bit-identical on every rank).
// Replaces the DDP gradient reduction the reference's accelerate loop would perform when its reducer is armed
// (opadpo/dpo_models/rl_trainer.py:155-175; torch DDP + NCCL ring all-reduce [UPSTREAM]).
//
// Why not NCCL here: the collective runs UNDER the backward's persistent 148-CTA tcgen05 GEMMs. NCCL's channels are CTAs with
// their own shared memory, so a GEMM CTA (≈200 KB smem) cannot be co-resident with them and whole GEMM waves serialise behind
// the collective (round 1: +21.5 ms of GEMM time per step at 8 GPUs). This kernel uses no shared memory and few registers: its
// CTAs fit beside a resident GEMM CTA on the same SM and only compete for issue slots; the wire carries bf16 (half the bytes),
// the reduction itself happens in the NVSwitch (NVLS), each rank pulling only its 1/world share.
//
// Work split. The slice is cut into units of 8 elements (16 B of staging). Region (r, b) — rank r, block b — is a contiguous
// run of units. Block b of EVERY rank packs the regions (*, b) of its local slice, so a per-block-index barrier across ranks
// (one signal word per (block, source rank), monotonic epochs, release/acquire at system scope) is all the synchronisation the
// reduce phase of block b on any rank needs; the same holds for the unpack phase after the exit barrier.
#include <cstdio>

#include <cuda_bf16.h>

#include "../../include/opadpo_b200.h"
#include "common.h"

namespace opadpo {
namespace {

constexpr int AR_THREADS = 512;

struct ArParams {
  float* grad;               // local fp32 slice
  uint4* stage;              // local bf16 staging of the slice (symmetric allocation), 16 B units
  uint64_t stage_mc;         // multicast address of the same bytes (0: no NVLS, use the peer mappings)
  uint64_t stage_peer[OPADPO_MAX_RANKS];
  uint64_t signal_peer[OPADPO_MAX_RANKS];  // uint32 [max_blocks][OPADPO_MAX_RANKS] on every rank
  int64_t units;             // count / 8
  int64_t per;               // units per region = ceil(units / (world * blocks))
  uint32_t epoch;            // entry barrier = epoch, exit barrier = epoch + 1
  int rank, world;
  int fp32_in_place;         // 1: `grad` itself is the symmetric buffer (stage_* map the fp32 slice), no staging, exact fp32 sums
  uint32_t* err;             // host-visible flag: set when a barrier timed out
  uint64_t timeout_ns;
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint64_t globaltimer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ uint4 ld_volatile_16(const void* p) {
  uint4 v;
  asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint4 multimem_ld_reduce_bf16x8(uint64_t mc_addr) {
  uint4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.bf16x2 {%0, %1, %2, %3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(mc_addr)
               : "memory");
  return v;
}
__device__ __forceinline__ void multimem_st_16(uint64_t mc_addr, uint4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc_addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}
__device__ __forceinline__ uint4 multimem_ld_reduce_f32x4(uint64_t mc_addr) {
  uint4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(mc_addr)
               : "memory");
  return v;
}
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float bf_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }

// Barrier among the blocks with this blockIdx.x on all ranks. Thread t < world signals rank t and waits for rank t's signal.
__device__ __forceinline__ void block_barrier_across_ranks(const ArParams& p, uint32_t epoch) {
  __syncthreads();
  if (threadIdx.x < p.world) {
    const int t = threadIdx.x;
    __threadfence_system();  // this block's stores (observed through the barrier above) before the signal
    uint32_t* theirs = reinterpret_cast<uint32_t*>(p.signal_peer[t]) + blockIdx.x * OPADPO_MAX_RANKS + p.rank;
    st_release_sys(theirs, epoch);
    const uint32_t* mine = reinterpret_cast<const uint32_t*>(p.signal_peer[p.rank]) + blockIdx.x * OPADPO_MAX_RANKS + t;
    const uint64_t t0 = globaltimer_ns();
    // epochs only grow: a peer that is already one barrier ahead has written a larger value (wrap-safe signed distance)
    while (static_cast<int32_t>(ld_acquire_sys(mine) - epoch) < 0) {
      if (globaltimer_ns() - t0 > p.timeout_ns) {
        *reinterpret_cast<volatile uint32_t*>(p.err) = 0x80000000u | (epoch & 0xffffu) << 8 | static_cast<uint32_t>(t);
        __threadfence_system();
        break;
      }
    }
  }
  __syncthreads();
}

// fp32 in place: the gradient bucket itself lives in symmetric memory. No staging and no rounding: region (rank, b) of the
// slice is summed over ranks (in the switch, or by peer loads in rank order) and written back to every rank. Units of 4 floats.
__device__ __forceinline__ void allreduce_fp32_in_place(const ArParams& p) {
  const int b = blockIdx.x, G = gridDim.x;
  block_barrier_across_ranks(p, p.epoch);  // every rank's slice is final (its compute stream reached this launch)
  const int64_t lo = (static_cast<int64_t>(p.rank) * G + b) * p.per;
  const int64_t hi = min(lo + p.per, p.units);
  if (p.stage_mc != 0) {
    constexpr int UN = 4;
    int64_t u = lo + threadIdx.x;
    for (; u + (UN - 1) * AR_THREADS < hi; u += UN * AR_THREADS) {
      uint4 v[UN];
#pragma unroll
      for (int k = 0; k < UN; ++k) v[k] = multimem_ld_reduce_f32x4(p.stage_mc + 16ull * (u + k * AR_THREADS));
#pragma unroll
      for (int k = 0; k < UN; ++k) multimem_st_16(p.stage_mc + 16ull * (u + k * AR_THREADS), v[k]);
    }
    for (; u < hi; u += AR_THREADS) multimem_st_16(p.stage_mc + 16ull * u, multimem_ld_reduce_f32x4(p.stage_mc + 16ull * u));
  } else {
    for (int64_t u = lo + threadIdx.x; u < hi; u += AR_THREADS) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int r = 0; r < p.world; ++r) {
        const uint4 w = ld_volatile_16(reinterpret_cast<const uint4*>(p.stage_peer[r]) + u);
        acc.x += __uint_as_float(w.x);
        acc.y += __uint_as_float(w.y);
        acc.z += __uint_as_float(w.z);
        acc.w += __uint_as_float(w.w);
      }
      for (int r = 0; r < p.world; ++r) reinterpret_cast<float4*>(p.stage_peer[r])[u] = acc;
    }
  }
  block_barrier_across_ranks(p, p.epoch + 1);  // every rank's share has landed here; nobody still reads this rank's slice
}

__global__ void __launch_bounds__(AR_THREADS) allreduce_bf16wire_kernel(const ArParams p) {
  const int b = blockIdx.x, G = gridDim.x;
  if (p.fp32_in_place) {
    allreduce_fp32_in_place(p);
    return;
  }
  // ---- phase 0: pack the regions (*, b) of the local fp32 slice into the local staging
  for (int r = 0; r < p.world; ++r) {
    const int64_t lo = (static_cast<int64_t>(r) * G + b) * p.per;
    const int64_t hi = min(lo + p.per, p.units);
    for (int64_t u = lo + threadIdx.x; u < hi; u += AR_THREADS) {
      const float4 a = __ldcs(reinterpret_cast<const float4*>(p.grad) + 2 * u);
      const float4 c = __ldcs(reinterpret_cast<const float4*>(p.grad) + 2 * u + 1);
      uint4 w;
      w.x = pack2(a.x, a.y);
      w.y = pack2(a.z, a.w);
      w.z = pack2(c.x, c.y);
      w.w = pack2(c.z, c.w);
      p.stage[u] = w;
    }
  }
  block_barrier_across_ranks(p, p.epoch);
  // ---- phase 1: region (rank, b): sum over ranks, result to every rank
  {
    const int64_t lo = (static_cast<int64_t>(p.rank) * G + b) * p.per;
    const int64_t hi = min(lo + p.per, p.units);
    if (p.stage_mc != 0) {
      constexpr int UN = 4;  // independent 16 B reductions in flight per thread
      int64_t u = lo + threadIdx.x;
      for (; u + (UN - 1) * AR_THREADS < hi; u += UN * AR_THREADS) {
        uint4 v[UN];
#pragma unroll
        for (int k = 0; k < UN; ++k) v[k] = multimem_ld_reduce_bf16x8(p.stage_mc + 16ull * (u + k * AR_THREADS));
#pragma unroll
        for (int k = 0; k < UN; ++k) multimem_st_16(p.stage_mc + 16ull * (u + k * AR_THREADS), v[k]);
      }
      for (; u < hi; u += AR_THREADS) multimem_st_16(p.stage_mc + 16ull * u, multimem_ld_reduce_bf16x8(p.stage_mc + 16ull * u));
    } else {
      for (int64_t u = lo + threadIdx.x; u < hi; u += AR_THREADS) {
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (int r = 0; r < p.world; ++r) {  // fixed order: the owner computes each sum once
          const uint4 w = ld_volatile_16(reinterpret_cast<const uint4*>(p.stage_peer[r]) + u);
          acc[0] += bf_lo(w.x);
          acc[1] += bf_hi(w.x);
          acc[2] += bf_lo(w.y);
          acc[3] += bf_hi(w.y);
          acc[4] += bf_lo(w.z);
          acc[5] += bf_hi(w.z);
          acc[6] += bf_lo(w.w);
          acc[7] += bf_hi(w.w);
        }
        uint4 w;
        w.x = pack2(acc[0], acc[1]);
        w.y = pack2(acc[2], acc[3]);
        w.z = pack2(acc[4], acc[5]);
        w.w = pack2(acc[6], acc[7]);
        for (int r = 0; r < p.world; ++r) reinterpret_cast<uint4*>(p.stage_peer[r])[u] = w;
      }
    }
  }
  block_barrier_across_ranks(p, p.epoch + 1);
  // ---- phase 2: unpack the regions (*, b), written by block b of every rank, back into the fp32 slice
  for (int r = 0; r < p.world; ++r) {
    const int64_t lo = (static_cast<int64_t>(r) * G + b) * p.per;
    const int64_t hi = min(lo + p.per, p.units);
    for (int64_t u = lo + threadIdx.x; u < hi; u += AR_THREADS) {
      const uint4 w = ld_volatile_16(p.stage + u);
      float4 a, c;
      a.x = bf_lo(w.x);
      a.y = bf_hi(w.x);
      a.z = bf_lo(w.y);
      a.w = bf_hi(w.y);
      c.x = bf_lo(w.z);
      c.y = bf_hi(w.z);
      c.z = bf_lo(w.w);
      c.w = bf_hi(w.w);
      reinterpret_cast<float4*>(p.grad)[2 * u] = a;
      reinterpret_cast<float4*>(p.grad)[2 * u + 1] = c;
    }
  }
}

}  // namespace
}  // namespace opadpo

using namespace opadpo;

extern "C" int opadpo_op_allreduce_bf16wire(const opadpo_allreduce_desc* d, void* stream) {
  if (!d) {
    set_error("opadpo_op_allreduce_bf16wire: null descriptor");
    return OPADPO_EINVAL;
  }
  if (d->world < 2 || d->world > OPADPO_MAX_RANKS || d->rank < 0 || d->rank >= d->world) {
    set_error("opadpo_op_allreduce_bf16wire: rank %d / world %d (2..%d ranks)", d->rank, d->world, OPADPO_MAX_RANKS);
    return OPADPO_EINVAL;
  }
  if (d->blocks < 1 || d->blocks > OPADPO_AR_MAX_BLOCKS) {
    set_error("opadpo_op_allreduce_bf16wire: blocks %d outside 1..%d", d->blocks, OPADPO_AR_MAX_BLOCKS);
    return OPADPO_EINVAL;
  }
  if (d->count <= 0 || d->count % 8 != 0 || (reinterpret_cast<uintptr_t>(d->grad) & 31) != 0 ||
      (reinterpret_cast<uintptr_t>(d->stage_local) & 15) != 0 || (d->stage_mc & 15) != 0) {
    set_error("opadpo_op_allreduce_bf16wire: count %lld must be a positive multiple of 8, grad 32 B- and staging 16 B-aligned",
              static_cast<long long>(d->count));
    return OPADPO_EINVAL;
  }
  if (!d->grad || !d->stage_local || !d->err_flag) {
    set_error("opadpo_op_allreduce_bf16wire: null buffer");
    return OPADPO_EINVAL;
  }
  ArParams p{};
  p.grad = d->grad;
  p.stage = static_cast<uint4*>(d->stage_local);
  p.stage_mc = d->stage_mc;
  for (int r = 0; r < d->world; ++r) {
    if (!d->signal_peer[r] || (d->stage_mc == 0 && !d->stage_peer[r])) {
      set_error("opadpo_op_allreduce_bf16wire: missing peer mapping for rank %d", r);
      return OPADPO_EINVAL;
    }
    p.stage_peer[r] = d->stage_peer[r];
    p.signal_peer[r] = d->signal_peer[r];
  }
  p.fp32_in_place = d->fp32_in_place != 0;
  p.units = p.fp32_in_place ? d->count / 4 : d->count / 8;
  const int64_t regions = static_cast<int64_t>(d->world) * d->blocks;
  p.per = (p.units + regions - 1) / regions;
  p.epoch = d->epoch;
  p.rank = d->rank;
  p.world = d->world;
  p.err = d->err_flag;
  p.timeout_ns = d->timeout_ms > 0 ? static_cast<uint64_t>(d->timeout_ms) * 1000000ull : 60000000000ull;
  allreduce_bf16wire_kernel<<<d->blocks, AR_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(p);
  count_launch();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("allreduce_bf16wire launch: %s", cudaGetErrorString(e));
    return OPADPO_ECUDA;
  }
  return OPADPO_OK;
}
