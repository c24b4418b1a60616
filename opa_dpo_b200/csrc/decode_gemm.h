// Weight-streaming GEMM of the rollout's decode step (see decode_gemm.cu): part[split, m, n] = alpha * x[m, :] . W[n, :]
// over the split's k-range, for Mb <= 128 activation rows. fp32 slices [splits, Mb, ldc], reduced by the consumer.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdint>

namespace opadpo {

struct DecodeGemmArgs {
  int Mb = 0, N = 0, K = 0;
  const __nv_bfloat16* x = nullptr;  // [Mb, K] K-major
  int64_t ldx = 0;
  const __nv_bfloat16* W = nullptr;  // [N, K] K-major (a Linear's weight as stored)
  int64_t ldw = 0;
  float* part = nullptr;             // [splits, Mb, ldc]
  int64_t ldc = 0;
  int splits = 1;                    // a value returned by decode_gemm_pick_splits
  float alpha = 1.f;
};

int decode_gemm_launch(const DecodeGemmArgs& a, cudaStream_t s);
// K-split count that balances the persistent grid against the extra fp32 slice traffic (max_splits 0 = up to 16)
int decode_gemm_pick_splits(int Mb, int N, int K, int max_splits);
const char* decode_gemm_last_error();

}  // namespace opadpo
