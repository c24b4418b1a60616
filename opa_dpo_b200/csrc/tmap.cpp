#include "tmap.h"

#include <cuda_runtime.h>
#include <cudaTypedefs.h>

#include <cstdio>
#include <mutex>

namespace opadpo {

static thread_local char g_tmap_err[256] = {0};
const char* tmap_last_error() { return g_tmap_err; }

static PFN_cuTensorMapEncodeTiled_v12000 get_encode() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  static std::once_flag once;
  std::call_once(once, []() {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    if (e == cudaSuccess && q == cudaDriverEntryPointSuccess) fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  });
  return fn;
}

int make_tmap_bf16_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                      uint32_t box_rows, uint32_t box_cols) {
  auto enc = get_encode();
  if (!enc) {
    snprintf(g_tmap_err, sizeof(g_tmap_err), "cuTensorMapEncodeTiled entry point unavailable");
    return -1;
  }
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || ((ld * 2) & 15) != 0) {
    snprintf(g_tmap_err, sizeof(g_tmap_err), "tensor map needs 16B-aligned base and row pitch (base=%p ld=%llu)", base,
             (unsigned long long)ld);
    return -2;
  }
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    snprintf(g_tmap_err, sizeof(g_tmap_err),
             "cuTensorMapEncodeTiled failed (%d): rows=%llu cols=%llu ld=%llu box=%ux%u", (int)r,
             (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld, box_rows, box_cols);
    return -3;
  }
  return 0;
}

}  // namespace opadpo
