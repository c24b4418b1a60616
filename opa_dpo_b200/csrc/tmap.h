// Host-side TMA tensor-map construction (cuTensorMapEncodeTiled resolved through the runtime's
// driver entry point, so the library links against cudart only).
#pragma once
#include <cuda.h>
#include <cstdint>

namespace opadpo {

// 2-D bf16 tensor map over a row-major [rows, cols] view with leading dimension ld (elements).
// Box = box_rows x box_cols (box_cols*2 bytes must be 128 for the 128B swizzle used by every kernel).
// Out-of-bounds elements are filled with zeros. Returns 0 on success.
int make_tmap_bf16_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                      uint32_t box_rows, uint32_t box_cols);

const char* tmap_last_error();

}  // namespace opadpo
