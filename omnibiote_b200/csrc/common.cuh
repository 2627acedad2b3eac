// Shared host-side helpers for the C-ABI library: error reporting, launch checks.
#pragma once
#include <cuda_runtime.h>
#include <cuda.h>
#include <stdint.h>
#include <stdio.h>

namespace obt {

// Error codes returned across the C ABI (0 = success). See include/omnibiote_b200.h.
enum : int {
  OBT_OK = 0,
  OBT_ERR_INVALID = -1,   // bad argument (null pointer, misaligned stride, unsupported shape)
  OBT_ERR_CUDA = -2,      // CUDA runtime / driver error
  OBT_ERR_UNSUPPORTED = -3,
};

void set_last_error(const char* fmt, ...);
int check_launch(const char* what);

// Builds (or fetches from the cache) a 2D bf16 tensor map with 128B swizzle.
//   dim0 = contiguous extent (elements), dim1 = number of rows, stride1 = row pitch in elements.
//   box0 must be 64 (128 bytes); box1 <= 256.
int get_tensor_map_2d(CUtensorMap* out, const void* base, uint64_t dim0, uint64_t dim1, uint64_t stride1_elems,
                      uint32_t box0, uint32_t box1);
// 3D variant: dim0 contiguous, dim1 with stride1, dim2 with stride2 (elements).
int get_tensor_map_3d(CUtensorMap* out, const void* base, uint64_t dim0, uint64_t dim1, uint64_t dim2,
                      uint64_t stride1_elems, uint64_t stride2_elems, uint32_t box0, uint32_t box1, uint32_t box2);

int sm_count();

}  // namespace obt

#define OBT_REQUIRE(cond, ...)                 \
  do {                                         \
    if (!(cond)) {                             \
      ::obt::set_last_error(__VA_ARGS__);      \
      return ::obt::OBT_ERR_INVALID;           \
    }                                          \
  } while (0)
