// Host-side shared state of libdove_b200: error reporting, device info, tensor-map encoding.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/dove_b200.h"

namespace dove {

typedef __nv_bfloat16 bf16;

int set_error(int code, const char* fmt, ...);
int check_cuda(cudaError_t e, const char* what);
int num_sms();
int ensure_init();

// Encode a bf16 tiled tensor map (rank <= 5).  dims/strides innermost first; strides in BYTES for dims 1..rank-1.
// elem_strides may be null (all 1).  128B swizzle, zero OOB fill.
int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box, const uint32_t* elem_strides);

#define DOVE_CHECK_ARG(cond, ...)                                   \
  do {                                                              \
    if (!(cond)) return dove::set_error(DOVE_E_BAD_ARG, __VA_ARGS__); \
  } while (0)

#define DOVE_LAUNCH_CHECK(name)                                          \
  do {                                                                   \
    cudaError_t e__ = cudaGetLastError();                                \
    if (e__ != cudaSuccess) return dove::check_cuda(e__, name);          \
  } while (0)

}  // namespace dove
