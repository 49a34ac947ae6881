// libdove_b200 host-side state: init, error reporting, tensor-map encoding (driver entry point resolved at
// run time through the CUDA runtime, so the library has no link-time dependency on libcuda).
#include "common.cuh"

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cudaTypedefs.h>

namespace dove {

static thread_local char g_err[512] = "";
static int g_device = -1;
static int g_num_sms = 0;
static PFN_cuTensorMapEncodeTiled_v12000 g_encode = nullptr;

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return DOVE_OK;
  return set_error(DOVE_E_CUDA, "%s: %s", what, cudaGetErrorString(e));
}

int num_sms() { return g_num_sms; }

static int g_opt_conv2cta = 1;
int get_option_conv2cta() { return g_opt_conv2cta; }
int set_attn_variant(int v);

int ensure_init() {
  if (g_device < 0) return set_error(DOVE_E_NOT_INIT, "dove_init(device) has not been called");
  return DOVE_OK;
}

int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                   const uint64_t* strides_bytes, const uint32_t* box, const uint32_t* elem_strides) {
  cuuint64_t gdim[5];
  cuuint64_t gstr[4];
  cuuint32_t bx[5];
  cuuint32_t es[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    es[i] = elem_strides ? elem_strides[i] : 1;
    if (i > 0) gstr[i - 1] = strides_bytes[i - 1];
  }
  CUresult r = g_encode(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, static_cast<cuuint32_t>(rank),
                        const_cast<void*>(base), gdim, gstr, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_error(DOVE_E_CUDA, "cuTensorMapEncodeTiled failed (CUresult %d, rank %d, dims %llu/%llu box %u/%u)",
                     static_cast<int>(r), rank, static_cast<unsigned long long>(dims[0]),
                     static_cast<unsigned long long>(rank > 1 ? dims[1] : 0), box[0], rank > 1 ? box[1] : 0);
  return DOVE_OK;
}

}  // namespace dove

using namespace dove;

extern "C" int dove_set_option(const char* name, int value) {
  if (name && !strcmp(name, "conv2cta")) {
    g_opt_conv2cta = value;
    return DOVE_OK;
  }
  if (name && !strcmp(name, "attn_variant")) return set_attn_variant(value);
  return set_error(DOVE_E_BAD_ARG, "unknown option %s", name ? name : "(null)");
}

extern "C" int dove_abi_version(void) { return DOVE_ABI_VERSION; }

extern "C" const char* dove_last_error(void) { return g_err; }

extern "C" int dove_num_sms(void) { return g_num_sms; }

extern "C" int dove_init(int device) {
  if (g_device == device && g_encode) return DOVE_OK;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    return set_error(DOVE_E_CUDA, "no CUDA device available (%s): libdove_b200 has no CPU fallback",
                     cudaGetErrorString(e));
  if (int r = check_cuda(cudaSetDevice(device), "cudaSetDevice")) return r;
  cudaDeviceProp prop;
  if (int r = check_cuda(cudaGetDeviceProperties(&prop, device), "cudaGetDeviceProperties")) return r;
  if (prop.major != 10)
    return set_error(DOVE_E_UNSUPPORTED, "device %d is sm_%d%d; libdove_b200 is built for sm_100a only", device,
                     prop.major, prop.minor);
  g_num_sms = prop.multiProcessorCount;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !fn)
    return set_error(DOVE_E_CUDA, "cannot resolve cuTensorMapEncodeTiled (%s)", cudaGetErrorString(e));
  g_encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
  g_device = device;
  return DOVE_OK;
}

extern "C" int dove_h2d_box_async(const void* src, int64_t src_row_pitch, int64_t src_rows_per_plane, void* dst,
                                  int64_t row_bytes, int64_t rows, int64_t planes, void* stream) {
  if (int e = ensure_init()) return e;
  DOVE_CHECK_ARG(src && dst && row_bytes > 0 && rows > 0 && planes > 0, "h2d_box: empty copy");
  DOVE_CHECK_ARG(src_row_pitch >= row_bytes && src_rows_per_plane >= rows, "h2d_box: box larger than the source planes");
  cudaMemcpy3DParms p{};
  p.srcPtr = make_cudaPitchedPtr(const_cast<void*>(src), static_cast<size_t>(src_row_pitch),
                                 static_cast<size_t>(src_row_pitch), static_cast<size_t>(src_rows_per_plane));
  p.dstPtr = make_cudaPitchedPtr(dst, static_cast<size_t>(row_bytes), static_cast<size_t>(row_bytes),
                                 static_cast<size_t>(rows));
  p.extent = make_cudaExtent(static_cast<size_t>(row_bytes), static_cast<size_t>(rows), static_cast<size_t>(planes));
  p.kind = cudaMemcpyHostToDevice;
  return check_cuda(cudaMemcpy3DAsync(&p, static_cast<cudaStream_t>(stream)), "cudaMemcpy3DAsync(h2d_box)");
}
