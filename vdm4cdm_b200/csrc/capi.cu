// Library-level entry points: version, thread-local error string, device check.
#include "common.cuh"

namespace vdm {
static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace vdm

extern "C" int vdm_version(void) { return 100; /* 0.1.0 */ }

extern "C" const char* vdm_last_error_string(void) { return vdm::g_err; }

extern "C" int vdm_device_supported(int dev) {
  cudaDeviceProp prop;
  cudaError_t e = cudaGetDeviceProperties(&prop, dev);
  if (e != cudaSuccess) {
    vdm::set_error("cudaGetDeviceProperties(%d): %s", dev, cudaGetErrorString(e));
    return VDM_E_CUDA;
  }
  if (prop.major != 10) {
    vdm::set_error("device %d is sm_%d%d; this library carries sm_100a code only", dev, prop.major,
                   prop.minor);
    return VDM_E_UNSUPPORTED;
  }
  return VDM_OK;
}

// *counter += 1 on the stream (step counter of a replayed CUDA graph).
__global__ void increment_kernel(int32_t* c) { *c += 1; }

extern "C" int vdm_increment(int32_t* counter, void* stream) {
  VDM_CHECK_ARG(counter != nullptr, "vdm_increment: counter is NULL");
  increment_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(counter);
  VDM_CHECK_LAUNCH();
  return VDM_OK;
}
