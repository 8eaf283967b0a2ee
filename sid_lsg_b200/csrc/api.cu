// Error reporting / version of the C ABI (include/sidlsg.h).  The library never allocates device
// memory, never synchronises and launches only on the stream it is handed; errors come back as a
// negative status plus a thread-local message (the Python shim raises RuntimeError, matching the
// TORCH_CHECK behaviour of the reference's own plugin idiom, /root/reference/torch_utils/ops/bias_act.cpp:23-54).
#include "common.cuh"
#include <stdarg.h>
#include <string.h>

namespace sidlsg {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return SIDLSG_ERR_CUDA;
  }
  return SIDLSG_OK;
}

}  // namespace sidlsg

extern "C" const char* sidlsg_last_error() { return sidlsg::g_err; }
extern "C" int sidlsg_version() { return 100; }
extern "C" int sidlsg_device_arch(int device) {
  cudaDeviceProp p;
  if (cudaGetDeviceProperties(&p, device) != cudaSuccess) {
    sidlsg::set_error("cudaGetDeviceProperties(%d) failed", device);
    return SIDLSG_ERR_CUDA;
  }
  return p.major * 10 + p.minor;
}
