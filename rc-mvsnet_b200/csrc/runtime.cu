// Error reporting + device check.
#include <stdarg.h>

#include "common.cuh"

namespace rcmvs {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace rcmvs

extern "C" int rcmvs_version(void) { return 100; }
extern "C" const char* rcmvs_last_error(void) { return rcmvs::g_err; }

extern "C" int rcmvs_check_device(void) {
  int dev = 0;
  cudaDeviceProp p;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&p, dev) != cudaSuccess) {
    rcmvs::set_error("rcmvs_check_device: no CUDA device: %s", cudaGetErrorString(cudaGetLastError()));
    return RCMVS_ERR_CUDA;
  }
  if (p.major != 10) {
    rcmvs::set_error("rcmvs_check_device: librcmvs is built for sm_100a only, device is sm_%d%d", p.major, p.minor);
    return RCMVS_ERR_ARCH;
  }
  return RCMVS_OK;
}
