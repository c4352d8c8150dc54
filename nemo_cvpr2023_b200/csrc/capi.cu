// Error plumbing + device queries of the C ABI.
#include <stdarg.h>
#include <string.h>
#include "common.cuh"

namespace nb {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace nb

extern "C" const char* nemo_b200_last_error(void) { return nb::g_err; }
extern "C" int nemo_b200_version(void) { return 100; }  // 0.1.0
extern "C" int nemo_b200_sm_arch(void) {
  int dev = 0, major = 0, minor = 0;
  NB_CUDA(cudaGetDevice(&dev));
  NB_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  NB_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
  return major * 10 + minor;
}
