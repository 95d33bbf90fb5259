// libcm2: version / error plumbing.
#include "common.cuh"
#include <stdlib.h>
#include <string.h>

namespace cm2 {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
bool pdl_enabled() {
  static const int on = getenv("CM2_PDL") ? atoi(getenv("CM2_PDL")) : 1;
  return on != 0;
}
}  // namespace cm2

extern "C" int cm2_version(void) { return CM2_VERSION; }
extern "C" const char* cm2_last_error(void) { return cm2::g_err; }

extern "C" int cm2_device_info(int* sm_count, int* cc_major, int* cc_minor) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    cudaGetLastError();
    if (sm_count) *sm_count = 0;
    if (cc_major) *cc_major = 0;
    if (cc_minor) *cc_minor = 0;
    cm2::set_error("no CUDA device");
    return CM2_ERR_CUDA;
  }
  cudaDeviceProp p;
  if (cudaGetDeviceProperties(&p, dev) != cudaSuccess) {
    cudaGetLastError();
    cm2::set_error("cudaGetDeviceProperties failed");
    return CM2_ERR_CUDA;
  }
  if (sm_count) *sm_count = p.multiProcessorCount;
  if (cc_major) *cc_major = p.major;
  if (cc_minor) *cc_minor = p.minor;
  return CM2_OK;
}
