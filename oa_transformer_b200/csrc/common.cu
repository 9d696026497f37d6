// Library-level entry points: version, error string, device check.
#include <stdlib.h>
#include <string.h>

#include "oat_host.h"

namespace oat {

char* last_error_buffer() {
  static thread_local char buf[512] = {0};
  return buf;
}

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(last_error_buffer(), 512, fmt, ap);
  va_end(ap);
  return code;
}

bool pdl_enabled() {
  static const bool on = [] {
    // measured on the benchmark step: 59.4-60.3 ms with, 58.7-59.5 ms without - no gain (the persistent kernels own every
    // SM until their last CTA exits), so programmatic dependent launch stays opt-in
    const char* e = getenv("OAT_PDL");
    return e != nullptr && atoi(e) != 0;
  }();
  return on;
}

int num_sms() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

}  // namespace oat

extern "C" int oat_version(void) { return 100; }

extern "C" const char* oat_last_error(void) { return oat::last_error_buffer(); }

extern "C" int oat_device_check(void) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return oat::set_error(OAT_ERR_CUDA, "cudaGetDevice: %s", cudaGetErrorString(e));
  int major = 0;
  e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (e != cudaSuccess) return oat::set_error(OAT_ERR_CUDA, "cudaDeviceGetAttribute: %s", cudaGetErrorString(e));
  if (major != 10)
    return oat::set_error(OAT_ERR_ARCH, "liboat is built for sm_100a only; device %d has compute capability %d.x",
                          dev, major);
  return OAT_OK;
}
