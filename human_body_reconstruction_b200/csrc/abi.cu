// Error plumbing and small queries of the C ABI (include/hbr.h).
#include "common.cuh"
#include <string.h>

namespace hbr {

char* err_buf() {
  static thread_local char buf[512] = {0};
  return buf;
}

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(err_buf(), 512, fmt, ap);
  va_end(ap);
  return code;
}

int sm_count() {
  static thread_local int cached_dev = -1, cached = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev != cached_dev) {
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, dev) != cudaSuccess) return 148;
    cached = p.multiProcessorCount;
    cached_dev = dev;
  }
  return cached;
}

}  // namespace hbr

extern "C" int hbr_abi_version(void) { return HBR_ABI_VERSION; }
extern "C" const char* hbr_last_error(void) { return hbr::err_buf(); }
