// Error plumbing, version and device queries for libyolact_b200.so.
#include "common.cuh"
#include <string.h>

namespace yb {
static thread_local char g_err[512] = "";
std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace yb

extern "C" int yb_version(void) { return YB_VERSION; }
extern "C" const char* yb_last_error(void) { return yb::g_err; }
extern "C" uint64_t yb_launch_count(void) { return yb::g_launches.load(); }

extern "C" int yb_device_info(int* sm_count, int* cc_major, int* cc_minor) {
  int dev = 0;
  YB_CHECK_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  YB_CHECK_CUDA(cudaGetDeviceProperties(&prop, dev));
  if (sm_count) *sm_count = prop.multiProcessorCount;
  if (cc_major) *cc_major = prop.major;
  if (cc_minor) *cc_minor = prop.minor;
  return YB_OK;
}
