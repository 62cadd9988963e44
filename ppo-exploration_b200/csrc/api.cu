// Library-level plumbing of libppx: thread-local error string, launch counter, device info.
#include <atomic>
#include "common.cuh"

namespace ppx {
namespace {
thread_local char g_err[512] = "";
std::atomic<uint64_t> g_launches{0};
}  // namespace

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }
int sm_count() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0)
      sms = 148;   // B200
  }
  return sms;
}
}  // namespace ppx

extern "C" const char* ppx_last_error(void) { return ppx::g_err; }
extern "C" int ppx_version(void) { return 100; }
extern "C" uint64_t ppx_launch_count(void) { return ppx::g_launches.load(std::memory_order_relaxed); }
extern "C" int ppx_device_info(int* sm_count, int* cc_major, int* cc_minor) {
  int dev = 0;
  PPX_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  PPX_CUDA(cudaGetDeviceProperties(&prop, dev));
  if (sm_count) *sm_count = prop.multiProcessorCount;
  if (cc_major) *cc_major = prop.major;
  if (cc_minor) *cc_minor = prop.minor;
  return PPX_OK;
}
