// lip_api.cu — error reporting and device queries of the C ABI.
#include <atomic>

#include "lip_common.cuh"

namespace lip {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

const char* get_error() { return g_err; }

static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
long long launches() { return g_launches.load(std::memory_order_relaxed); }

}  // namespace lip

extern "C" {

const char* lip_last_error(void) { return lip::get_error(); }

int lip_version(void) { return 100; }

int64_t lip_launch_count(void) { return (int64_t)lip::launches(); }

int lip_device_is_sm100(void) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    lip::set_error("cudaGetDevice: %s", cudaGetErrorString(e));
    return LIP_ERR_CUDA;
  }
  int major = 0, minor = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  return (major == 10 && minor == 0) ? 1 : 0;
}

}  // extern "C"
