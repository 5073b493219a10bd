// Host-side support shared by libsinc_b200.so and the diagnostics library: launch accounting and the immutable
// per-device caches (device attributes, dynamic shared-memory opt-in per kernel).  No mutable settings live here.
#include <atomic>
#include <map>
#include <mutex>
#include <utility>

#include "sinc_internal.h"

namespace sinc {

static std::atomic<uint64_t> g_launches{0};

void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }
uint64_t launch_count() { return g_launches.load(std::memory_order_relaxed); }

cudaError_t ensure_dynamic_smem(const void* kernel, int bytes) {
  static std::mutex mu;
  static std::map<std::pair<const void*, int>, int> done;  // (kernel, device) -> bytes already granted
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  std::lock_guard<std::mutex> lock(mu);
  auto key = std::make_pair(kernel, dev);
  auto it = done.find(key);
  if (it != done.end() && it->second >= bytes) return cudaSuccess;
  e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess) done[key] = bytes;
  return e;
}

int get_device_info(DeviceInfo* out) {
  static std::mutex mu;
  static DeviceInfo cache[64];
  static bool have[64] = {false};
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return (int)e;
  if (dev < 0 || dev >= 64) return SINC_E_DEVICE;
  std::lock_guard<std::mutex> lock(mu);
  if (!have[dev]) {
    DeviceInfo di;
    if ((e = cudaDeviceGetAttribute(&di.sm_count, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return (int)e;
    if ((e = cudaDeviceGetAttribute(&di.cc_major, cudaDevAttrComputeCapabilityMajor, dev)) != cudaSuccess) return (int)e;
    if ((e = cudaDeviceGetAttribute(&di.cc_minor, cudaDevAttrComputeCapabilityMinor, dev)) != cudaSuccess) return (int)e;
    if ((e = cudaDeviceGetAttribute(&di.max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev)) != cudaSuccess)
      return (int)e;
    cache[dev] = di;
    have[dev] = true;
  }
  *out = cache[dev];
  return 0;
}

}  // namespace sinc
