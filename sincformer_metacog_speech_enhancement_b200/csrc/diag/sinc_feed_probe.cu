// Diagnostic (libsinc_b200_diag.so): how fast can one B200 stream the (B,C,T) fp32 activation from HBM into shared
// memory with the access patterns the backward kernels use -- and with the alternatives?  No arithmetic, no consumers:
// 148 persistent CTAs, each walks a contiguous range of 128-sample stages of the 64-row activation and only waits
// for the data to land.  The achieved rate is the ceiling of the "feed" the tensor-core kernels can be given.
//   mode 0: 1-D bulk copies, one per row and stage (64 x 512 B), `depth` stages in flight   (grad-filter staging, r01)
//   mode 1: 1-D bulk copies of 1 KB rows (256-sample stages, 64 x 1 KB)
//   mode 2: 1-D bulk copies of 2 KB rows (512-sample stages)
//   mode 3: plain 16-byte loads (LDG.128, each warp 512 contiguous bytes of one row), 8 per thread in flight
#include "sinc_b200_diag.h"
#include "sinc_internal.h"
#include "sinc_tc.cuh"

namespace sinc {

struct FeedArgs {
  const float* g;
  int64_t B, T;
  int C, mode, depth;
  int64_t nstages;          // B * ceil(T / samples_per_stage)
  int64_t stages_per_row;
  float* sink;
};

__global__ void __launch_bounds__(256, 1) feed_probe_kernel(const FeedArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full[8];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int sps = a.mode == 0 ? 128 : a.mode == 1 ? 256 : a.mode == 2 ? 512 : 128;   // samples per stage
  const int depth = a.depth;
  const int64_t s0 = a.nstages * blockIdx.x / gridDim.x, s1 = a.nstages * (blockIdx.x + 1) / gridDim.x;
  if (tid == 0) {
    for (int i = 0; i < 8; ++i) tc::mbar_init(&full[i], 1);
    tc::mbar_init_fence();
  }
  __syncthreads();
  float acc = 0.f;
  if (a.mode <= 2) {
    const uint32_t stage_bytes = 64u * sps * 4u;
    auto issue = [&](int64_t s) {   // whole warp 0 calls; one lane issues
      const int slot = (int)((s - s0) % depth);
      const int64_t b = s / a.stages_per_row, st = s - b * a.stages_per_row;
      int64_t left = a.T - st * sps;
      const uint32_t bytes = (uint32_t)((left < sps ? left : sps) * 4);
      if (tc::elect_one()) {
        tc::mbar_expect_tx(&full[slot], 64u * bytes);
        const uint8_t* src = reinterpret_cast<const uint8_t*>(a.g) + ((b * a.C) * a.T + st * sps) * 4;
        uint32_t dst = tc::smem_addr(smem) + slot * stage_bytes;
        const uint32_t bar = tc::smem_addr(&full[slot]);
        for (int r = 0; r < 64; ++r) {
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                       "l"(src), "r"(bytes), "r"(bar)
                       : "memory");
          src += a.T * 4;
          dst += sps * 4;
        }
      }
      __syncwarp();
    };
    if (warp == 0) {
      for (int64_t s = s0; s < s1 && s < s0 + depth; ++s) issue(s);
      for (int64_t s = s0; s < s1; ++s) {
        const int slot = (int)((s - s0) % depth);
        tc::mbar_wait(&full[slot], (uint32_t)(((s - s0) / depth) & 1));
        acc += reinterpret_cast<const float*>(smem + slot * stage_bytes)[lane];   // touch the data
        __syncwarp();
        if (s + depth < s1) issue(s + depth);
      }
    }
  } else {
    // plain loads: warp w takes rows 8w .. 8w+7 of a 128-sample stage: 512 contiguous bytes per row = one LDG.128 per lane
    for (int64_t s = s0; s < s1; ++s) {
      const int64_t b = s / a.stages_per_row, st = s - b * a.stages_per_row;
      const float4* base = reinterpret_cast<const float4*>(a.g + (b * a.C + 8 * warp) * a.T + st * 128) + lane;
      float4 v[8];
#pragma unroll
      for (int r = 0; r < 8; ++r) v[r] = __ldg(base + r * (a.T / 4));
#pragma unroll
      for (int r = 0; r < 8; ++r) acc += v[r].x + v[r].w;
    }
  }
  if (acc == 1.2345f) a.sink[0] = acc;
}

}  // namespace sinc

extern "C" int sinc_feed_probe(const float* g, int64_t B, int32_t C, int64_t T, int mode, int depth, float* sink,
                               void* stream) {
  using namespace sinc;
  if (!g || !sink) return SINC_E_NULL;
  if (C != 64 || T % 512 != 0 || mode < 0 || mode > 3 || depth < 1 || depth > 8) return SINC_E_SHAPE;
  DeviceInfo di;
  int r = get_device_info(&di);
  if (r) return r;
  const int sps = mode == 0 ? 128 : mode == 1 ? 256 : mode == 2 ? 512 : 128;
  FeedArgs a{g, B, T, C, mode, depth, B * (T / sps), T / sps, sink};
  const size_t smem = mode <= 2 ? (size_t)depth * 64 * sps * 4 + 1024 : 1024;
  if (smem > (size_t)di.max_smem_optin) return SINC_E_SHAPE;
  cudaError_t e = ensure_dynamic_smem(reinterpret_cast<const void*>(feed_probe_kernel), (int)smem);
  if (e != cudaSuccess) return (int)e;
  feed_probe_kernel<<<di.sm_count, 256, smem, (cudaStream_t)stream>>>(a);
  count_launch();
  return (int)cudaGetLastError();
}
