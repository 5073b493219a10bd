// Hardware probe for the tcgen05 operand layouts the tensor-core engine relies on (diagnostic entry point
// sinc_tc_probe, used by tests/test_gpu_tc_probe.py).  The host supplies raw shared-memory images, raw
// matrix descriptors (without the start address) and the instruction descriptor; the kernel runs `ksteps`
// MMAs into one accumulator and returns D[128][N].  This is how the in-place Toeplitz descriptors
// (overlapping core matrices, LBO = 16 B) and the TMEM A-operand layout were validated on B200.
#include <type_traits>

#include "sinc_b200_diag.h"
#include "sinc_internal.h"
#include "sinc_tc.cuh"

namespace sinc {

struct ProbeArgs {
  const uint8_t* imgA;   // smem image of A, or (a_from_tmem) row-major float [128][a_cols]
  const uint8_t* imgB;
  int bytesA, bytesB;
  unsigned long long descA, descB;  // start-address field zero
  uint32_t idesc;
  int ksteps;
  int a_step;  // bytes per k-step (smem A) or TMEM columns per k-step
  int b_step;  // bytes per k-step
  int a_from_tmem, a_cols;
  int N;
  float* D;    // [128][N]
  int* status; // 0 ok, 1 = mbarrier wait timed out
};

__global__ void __launch_bounds__(128, 1) tc_probe_kernel(const ProbeArgs a) {
  extern __shared__ __align__(1024) uint8_t psm[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_slot;
  uint8_t* sA = psm;
  uint8_t* sB = psm + ((a.a_from_tmem ? 0 : a.bytesA) + 1023) / 1024 * 1024;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (!a.a_from_tmem)
    for (int i = tid; i < a.bytesA / 4; i += 128) reinterpret_cast<uint32_t*>(sA)[i] = reinterpret_cast<const uint32_t*>(a.imgA)[i];
  for (int i = tid; i < a.bytesB / 4; i += 128) reinterpret_cast<uint32_t*>(sB)[i] = reinterpret_cast<const uint32_t*>(a.imgB)[i];
  if (tid == 0) {
    tc::mbar_init(&bar, 1);
    tc::mbar_init_fence();
  }
  if (warp == 0) {
    tc::tmem_alloc(&tmem_base_slot, 512);
    tc::tmem_relinquish();
  }
  tc::fence_async_smem();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = tmem_base_slot;
  const uint32_t lane_base = (uint32_t)(32 * warp) << 16;
  const uint32_t a_col0 = 256;
  if (a.a_from_tmem) {
    const float* row = reinterpret_cast<const float*>(a.imgA) + (size_t)tid * a.a_cols;
    for (int c = 0; c < a.a_cols; c += 8) {
      uint32_t v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = __float_as_uint(row[c + i]);
      tc::tmem_st8(tmem + lane_base + a_col0 + c, v);
    }
    tc::tmem_st_wait();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
  }
  if (tid == 0) {
    const bool f16 = ((a.idesc >> 7) & 7u) < 2u;  // A format 0/1 = f16/bf16 -> kind::f16, 2 = tf32 -> kind::tf32
    for (int s = 0; s < a.ksteps; ++s) {
      const uint64_t bd = tc::desc_at(a.descB, tc::smem_addr(sB) + s * a.b_step);
      if (a.a_from_tmem) {
        if (f16) tc::mma_f16_ts(tmem, tmem + a_col0 + s * a.a_step, bd, a.idesc, s > 0);
        else tc::mma_tf32_ts(tmem, tmem + a_col0 + s * a.a_step, bd, a.idesc, s > 0);
      } else {
        const uint64_t ad = tc::desc_at(a.descA, tc::smem_addr(sA) + s * a.a_step);
        if (f16) tc::mma_f16_ss(tmem, ad, bd, a.idesc, s > 0);
        else tc::mma_tf32_ss(tmem, ad, bd, a.idesc, s > 0);
      }
    }
    tc::mma_commit(&bar);
  }
  const bool ok = tc::mbar_wait(&bar, 0);
  tc::fence_after_sync();
  if (!ok && tid == 0) *a.status = 1;
  if (ok) {
    for (int n0 = 0; n0 < a.N; n0 += 8) {
      uint32_t v[8];
      tc::tmem_ld8(tmem + lane_base + n0, v);
      tc::tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 8; ++i) a.D[(size_t)tid * a.N + n0 + i] = __uint_as_float(v[i]);
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

}  // namespace sinc

extern "C" int sinc_tc_probe(const void* imgA, int bytesA, const void* imgB, int bytesB, unsigned long long descA,
                             unsigned long long descB, unsigned int idesc, int ksteps, int a_step, int b_step,
                             int a_from_tmem, int a_cols, int N, float* D_out, int* status, void* stream) {
  using namespace sinc;
  if (!imgA || !imgB || !D_out || !status) return SINC_E_NULL;
  if (N < 8 || N > 256 || (N % 8) || ksteps < 1 || bytesA < 0 || bytesB < 0 || (bytesA % 4) || (bytesB % 4)) return SINC_E_SHAPE;
  if (a_from_tmem && (a_cols % 8 || a_cols > 256)) return SINC_E_SHAPE;
  DeviceInfo di;
  int r = get_device_info(&di);
  if (r) return r;
  if (di.cc_major != 10) return SINC_E_DEVICE;
  ProbeArgs a;
  a.imgA = (const uint8_t*)imgA;
  a.imgB = (const uint8_t*)imgB;
  a.bytesA = bytesA;
  a.bytesB = bytesB;
  a.descA = descA;
  a.descB = descB;
  a.idesc = idesc;
  a.ksteps = ksteps;
  a.a_step = a_step;
  a.b_step = b_step;
  a.a_from_tmem = a_from_tmem;
  a.a_cols = a_cols;
  a.N = N;
  a.D = D_out;
  a.status = status;
  const size_t smem = ((a_from_tmem ? 0 : bytesA) + 1023) / 1024 * 1024 + (bytesB + 1023) / 1024 * 1024 + 1024;
  if (smem > (size_t)di.max_smem_optin) return SINC_E_UNSUPPORTED;
  cudaError_t e = ensure_dynamic_smem(reinterpret_cast<const void*>(tc_probe_kernel), (int)smem);
  if (e != cudaSuccess) return (int)e;
  tc_probe_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(a);
  count_launch();
  return (int)cudaGetLastError();
}

// ---- MMA issue-rate probe (diagnostic): cycles per tcgen05.mma kind::f16, M = 128, for a given N, number of
// accumulators rotated through, and A operand source -- the numbers DESIGN.md quotes for "what an N = 32 MMA costs".
namespace sinc {

struct RateArgs {
  int N, nacc, count, a_from_tmem;
  int pattern;        // 0: one B descriptor; 1: forward-kernel sequence (17 K-steps x 3 products, B start += 32 bytes per
                      // K-step, hi/lo operands); 2: the same with every B start address 128-byte aligned
  int stress;         // what warps 1-3 do meanwhile: 0 nothing, 1 tcgen05.st into unused TMEM columns, 2 LDS.128 from the
                      // B operand's shared memory, 3 tcgen05.ld of the accumulator columns
  int nonzero;        // operands filled with a non-trivial bit pattern (switching activity) instead of zeros
  long long* cycles;  // [1]
  int* status;
};

__global__ void __launch_bounds__(128, 1) tc_mma_rate_kernel(const RateArgs a) {
  extern __shared__ __align__(1024) uint8_t rsm[];  // B: 256 rows x 32 bytes (zeros), A (SS form): 128 rows x 32 bytes
  __shared__ uint64_t bar, bar2;   // bar2: target of the extra commits of pattern 3 (never waited on)
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 16384 / 4; i += 128)
    reinterpret_cast<uint32_t*>(rsm)[i] = a.nonzero ? 0x3c003c00u ^ (uint32_t)(i * 0x9e3779b1u & 0x03ff03ffu) : 0u;
  if (tid == 0) {
    tc::mbar_init(&bar, 1);
    tc::mbar_init(&bar2, 1000000);
    tc::mbar_init_fence();
  }
  if (warp == 0) {
    tc::tmem_alloc(&slot, 512);
    tc::tmem_relinquish();
  }
  tc::fence_async_smem();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tm = slot;
  {  // zero the A columns [448, 512) of every lane
    uint32_t z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (a.nonzero)
      for (int j = 0; j < 8; ++j) z[j] = 0x3c003c00u ^ (uint32_t)((tid * 8 + j) * 0x85ebca6bu & 0x03ff03ffu);
    for (int c = 0; c < 64; c += 8) tc::tmem_st8(tm + ((uint32_t)(32 * warp) << 16) + 448 + c, z);
    tc::tmem_st_wait();
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  if (warp == 0) {
    const uint32_t tmu = tc::uniform((int)tm);
    const int N = tc::uniform(a.N), nacc = tc::uniform(a.nacc), count = tc::uniform(a.count);
    const uint32_t idesc = tc::make_idesc_f16(128, N, tc::kFmtBF16, tc::kFmtBF16);
    const uint64_t bd = tc::desc_at(tc::make_smem_desc(16, 128), tc::uniform((int)tc::smem_addr(rsm)));
    const uint64_t ad = tc::desc_at(tc::make_smem_desc(16, 128), tc::uniform((int)tc::smem_addr(rsm + 8192)));
    long long t0 = 0;
    const uint32_t b0 = tc::uniform((int)tc::smem_addr(rsm));  // (warp-wide shuffle: outside the elected-lane block)
    if (tc::elect_one()) {
      t0 = clock64();
      int acc = 0;
      if (a.pattern == 3) {
        // the forward sequence with a tcgen05.commit (to a barrier nobody waits on) every KC K-steps = 3*KC MMAs:
        // the kernels commit to release ring slots -- does that cost the tensor pipe anything?
        const uint64_t bd0 = tc::make_smem_desc(16, 128);
        auto body = [&](auto kc) {
          constexpr int KC = decltype(kc)::value;
          for (int i = 0; i < count; i += 51) {
#pragma unroll
            for (int q = 0; q < 17; ++q) {
              const uint64_t dh = tc::desc_at(bd0, b0 + 32 * q), dl = tc::desc_at(bd0, b0 + 4096 + 32 * q);
              const uint32_t ah = tmu + 448 + 8 * (q & 3), al = ah + 32;
              tc::mma_f16_ts(tmu, ah, dh, idesc, 1);
              tc::mma_f16_ts(tmu, ah, dl, idesc, 1);
              tc::mma_f16_ts(tmu, al, dh, idesc, 1);
              if (q % KC == KC - 1) tc::mma_commit(&bar2);
            }
          }
        };
        if (nacc == 1) body(std::integral_constant<int, 1>{});
        else if (nacc == 2) body(std::integral_constant<int, 2>{});
        else if (nacc == 4) body(std::integral_constant<int, 4>{});
        else body(std::integral_constant<int, 8>{});
      } else if (a.pattern) {
        const uint64_t bd0 = tc::make_smem_desc(16, 128);
        const uint32_t step = a.pattern == 2 ? 128u : 32u;
        for (int i = 0; i < count; i += 51) {
          for (int q = 0; q < 17; ++q) {
            const uint64_t dh = tc::desc_at(bd0, b0 + step * q), dl = tc::desc_at(bd0, b0 + 4096 + step * q);
            const uint32_t ah = tmu + 448 + 8 * (q & 3), al = ah + 32;
            // nacc = 1: one accumulator; nacc = 2: consecutive MMAs alternate between two accumulators (the
            // grad-filter issue order); nacc = 3: runs of three MMAs per accumulator, two accumulators
            const uint32_t d1 = tmu + (uint32_t)N;
            tc::mma_f16_ts(nacc == 3 && (q & 1) ? d1 : tmu, ah, dh, idesc, 1);
            tc::mma_f16_ts(nacc == 2 ? d1 : (nacc == 3 && (q & 1) ? d1 : tmu), ah, dl, idesc, 1);
            tc::mma_f16_ts(nacc == 3 && (q & 1) ? d1 : tmu, al, dh, idesc, 1);
          }
        }
      } else
      for (int i = 0; i < count; ++i) {
        const uint32_t d = tmu + (uint32_t)(acc * N);
        if (a.a_from_tmem)
          tc::mma_f16_ts(d, tmu + 448 + 8 * (i & 7), bd, idesc, 1);
        else
          tc::mma_f16_ss(d, ad, bd, idesc, 1);
        acc = acc + 1 == nacc ? 0 : acc + 1;
      }
      tc::mma_commit(&bar);
    }
    __syncwarp();
    const bool ok = tc::mbar_wait(&bar, 0);
    if (tc::elect_one()) {
      if (blockIdx.x == 0) a.cycles[0] = clock64() - t0;
      if (!ok) *a.status = 1;
      *reinterpret_cast<volatile uint32_t*>(&slot) = 0xffffffffu;  // stops the background traffic of warps 1-3
    }
    __syncwarp();
  }
  else if (a.stress) {
    // background traffic until warp 0 reports (the flag is the cycle count it writes at the end)
    volatile uint32_t* done = &slot;  // reused as the 'finished' flag: warp 0 sets it to 0xffffffff
    uint32_t z[8] = {1, 2, 3, 4, 5, 6, 7, 8};
    uint32_t sink = 0;
    const uint32_t lb = ((uint32_t)(32 * warp) << 16);
    while (*done != 0xffffffffu) {
      for (int r = 0; r < 16; ++r) {
        if (a.stress == 1) {
          tc::tmem_st8(tm + lb + 384 + 8 * (r & 7), z);
        } else if (a.stress == 2) {
          uint32_t v0, v1, v2, v3;
          asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                       : "=r"(v0), "=r"(v1), "=r"(v2), "=r"(v3)
                       : "r"(tc::smem_addr(rsm + 16 * ((tid + 37 * r) & 511))));
          sink += v0 + v3;
        } else {
          uint32_t v[8];
          tc::tmem_ld8(tm + lb + 8 * (r & 7), v);
          tc::tmem_ld_wait();
          sink += v[0];
        }
      }
      if (a.stress == 1) tc::tmem_st_wait();
    }
    if (sink == 0x12345678u) a.cycles[0] = 0;
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tm, 512);
}

// ---- register fragment of tcgen05.ld 16x256b.x8 (the grad-input epilogue's load shape) ------------------------------
__global__ void __launch_bounds__(32, 1) ld16x256_probe_kernel(uint32_t* out) {
  __shared__ uint32_t slot;
  tc::tmem_alloc(&slot, 64);
  tc::tmem_relinquish();
  tc::fence_before_sync();
  __syncwarp();
  tc::fence_after_sync();
  const uint32_t tm = slot;
  const int lane = threadIdx.x;
  for (int c0 = 0; c0 < 64; c0 += 8) {
    uint32_t v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = 256u * lane + (uint32_t)(c0 + e);
    tc::tmem_st8(tm + c0, v);   // 32x32b: thread = TMEM lane
  }
  tc::tmem_st_wait();
  tc::fence_before_sync();
  __syncwarp();
  tc::fence_after_sync();
  uint32_t lo[32], hi[32];
  tc::tmem_ld16x256_x8(tm, lo);                  // lanes 0..15
  tc::tmem_ld16x256_x8(tm + (16u << 16), hi);    // lanes 16..31
  tc::tmem_ld_wait();
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    out[lane * 64 + i] = lo[i];
    out[lane * 64 + 32 + i] = hi[i];
  }
  tc::fence_before_sync();
  __syncwarp();
  tc::tmem_dealloc(tm, 64);
}

// ---- what a wait on an ALREADY COMPLETED mbarrier phase costs (cycles per wait, one thread, dependent chain) ----
// mode 0: mbarrier.try_wait with the suspend-time hint (tc::mbar_try_wait), 1: try_wait without a hint,
// 2: mbarrier.test_wait (non-blocking), 3: tc::mbar_wait (the kernels' bounded wait: test_wait, then try_wait polls)
__global__ void __launch_bounds__(32, 1) mbar_cost_kernel(int mode, int count, long long* cycles) {
  __shared__ uint64_t bar;
  if (threadIdx.x == 0) {
    tc::mbar_init(&bar, 1);
    tc::mbar_init_fence();
    tc::mbar_arrive(&bar);  // phase 0 complete: waiting for parity 0 succeeds at once
  }
  __syncwarp();
  if (threadIdx.x == 0) {
    const uint32_t addr = tc::smem_addr(&bar);
    uint32_t acc = 0;
    const long long t0 = clock64();
    for (int i = 0; i < count; ++i) {
      uint32_t ok = 0;
      if (mode == 0) {
        ok = tc::mbar_try_wait(&bar, acc & 0u) ? 1u : 0u;
      } else if (mode == 1) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(ok) : "r"(addr), "r"(acc & 0u) : "memory");
      } else if (mode == 2) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(ok) : "r"(addr), "r"(acc & 0u) : "memory");
      } else {
        ok = tc::mbar_wait(&bar, acc & 0u) ? 1u : 0u;
      }
      acc += ok;  // the next wait depends on this one
    }
    cycles[0] = clock64() - t0;
    cycles[1] = acc;
  }
}

}  // namespace sinc

extern "C" int sinc_tc_ld16x256_probe(uint32_t* out2048, void* stream) {
  if (!out2048) return SINC_E_NULL;
  sinc::ld16x256_probe_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(out2048);
  sinc::count_launch();
  return (int)cudaGetLastError();
}

extern "C" int sinc_mbar_wait_cost(int mode, int count, long long* cycles_out2, void* stream) {
  if (!cycles_out2 || count < 1 || mode < 0 || mode > 3) return SINC_E_SHAPE;
  sinc::mbar_cost_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(mode, count, cycles_out2);
  sinc::count_launch();
  return (int)cudaGetLastError();
}

extern "C" int sinc_tc_mma_rate(int N, int nacc, int count, int a_from_tmem, long long* cycles_out, int* status,
                                void* stream) {
  using namespace sinc;
  if (!cycles_out || !status) return SINC_E_NULL;
  if (N < 16 || N > 256 || (N % 16) || nacc < 1 || nacc * N > 448 || count < 1) return SINC_E_SHAPE;
  DeviceInfo di;
  int r = get_device_info(&di);
  if (r) return r;
  if (di.cc_major != 10) return SINC_E_DEVICE;
  // a_from_tmem: bit 0 = A operand in TMEM, bit 1 = non-zero operands, bits 2..3 = pattern, bits 4..5 = background
  // traffic of the other warps (RateArgs::stress), bits 8..15 = CTAs (0 = one)
  const int grid = ((a_from_tmem >> 8) & 0xff) ? ((a_from_tmem >> 8) & 0xff) : 1;
  RateArgs a{N, nacc, count, a_from_tmem & 1, (a_from_tmem >> 2) & 3, (a_from_tmem >> 4) & 3, (a_from_tmem >> 1) & 1,
             cycles_out, status};
  const size_t smem = 16384 + 1024;
  cudaError_t e = ensure_dynamic_smem(reinterpret_cast<const void*>(tc_mma_rate_kernel), (int)smem);
  if (e != cudaSuccess) return (int)e;
  tc_mma_rate_kernel<<<grid, 128, smem, (cudaStream_t)stream>>>(a);
  count_launch();
  return (int)cudaGetLastError();
}
