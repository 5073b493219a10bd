// Filter synthesis, device side: the arithmetic of SincConv1d.forward before F.conv1d
// (reference agents/perception.py:89-115), evaluated in fp64 on the fp32 *values* of the module's
// parameters and buffers, one warp per filter.  Every convolution kernel calls this in its prologue
// (no separate synthesis launch, no filter tensor in HBM).
//
// Accuracy target: stay within ~3e-7 of the reference evaluated in fp64, so that OUR distance from the fp32
// reference is the reference's own round-off (2e-6 .. 2e-5) and not the sum of two independent fp32 errors
// (SURVEY.md section 7 step 3, H4).  See synth_filter_warp for how that is done without fp64 transcendentals.
#pragma once
#include "sinc_internal.h"

namespace sinc {

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Per-filter scalars (perception.py:89-96).
struct ChanCoef {
  double f_lo;
  double f_hi;
  bool unclamped;  // gradient of torch.clamp(max=) passes where input <= max
};

__device__ __forceinline__ ChanCoef chan_coef(const Bank& bk, int c) {
  const double lo_p = (double)bk.low_hz[c];
  const double bd_p = (double)bk.band_hz[c];
  const double low = bk.min_low + fabs(lo_p);             // :89
  const double pre = low + bk.min_band + fabs(bd_p);      // :91
  const double nyq = bk.sr / 2.0;                         // :92
  ChanCoef cc;
  cc.unclamped = pre <= nyq;
  const double high = pre > nyq ? nyq : pre;              // clamp(max=); NaN passes through
  cc.f_lo = low / bk.sr;                                  // :95
  cc.f_hi = high / bk.sr;                                 // :96
  return cc;
}

// One tap of the un-normalised left half (perception.py:99-101).
__device__ __forceinline__ double left_tap(const ChanCoef& cc, double nj) {
  return (sin(cc.f_hi * nj) - sin(cc.f_lo * nj)) / (nj * 0.5 + 1e-8);
}

// Whole-warp call.  Synthesises filter `c` and calls emit(k, w_k) (w_k as double) for every tap k in [0,K)
// exactly once (from varying lanes).
//
// Arithmetic: the per-filter scalars are fp64; each tap is evaluated in fp32 through the cancellation-free
// identity  sin(a) - sin(b) = 2 cos((a+b)/2) sin((a-b)/2)  and the L1 norm is accumulated in fp64.  The
// reference's own form sin(f_hi n) - sin(f_lo n) cancels 10-25x in fp32 (that is why it sits 2e-6 .. 2e-5 from
// its fp64 evaluation); the product form has no cancellation, so fp32 gives ~3e-7 of the fp64 value at a
// sixteenth of the latency of fp64 sin (ncu r01m: the fp64 prologue was 10 % of the tensor-core kernels).
template <class Emit>
__device__ __noinline__ void synth_filter_warp(const Bank& bk, int c, int lane, Emit emit,
                                               double* S_out = nullptr) {
  const int K = bk.K;
  const int H = (K - 1) / 2;
  const ChanCoef cc = chan_coef(bk, c);
  const float fs = (float)(0.5 * (cc.f_hi + cc.f_lo));  // (a+b)/2 = fs * n
  const float fd = (float)(0.5 * (cc.f_hi - cc.f_lo));  // (a-b)/2 = fd * n
  auto left32 = [&](int j) -> float {
    const float nj = bk.n[j];
    return 2.0f * cosf(fs * nj) * sinf(fd * nj) / (nj * 0.5f + 1e-8f);
  };
  constexpr int R = 4;  // taps cached in registers per lane: covers H <= 128 (K <= 257)
  float ul[R], ur[R];
  double sumabs = 0.0;
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int j = r * 32 + lane;
    float l = 0.f, rr = 0.f;
    if (j < H) {
      const float left = left32(j);
      l = left * bk.window[j];           // :105-111, tap j
      rr = left * bk.window[K - 1 - j];  // mirrored tap K-1-j (the window is NOT symmetric)
      sumabs += (double)fabsf(l) + (double)fabsf(rr);
    }
    ul[r] = l;
    ur[r] = rr;
  }
  for (int j = R * 32 + lane; j < H; j += 32) {  // long filters: second pass recomputes these
    const float left = left32(j);
    sumabs += (double)fabsf(left * bk.window[j]) + (double)fabsf(left * bk.window[K - 1 - j]);
  }
  const float uc = (float)(2.0 * (cc.f_hi - cc.f_lo)) * bk.window[H];  // :104
  if (lane == 0) sumabs += (double)fabsf(uc);
  const double S = warp_sum(sumabs) + 1e-8;  // :112
  if (S_out) *S_out = S;
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int j = r * 32 + lane;
    if (j < H) {
      emit(j, (double)ul[r] / S);
      emit(K - 1 - j, (double)ur[r] / S);
    }
  }
  for (int j = R * 32 + lane; j < H; j += 32) {
    const float left = left32(j);
    emit(j, (double)(left * bk.window[j]) / S);
    emit(K - 1 - j, (double)(left * bk.window[K - 1 - j]) / S);
  }
  if (lane == 0) emit(H, (double)uc / S);
}

// ---- small PTX helpers used by several kernels -------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
// Ampere-style async copies with zero fill: bytes beyond src_bytes are written as zero.
__device__ __forceinline__ void cp_async16(void* dst_smem, const void* src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(smem_u32(dst_smem)), "l"(src),
               "r"(src_bytes)
               : "memory");
}
__device__ __forceinline__ void cp_async4(void* dst_smem, const void* src, int src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" ::"r"(smem_u32(dst_smem)), "l"(src),
               "r"(src_bytes)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

}  // namespace sinc
