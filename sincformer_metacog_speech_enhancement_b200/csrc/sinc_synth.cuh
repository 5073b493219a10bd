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
  const double inv_sr = 1.0 / bk.sr;                      // one fp64 division (independent of the parameter loads)
  cc.f_lo = low * inv_sr;                                 // :95   (differs from low / sr in the last bit of a double)
  cc.f_hi = high * inv_sr;                                // :96
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
    return __fdividef(2.0f * cosf(fs * nj) * sinf(fd * nj), nj * 0.5f + 1e-8f);  // 2 ulp quotient: inside the 3e-7 target
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
  // one fp64 reciprocal instead of a division per tap (the quotient differs from u / S in the last bit of a double,
  // far below the fp32 rounding of the result); the prologue of every convolution kernel runs this function
  const double rS = 1.0 / S;
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int j = r * 32 + lane;
    if (j < H) {
      emit(j, (double)ul[r] * rS);
      emit(K - 1 - j, (double)ur[r] * rS);
    }
  }
  for (int j = R * 32 + lane; j < H; j += 32) {
    const float left = left32(j);
    emit(j, (double)(left * bk.window[j]) * rS);
    emit(K - 1 - j, (double)(left * bk.window[K - 1 - j]) * rS);
  }
  if (lane == 0) emit(H, (double)uc * rS);
}

// ---- filter table in the workspace -----------------------------------------------------------------------------
// The grad-filter kernel (16-bit tcgen05 engine) has idle warps at its start; they synthesise the filters ONCE and
// leave them in the workspace, so that the grad-input kernel -- launched later on the same stream -- does not spend
// its prologue synthesising all 64 filters in every CTA (ncu r04b: 14 % of that kernel).  The table is self-
// validating, because a workspace is caller memory that may be recycled between calls: a header holds an exact copy
// of everything the synthesis depends on (C, K, sample rate, minima, low_hz_, band_hz_, window, n_) and a checksum
// per row; the consumer compares the header with ITS filter bank word by word and re-computes the row checksums
// while it copies the rows, and synthesises the filters itself on any mismatch.
constexpr int kTabHdrWords = 8 + 64 + 64 + 256 + 128;       // 520 32-bit words
constexpr int kTabRowsumOff = 2112;                         // bytes: 64 x uint64 row checksums
constexpr int kTabDataOff = 4096;                           // bytes: float W[64][256], zero outside [0,C) x [0,K)
static_assert(kTabDataOff + 64 * 256 * 4 == kFilterTableBytes, "workspace_layout reserves the table");

__device__ __forceinline__ bool tab_shape_ok(const Bank& bk) { return bk.C <= 64 && bk.K <= 256; }

// word i of the header for this filter bank
__device__ __forceinline__ uint32_t tab_header_word(const Bank& bk, int i) {
  const int H = (bk.K - 1) / 2;
  if (i < 8) {
    if (i == 0) return (uint32_t)bk.C;
    if (i == 1) return (uint32_t)bk.K;
    const double d = i < 4 ? bk.sr : i < 6 ? bk.min_low : bk.min_band;
    const unsigned long long u = (unsigned long long)__double_as_longlong(d);
    return (i & 1) ? (uint32_t)(u >> 32) : (uint32_t)u;
  }
  i -= 8;
  if (i < 64) return i < bk.C ? __float_as_uint(bk.low_hz[i]) : 0u;
  i -= 64;
  if (i < 64) return i < bk.C ? __float_as_uint(bk.band_hz[i]) : 0u;
  i -= 64;
  if (i < 256) return i < bk.K ? __float_as_uint(bk.window[i]) : 0u;
  i -= 256;
  return i < H ? __float_as_uint(bk.n[i]) : 0u;
}

__device__ __forceinline__ unsigned long long tab_mix(float w, int k) {
  return (unsigned long long)__float_as_uint(w) * (unsigned long long)(2 * k + 1);
}
__device__ __forceinline__ unsigned long long warp_sum_u64(unsigned long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Whole-warp call: row c of the table (+ its checksum).
__device__ __forceinline__ void tab_write_row(const Bank& bk, uint8_t* tab, int c, int lane) {
  float* row = reinterpret_cast<float*>(tab + kTabDataOff) + c * 256;
  for (int k = lane; k < 256; k += 32)
    if (c >= bk.C || k >= bk.K) row[k] = 0.f;
  unsigned long long sum = 0ull;
  if (c < bk.C)
    synth_filter_warp(bk, c, lane, [&](int k, double w) {
      const float wf = (float)w;
      row[k] = wf;
      sum += tab_mix(wf, k);
    });
  sum = warp_sum_u64(sum);
  if (lane == 0) reinterpret_cast<unsigned long long*>(tab + kTabRowsumOff)[c] = sum;
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
