// The synthesis chain rule: reduced grad-filter G[c, :] -> d low_hz_[c], d band_hz_[c], in fp64, closed form
// of the autograd graph behind agents/perception.py:89-115.  Shared by the stand-alone finalize_kernel (sinc_fma.cu) and by
// the tail of the 16-bit grad-filter kernel (sinc_tc16_bwd_filter.cu), which finalises inside the kernel that produced the
// partial tiles.  Block-wide calls: every one of the NT threads of the CTA must take part.
#pragma once
#include "sinc_internal.h"
#include "sinc_synth.cuh"

namespace sinc {

__device__ __forceinline__ double sgn(double v) { return (double)((v > 0.0) - (v < 0.0)); }

// Two block-wide sums behind one pair of barriers (scratch: 2 x 32 doubles); fixed order: bitwise reproducible.
template <int NT>
__device__ __forceinline__ void block_sum2(double& a, double& b, double* scratch) {
  a = warp_sum(a);
  b = warp_sum(b);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) {
    scratch[warp] = a;
    scratch[32 + warp] = b;
  }
  __syncthreads();
  double sa = 0.0, sb = 0.0;
#pragma unroll 4
  for (int w = 0; w < NT / 32; ++w) {
    sa += scratch[w];
    sb += scratch[32 + w];
  }
  a = sa;
  b = sb;
}

// Everything of the chain rule that does not depend on G: the un-normalised taps us[K] and the cosine factors chi[H], clo[H].
// Called by `nthr` threads numbered t = 0..nthr-1 (any subset of the CTA); the caller synchronises afterwards.
__device__ __forceinline__ void finalize_factors(const Bank& bank, const ChanCoef& cc, int t, int nthr, double* us,
                                                 double* chi, double* clo) {
  const int K = bank.K, H = (K - 1) / 2;
  for (int j = t; j < H; j += nthr) {
    const double nj = (double)bank.n[j];
    const double left = left_tap(cc, nj);
    us[j] = left * (double)bank.window[j];
    us[K - 1 - j] = left * (double)bank.window[K - 1 - j];
    const double d = nj * 0.5 + 1e-8;
    chi[j] = nj * cos(cc.f_hi * nj) / d;
    clo[j] = nj * cos(cc.f_lo * nj) / d;
  }
  if (t == nthr - 1) us[H] = 2.0 * (cc.f_hi - cc.f_lo) * (double)bank.window[H];
}

// Gs[K] (reduced grad-filter of filter c), us / chi / clo from finalize_factors, all in shared memory and visible to the
// whole CTA (the caller has synchronised) -> grad_low[c], grad_band[c].  scratch: 64 doubles.
template <int NT>
__device__ __forceinline__ void finalize_chain_rule(const Bank& bank, const ChanCoef& cc, int c, const double* Gs,
                                                    const double* us, const double* chi, const double* clo,
                                                    double* scratch, float* grad_low, float* grad_band) {
  const int K = bank.K, H = (K - 1) / 2;
  const int tid = threadIdx.x;
  // the two parameters whose signs close the chain rule: loaded now, used by thread 0 at the very end
  const float p_low = tid == 0 ? bank.low_hz[c] : 0.f, p_band = tid == 0 ? bank.band_hz[c] : 0.f;
  double sabs = 0.0, dot = 0.0;
  for (int k = tid; k < K; k += NT) {
    sabs += fabs(us[k]);
    dot += Gs[k] * us[k];
  }
  block_sum2<NT>(sabs, dot, scratch);
  const double S = sabs + 1e-8;

  // P_k = dL/du_k * window_k,  dL/du_k = G_k/S - sign(u_k) * dot / S^2   (one fp64 reciprocal instead of two inline
  // divisions per call: this code runs once per launch and is paid for in instruction fetch)
  const double rS = 1.0 / S, c2 = dot * rS * rS;
  auto P = [&](int k) { return (Gs[k] * rS - sgn(us[k]) * c2) * (double)bank.window[k]; };
  double dhi = 0.0, dlo = 0.0;
  for (int j = tid; j < H; j += NT) {
    const double pl = P(j) + P(K - 1 - j);
    dhi += pl * chi[j];
    dlo -= pl * clo[j];
  }
  block_sum2<NT>(dhi, dlo, scratch);
  if (tid == 0) {
    const double pc = P(H);
    const double d_fhi = dhi + 2.0 * pc;
    const double d_flo = dlo - 2.0 * pc;
    // two true divisions (one thread, once): with a reciprocal the compiler contracts d_flo * (1/sr) + d_high into an FMA,
    // and the exact cancellation d_low = (-2 pc + 2 pc) / sr of a one-tap filter (K = 1) turns into a rounding residue
    const double d_high = d_fhi / bank.sr;
    const double d_low_direct = d_flo / bank.sr;
    const double un = cc.unclamped ? 1.0 : 0.0;
    grad_band[c] = (float)(un * d_high * sgn((double)p_band));
    grad_low[c] = (float)((d_low_direct + un * d_high) * sgn((double)p_low));
  }
}

}  // namespace sinc
