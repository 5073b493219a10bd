// Next row N1: GroupNorm(8, C) + GELU right after the SincConv output, fused.
//     reference: agents/perception.py:157 (self.sinc_norm = nn.GroupNorm(8, D // 4)), :233-234
//                x = self.sinc_norm(x); x = F.gelu(x)
// The (B,C,T) activation is the largest tensor of the whole Perception Agent (256 B per input sample).  Unfused,
// torch reads / writes it 5 times in forward (GroupNorm statistics, normalise, GELU) and 7 times in backward.
// Here: the statistics come out of the convolution epilogue (sinc_tc16_fwd.cu, per-tile partial sums) or from one
// read pass (gn_stats_kernel), one pass applies normalisation + affine + GELU (read y, write z); backward is two
// passes (partial sums of du and du*yhat per channel chunk; then dy).  All kernels are HBM-bound streaming kernels:
// 16-byte accesses, grid sized in multiples of the SM count by the caller, fixed-order fp64 reductions of the
// partial sums (bitwise reproducible).
//
// Layout: y, z, dz, dy are (B,C,T) contiguous; a group's C/G channels are contiguous, so a (b,g) slab is one
// contiguous run of (C/G)*T elements.  16-byte accesses when T % 4 == 0 and the tensors are 16-byte aligned (template
// flag V); any other length or alignment takes the same kernels with 4-byte accesses -- the reference's
// gelu(GroupNorm(.)) accepts any length, so the drop-in does too (at roughly a quarter of the bandwidth).
#include <cuda_bf16.h>

#include "sinc_internal.h"

namespace sinc {

constexpr int GN_THREADS = 256;
constexpr int GN_CHUNK = 16384;  // elements of one channel row per CTA (64 KB): the per-CTA prologue (fp64 reduction of the
                                 // group's partial sums) is amortised over 16 float4 per thread, 4 in flight

__device__ __forceinline__ float gelu_f(float u) { return 0.5f * u * (1.f + erff(u * 0.70710678118654752440f)); }
__device__ __forceinline__ float gelu_grad_f(float u) {
  const float cdf = 0.5f * (1.f + erff(u * 0.70710678118654752440f));
  const float pdf = 0.39894228040143267794f * __expf(-0.5f * u * u);
  return cdf + u * pdf;
}

// block-wide sum of two doubles (all threads get the result); scratch: 2 * 32 doubles
__device__ __forceinline__ void block_sum2(double& a, double& b, double* scratch) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
  }
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = blockDim.x >> 5;
  __syncthreads();
  if (l == 0) {
    scratch[w] = a;
    scratch[32 + w] = b;
  }
  __syncthreads();
  a = 0.0;
  b = 0.0;
  for (int i = 0; i < nw; ++i) {  // fixed order
    a += scratch[i];
    b += scratch[32 + i];
  }
}

// ---- K1: per-chunk partial sums of y and y^2 ------------------------------------------------------------------
// grid = (B*C, chunks_per_row); partial[(b*C + c) * chunks_per_row + chunk] = (sum, sum of squares)
template <bool V>
__global__ void __launch_bounds__(GN_THREADS) gn_stats_kernel(const float* __restrict__ y, int64_t T, int chunks_per_row,
                                                              float2* __restrict__ partial) {
  __shared__ double scratch[64];
  const int64_t row = blockIdx.x;
  const int chunk = blockIdx.y;
  const int64_t t0 = (int64_t)chunk * GN_CHUNK;
  const int64_t n = (T - t0) < GN_CHUNK ? (T - t0) : GN_CHUNK;
  float s = 0.f, q = 0.f;
  if constexpr (V) {
    const float4* p = reinterpret_cast<const float4*>(y + row * T + t0);
#pragma unroll 4
    for (int i = threadIdx.x; i < n / 4; i += GN_THREADS) {
      const float4 v = __ldg(p + i);
      s += (v.x + v.y) + (v.z + v.w);
      q += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
    }
  } else {
    const float* p = y + row * T + t0;
#pragma unroll 4
    for (int i = threadIdx.x; i < n; i += GN_THREADS) {
      const float v = __ldg(p + i);
      s += v;
      q = fmaf(v, v, q);
    }
  }
  double ds = s, dq = q;
  block_sum2(ds, dq, scratch);
  if (threadIdx.x == 0) partial[row * chunks_per_row + chunk] = make_float2((float)ds, (float)dq);
}

// mean / rstd of group (b,g) from `np` partial pairs per channel row (fp64, fixed order); every thread gets them
__device__ __forceinline__ void group_moments(const float2* __restrict__ partial, int64_t first_row, int rows, int np,
                                              double count, float eps, double* scratch, float& mean, float& rstd) {
  double s = 0.0, q = 0.0;
  for (int i = threadIdx.x; i < rows * np; i += blockDim.x) {
    const float2 v = __ldg(partial + first_row * np + i);
    s += (double)v.x;
    q += (double)v.y;
  }
  block_sum2(s, q, scratch);
  const double m = s / count;
  double var = q / count - m * m;
  var = var < 0.0 ? 0.0 : var;
  mean = (float)m;
  rstd = (float)(1.0 / sqrt(var + (double)eps));
}

// ---- K2: z = gelu(gamma_c * (y - mean) * rstd + beta_c) -----------------------------------------------------
// grid = (B*C, chunks_per_row); also writes stats[(b*G + g)] = (mean, rstd) (first chunk of the group's first channel)
template <bool V>
__global__ void __launch_bounds__(GN_THREADS)
gn_apply_gelu_kernel(const float* __restrict__ y, const float2* __restrict__ partial, int np, int rows_per_group,
                     const float* __restrict__ gamma,
                     const float* __restrict__ beta, float* __restrict__ z, float2* __restrict__ stats, int64_t T, int C,
                     int G, int chunks_per_row, float eps) {
  __shared__ double scratch[64];
  const int64_t row = blockIdx.x;  // b*C + c
  const int c = (int)(row % C);
  const int64_t b = row / C;
  const int cpg = C / G, g = c / cpg;
  float mean, rstd;
  // partial sums of group (b,g): rows_per_group rows of np pairs each (cpg rows from gn_stats_kernel, one row from the
  // convolution epilogue)
  group_moments(partial, (b * G + g) * (int64_t)rows_per_group, rows_per_group, np, (double)cpg * (double)T, eps, scratch,
                mean, rstd);
  const int chunk = blockIdx.y;
  if (chunk == 0 && c == g * cpg && threadIdx.x == 0) stats[b * G + g] = make_float2(mean, rstd);
  // u = a * (y - mean) + beta: the subtraction first -- a*y + (beta - mean*a) cancels when the group's variance is tiny
  // (rstd ~ 1/sqrt(eps)); the kernel is HBM-bound, the extra FADD is free
  const float a = __ldg(gamma + c) * rstd, sh = __ldg(beta + c);
  const int64_t t0 = (int64_t)chunk * GN_CHUNK;
  const int64_t n = (T - t0) < GN_CHUNK ? (T - t0) : GN_CHUNK;
  if constexpr (V) {
    const float4* p = reinterpret_cast<const float4*>(y + row * T + t0);
    float4* o = reinterpret_cast<float4*>(z + row * T + t0);
#pragma unroll 4
    for (int i = threadIdx.x; i < n / 4; i += GN_THREADS) {
      const float4 v = __ldg(p + i);
      float4 r;
      r.x = gelu_f(fmaf(a, v.x - mean, sh));
      r.y = gelu_f(fmaf(a, v.y - mean, sh));
      r.z = gelu_f(fmaf(a, v.z - mean, sh));
      r.w = gelu_f(fmaf(a, v.w - mean, sh));
      o[i] = r;
    }
  } else {
    const float* p = y + row * T + t0;
    float* o = z + row * T + t0;
#pragma unroll 4
    for (int i = threadIdx.x; i < n; i += GN_THREADS) o[i] = gelu_f(fmaf(a, __ldg(p + i) - mean, sh));
  }
}

// ---- K3: backward partial sums per channel chunk: (sum du*yhat, sum du), du = dz * gelu'(u) ------------------
template <bool V>
__global__ void __launch_bounds__(GN_THREADS)
gn_bwd_stats_kernel(const float* __restrict__ dz, const float* __restrict__ y, const float2* __restrict__ stats,
                    const float* __restrict__ gamma, const float* __restrict__ beta, float2* __restrict__ partial, int64_t T,
                    int C, int G, int chunks_per_row) {
  __shared__ double scratch[64];
  const int64_t row = blockIdx.x;
  const int c = (int)(row % C);
  const int64_t b = row / C;
  const int g = c / (C / G);
  const float2 st = __ldg(stats + b * G + g);
  const float mean = st.x, rstd = st.y, ga = __ldg(gamma + c), be = __ldg(beta + c);
  const int chunk = blockIdx.y;
  const int64_t t0 = (int64_t)chunk * GN_CHUNK;
  const int64_t n = (T - t0) < GN_CHUNK ? (T - t0) : GN_CHUNK;
  float s1 = 0.f, s0 = 0.f;
  auto one = [&](float d, float v) {
    const float yh = (v - mean) * rstd;
    const float du = d * gelu_grad_f(fmaf(ga, yh, be));
    s1 = fmaf(du, yh, s1);
    s0 += du;
  };
  if constexpr (V) {
    const float4* pz = reinterpret_cast<const float4*>(dz + row * T + t0);
    const float4* py = reinterpret_cast<const float4*>(y + row * T + t0);
#pragma unroll 4
    for (int i = threadIdx.x; i < n / 4; i += GN_THREADS) {
      const float4 d = __ldg(pz + i), v = __ldg(py + i);
      one(d.x, v.x);
      one(d.y, v.y);
      one(d.z, v.z);
      one(d.w, v.w);
    }
  } else {
    const float* pz = dz + row * T + t0;
    const float* py = y + row * T + t0;
#pragma unroll 4
    for (int i = threadIdx.x; i < n; i += GN_THREADS) one(__ldg(pz + i), __ldg(py + i));
  }
  double d1 = s1, d0 = s0;
  block_sum2(d1, d0, scratch);
  if (threadIdx.x == 0) partial[row * chunks_per_row + chunk] = make_float2((float)d1, (float)d0);
}

// ---- K4: dy = rstd * (dyhat - m0 - yhat * m1),  dyhat = gamma_c * du, m = group means of dyhat, dyhat*yhat ----
template <bool V>
__global__ void __launch_bounds__(GN_THREADS)
gn_bwd_apply_kernel(const float* __restrict__ dz, const float* __restrict__ y, const float2* __restrict__ stats,
                    const float2* __restrict__ partial, const float* __restrict__ gamma, const float* __restrict__ beta,
                    float* __restrict__ dy, int64_t T, int C, int G, int chunks_per_row) {
  __shared__ double scratch[64];
  const int64_t row = blockIdx.x;
  const int c = (int)(row % C);
  const int64_t b = row / C;
  const int cpg = C / G, g = c / cpg;
  // group sums: sum_c gamma_c * (channel sums), fp64, fixed order
  double m1 = 0.0, m0 = 0.0;
  for (int i = threadIdx.x; i < cpg * chunks_per_row; i += blockDim.x) {
    const int cc = g * cpg + i / chunks_per_row;
    const float2 v = __ldg(partial + (b * C + cc) * chunks_per_row + (i % chunks_per_row));
    const double gm = (double)__ldg(gamma + cc);
    m1 += gm * (double)v.x;
    m0 += gm * (double)v.y;
  }
  block_sum2(m1, m0, scratch);
  const double cnt = (double)cpg * (double)T;
  const float f1 = (float)(m1 / cnt), f0 = (float)(m0 / cnt);
  const float2 st = __ldg(stats + b * G + g);
  const float mean = st.x, rstd = st.y, ga = __ldg(gamma + c), be = __ldg(beta + c);
  const int chunk = blockIdx.y;
  const int64_t t0 = (int64_t)chunk * GN_CHUNK;
  const int64_t n = (T - t0) < GN_CHUNK ? (T - t0) : GN_CHUNK;
  auto one = [&](float d, float v) -> float {
    const float yh = (v - mean) * rstd;
    const float dyh = ga * d * gelu_grad_f(fmaf(ga, yh, be));
    return rstd * (dyh - f0 - yh * f1);
  };
  if constexpr (V) {
    const float4* pz = reinterpret_cast<const float4*>(dz + row * T + t0);
    const float4* py = reinterpret_cast<const float4*>(y + row * T + t0);
    float4* po = reinterpret_cast<float4*>(dy + row * T + t0);
#pragma unroll 4
    for (int i = threadIdx.x; i < n / 4; i += GN_THREADS) {
      const float4 d = __ldg(pz + i), v = __ldg(py + i);
      po[i] = make_float4(one(d.x, v.x), one(d.y, v.y), one(d.z, v.z), one(d.w, v.w));
    }
  } else {
    const float* pz = dz + row * T + t0;
    const float* py = y + row * T + t0;
    float* po = dy + row * T + t0;
#pragma unroll 4
    for (int i = threadIdx.x; i < n; i += GN_THREADS) po[i] = one(__ldg(pz + i), __ldg(py + i));
  }
}

// ---- K5: dgamma_c = sum_{b,chunk} partial.x, dbeta_c = sum partial.y  (grid = C) -----------------------------
__global__ void __launch_bounds__(GN_THREADS)
gn_param_grad_kernel(const float2* __restrict__ partial, int64_t B, int C, int chunks_per_row, float* __restrict__ dgamma,
                     float* __restrict__ dbeta) {
  __shared__ double scratch[64];
  const int c = blockIdx.x;
  double s1 = 0.0, s0 = 0.0;
  const int64_t total = B * chunks_per_row;
  for (int64_t i = threadIdx.x; i < total; i += blockDim.x) {
    const int64_t b = i / chunks_per_row;
    const float2 v = __ldg(partial + (b * C + c) * chunks_per_row + (i % chunks_per_row));
    s1 += (double)v.x;
    s0 += (double)v.y;
  }
  block_sum2(s1, s0, scratch);
  if (threadIdx.x == 0) {
    dgamma[c] = (float)s1;
    dbeta[c] = (float)s0;
  }
}

// ---- row re-pitch (sinc_backward for ragged lengths, sinc_internal.h): dst[r][0..dpitch) = src[r][0..n), zero beyond n ----
template <typename E>
__global__ void __launch_bounds__(256) pad_rows_kernel(const E* __restrict__ src, int64_t spitch, E* __restrict__ dst,
                                                       int64_t dpitch, int64_t n) {
  const int64_t row = blockIdx.y;
  const E* s = src + row * spitch;
  E* d = dst + row * dpitch;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < dpitch; i += (int64_t)gridDim.x * 256)
    d[i] = i < n ? __ldg(s + i) : E(0);
}

cudaError_t launch_pad_rows(const void* src, int64_t src_pitch, void* dst, int64_t dst_pitch, int64_t rows, int64_t n,
                            int esz, cudaStream_t st) {
  if (rows > 65535LL * 32768) return cudaErrorInvalidValue;
  int64_t gx = (dst_pitch + 256 * 8 - 1) / (256 * 8);   // ~8 elements per thread
  if (gx < 1) gx = 1;
  if (gx > 1024) gx = 1024;
  // gridDim.y is limited to 65535: fold larger row counts into several launches
  for (int64_t r0 = 0; r0 < rows; r0 += 65535) {
    const int64_t nr = rows - r0 < 65535 ? rows - r0 : 65535;
    dim3 grid((unsigned)gx, (unsigned)nr);
    if (esz == 4)
      pad_rows_kernel<uint32_t><<<grid, 256, 0, st>>>(reinterpret_cast<const uint32_t*>(src) + r0 * src_pitch, src_pitch,
                                                      reinterpret_cast<uint32_t*>(dst) + r0 * dst_pitch, dst_pitch, n);
    else
      pad_rows_kernel<uint16_t><<<grid, 256, 0, st>>>(reinterpret_cast<const uint16_t*>(src) + r0 * src_pitch, src_pitch,
                                                      reinterpret_cast<uint16_t*>(dst) + r0 * dst_pitch, dst_pitch, n);
    count_launch();
  }
  return cudaGetLastError();
}

static int gn_check(int64_t B, int64_t T, int C, int G, const void* a, const void* b2) {
  if (B < 1 || T < 1 || C < 1 || G < 1 || C % G) return SINC_E_SHAPE;
  if (reinterpret_cast<uintptr_t>(a) % 4 || reinterpret_cast<uintptr_t>(b2) % 4) return SINC_E_ALIGN;
  if (B * C > 2147483647LL) return SINC_E_SHAPE;
  return 0;
}
// 16-byte accesses need every row start 16-byte aligned
static bool gn_vec(int64_t T, const void* a, const void* b2, const void* c3 = nullptr) {
  return T % 4 == 0 && reinterpret_cast<uintptr_t>(a) % 16 == 0 && reinterpret_cast<uintptr_t>(b2) % 16 == 0 &&
         reinterpret_cast<uintptr_t>(c3) % 16 == 0;
}

}  // namespace sinc

using namespace sinc;

extern "C" {

int sinc_gn_chunks(int64_t T) { return (int)((T + GN_CHUNK - 1) / GN_CHUNK); }

int sinc_gn_stats(const float* y, int64_t B, int64_t T, int32_t C, float* partial, void* stream) {
  if (!y || !partial) return SINC_E_NULL;
  int r = gn_check(B, T, C, 1, y, y);
  if (r) return r;
  const int cpr = sinc_gn_chunks(T);
  auto kern = gn_vec(T, y, y) ? gn_stats_kernel<true> : gn_stats_kernel<false>;
  kern<<<dim3((unsigned)(B * C), cpr), GN_THREADS, 0, (cudaStream_t)stream>>>(y, T, cpr, reinterpret_cast<float2*>(partial));
  count_launch();
  return (int)cudaGetLastError();
}

int sinc_gn_gelu_forward(const float* y, const float* partial, int32_t np, int32_t rows_per_group, const float* gamma,
                         const float* beta, float* z, float* stats, int64_t B, int64_t T, int32_t C, int32_t G, float eps,
                         void* stream) {
  if (!y || !partial || !gamma || !beta || !z || !stats) return SINC_E_NULL;
  int r = gn_check(B, T, C, G, y, z);
  if (r) return r;
  if (np < 1 || (rows_per_group != 1 && rows_per_group != C / G)) return SINC_E_SHAPE;
  const int cpr = sinc_gn_chunks(T);
  auto kern = gn_vec(T, y, z) ? gn_apply_gelu_kernel<true> : gn_apply_gelu_kernel<false>;
  kern<<<dim3((unsigned)(B * C), cpr), GN_THREADS, 0, (cudaStream_t)stream>>>(
      y, reinterpret_cast<const float2*>(partial), np, rows_per_group, gamma, beta, z, reinterpret_cast<float2*>(stats), T, C, G,
      cpr, eps);
  count_launch();
  return (int)cudaGetLastError();
}

int sinc_gn_gelu_backward(const float* dz, const float* y, const float* stats, const float* gamma, const float* beta,
                          float* dy, float* dgamma, float* dbeta, float* partial, int64_t B, int64_t T, int32_t C,
                          int32_t G, void* stream) {
  if (!dz || !y || !stats || !gamma || !beta || !dy || !partial) return SINC_E_NULL;
  if ((dgamma == nullptr) != (dbeta == nullptr)) return SINC_E_NULL;
  int r = gn_check(B, T, C, G, dz, y);
  if (r) return r;
  if (reinterpret_cast<uintptr_t>(dy) % 4) return SINC_E_ALIGN;
  const int cpr = sinc_gn_chunks(T);
  cudaStream_t st = (cudaStream_t)stream;
  const bool vec = gn_vec(T, dz, y, dy);
  auto kstats = vec ? gn_bwd_stats_kernel<true> : gn_bwd_stats_kernel<false>;
  auto kapply = vec ? gn_bwd_apply_kernel<true> : gn_bwd_apply_kernel<false>;
  kstats<<<dim3((unsigned)(B * C), cpr), GN_THREADS, 0, st>>>(dz, y, reinterpret_cast<const float2*>(stats), gamma,
                                                                             beta, reinterpret_cast<float2*>(partial), T, C, G, cpr);
  count_launch();
  kapply<<<dim3((unsigned)(B * C), cpr), GN_THREADS, 0, st>>>(dz, y, reinterpret_cast<const float2*>(stats),
                                                                             reinterpret_cast<const float2*>(partial), gamma, beta,
                                                                             dy, T, C, G, cpr);
  count_launch();
  if (dgamma) {
    gn_param_grad_kernel<<<C, GN_THREADS, 0, st>>>(reinterpret_cast<const float2*>(partial), B, C, cpr, dgamma, dbeta);
    count_launch();
  }
  return (int)cudaGetLastError();
}

}  // extern "C"
