// FP32-FMA engine of the SincConv hot path (sm_100a).  General in C, K, T; also the parity
// stepping stone for the tcgen05 engine (sinc_tc.cu).
//
//   fwd_fma_kernel         y[b,c,t]  = sum_k w[c,k] x[b,t+k-H]            (perception.py:117-118)
//   bwd_input_fma_kernel   gx[b,t]   = sum_c sum_k w[c,k] g[b,c,t-k+H]    (ConvolutionBackward0, grad-input)
//   bwd_filter_fma_kernel  G[c,k]    = sum_b sum_t g[b,c,t] x[b,t+k-H]    (ConvolutionBackward0, grad-weight)
//   finalize_kernel        G -> dL/dlow_hz_, dL/dband_hz_ in fp64          (autograd of perception.py:89-115)
//
// All convolution kernels are persistent (grid = a multiple of the SM count) and synthesise the
// filters they need into shared memory in their prologue (sinc_synth.cuh).
#include <cuda_bf16.h>

#include "sinc_internal.h"
#include "sinc_finalize.cuh"
#include "sinc_synth.cuh"

namespace sinc {

// =================================================================================================
// forward
// =================================================================================================
constexpr int FW_THREADS = 256;
constexpr int FW_TT = 512;   // time samples per tile: 2 warps-in-time x 2 groups x 32 lanes x 4
constexpr int FW_XPT = 6;    // x-tile elements prefetched per thread (tile length <= 1536)

struct FwdArgs {
  Bank bank;
  const float* x;
  int64_t xsb;
  void* y;
  int64_t B, T;
  int Kp;  // K rounded up to a multiple of 4 (zero taps)
  int64_t ntiles;
  int64_t tiles_per_row;
  int vec_ok;  // y rows are 16-byte (fp32) / 8-byte (bf16) aligned at every t % 4 == 0
};

// Thread tile: 16 filters x (2 groups x 4 consecutive samples).  Warp w: filters 16*(w&3).., time half
// w>>2.  Per 4 taps a thread issues 16 broadcast LDS.128 of w, 2 conflict-free LDS.128 of x and 512 FMA.
template <bool BF16>
__global__ void __launch_bounds__(FW_THREADS, 1) fwd_fma_kernel(const FwdArgs a) {
  extern __shared__ __align__(16) float smem[];
  const int Kp = a.Kp;
  float* ws = smem;            // [Kp][64]  tap-major, filter contiguous
  float* xs = smem + Kp * 64;  // [FW_TT + Kp + 8]
  const int XS = FW_TT + Kp + 8;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int C = a.bank.C, K = a.bank.K, H = (K - 1) / 2;
  const int c0 = blockIdx.y * kChanBlock;
  const int64_t T = a.T;

  for (int i = tid; i < Kp * 64; i += FW_THREADS) ws[i] = 0.f;
  __syncthreads();
  for (int cl = warp; cl < kChanBlock; cl += FW_THREADS / 32) {
    const int c = c0 + cl;
    if (c < C) synth_filter_warp(a.bank, c, lane, [&](int k, double w) { ws[k * 64 + cl] = (float)w; });
  }

  const int cg = warp & 3, th = warp >> 2;
  const int base_a = th * 256 + 4 * lane;
  const int base_b = base_a + 128;

  int64_t tile = blockIdx.x;
  float pre[FW_XPT];
  auto fetch = [&](int64_t tl) {
    const int64_t b = tl / a.tiles_per_row;
    const int64_t t0 = (tl - b * a.tiles_per_row) * FW_TT;
    const float* xr = a.x + b * a.xsb;
#pragma unroll
    for (int r = 0; r < FW_XPT; ++r) {
      const int i = tid + r * FW_THREADS;
      const int64_t t = t0 + i - H;
      pre[r] = (i < XS && t >= 0 && t < T) ? __ldg(xr + t) : 0.f;
    }
  };
  auto stash = [&]() {
#pragma unroll
    for (int r = 0; r < FW_XPT; ++r) {
      const int i = tid + r * FW_THREADS;
      if (i < XS) xs[i] = pre[r];
    }
  };
  if (tile < a.ntiles) {
    fetch(tile);
    stash();
  }
  __syncthreads();

  while (tile < a.ntiles) {
    const int64_t b = tile / a.tiles_per_row;
    const int64_t t0 = (tile - b * a.tiles_per_row) * FW_TT;
    const int64_t next = tile + gridDim.x;
    if (next < a.ntiles) fetch(next);

    float acc[2][16][4];
#pragma unroll
    for (int g = 0; g < 2; ++g)
#pragma unroll
      for (int c = 0; c < 16; ++c)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[g][c][j] = 0.f;

    float4 xa_lo = *reinterpret_cast<const float4*>(xs + base_a);
    float4 xb_lo = *reinterpret_cast<const float4*>(xs + base_b);
    const float* wp = ws + cg * 16;
#pragma unroll 1
    for (int k4 = 0; k4 < Kp; k4 += 4) {
      const float4 xa_hi = *reinterpret_cast<const float4*>(xs + base_a + k4 + 4);
      const float4 xb_hi = *reinterpret_cast<const float4*>(xs + base_b + k4 + 4);
      const float xa[8] = {xa_lo.x, xa_lo.y, xa_lo.z, xa_lo.w, xa_hi.x, xa_hi.y, xa_hi.z, xa_hi.w};
      const float xb[8] = {xb_lo.x, xb_lo.y, xb_lo.z, xb_lo.w, xb_hi.x, xb_hi.y, xb_hi.z, xb_hi.w};
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const float4* w4 = reinterpret_cast<const float4*>(wp + (k4 + kk) * 64);
        const float4 w0 = w4[0], w1 = w4[1], w2 = w4[2], w3 = w4[3];
        const float wv[16] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w,
                              w2.x, w2.y, w2.z, w2.w, w3.x, w3.y, w3.z, w3.w};
#pragma unroll
        for (int c = 0; c < 16; ++c)
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            acc[0][c][j] = fmaf(wv[c], xa[kk + j], acc[0][c][j]);
            acc[1][c][j] = fmaf(wv[c], xb[kk + j], acc[1][c][j]);
          }
      }
      xa_lo = xa_hi;
      xb_lo = xb_hi;
    }

    // epilogue: each warp store instruction writes 32 lanes x 16 B = 512 contiguous bytes of one row
#pragma unroll
    for (int g = 0; g < 2; ++g) {
      const int64_t t = t0 + th * 256 + g * 128 + 4 * lane;
#pragma unroll
      for (int c = 0; c < 16; ++c) {
        const int ch = c0 + cg * 16 + c;
        if (ch < C && t < T) {
          const int64_t off = (b * C + ch) * T + t;
          if (a.vec_ok) {
            if (BF16) {
              __nv_bfloat162 p0 = __floats2bfloat162_rn(acc[g][c][0], acc[g][c][1]);
              __nv_bfloat162 p1 = __floats2bfloat162_rn(acc[g][c][2], acc[g][c][3]);
              uint2 v;
              v.x = *reinterpret_cast<uint32_t*>(&p0);
              v.y = *reinterpret_cast<uint32_t*>(&p1);
              *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(a.y) + off) = v;
            } else {
              *reinterpret_cast<float4*>(reinterpret_cast<float*>(a.y) + off) =
                  make_float4(acc[g][c][0], acc[g][c][1], acc[g][c][2], acc[g][c][3]);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
              if (t + j < T) {
                if (BF16)
                  reinterpret_cast<__nv_bfloat16*>(a.y)[off + j] = __float2bfloat16_rn(acc[g][c][j]);
                else
                  reinterpret_cast<float*>(a.y)[off + j] = acc[g][c][j];
              }
          }
        }
      }
    }
    __syncthreads();  // everyone is done reading xs
    if (next < a.ntiles) stash();
    __syncthreads();
    tile = next;
  }
}

// =================================================================================================
// backward: grad-filter partial tiles
// =================================================================================================
constexpr int BF_THREADS = 256;
constexpr int BF_TT = 128;                               // time samples per tile
constexpr int BF_XR = BF_TT + kTapBlock + 8;             // raw x-chunk length
__host__ __device__ constexpr int bf_pad(int i) { return i + ((i >> 5) << 2); }  // +4 words per 32
constexpr int BF_XSP = ((bf_pad(BF_XR) + 15) / 16) * 16;
constexpr int BF_FLUSH = 64;  // tiles accumulated in registers between flushes (8192 samples)

struct BwdFilterArgs {
  const float* x;
  int64_t xsb;
  const void* g;
  int64_t B, T;
  int C, K;
  float* partials;  // [gridDim.x][Cpad][Kpad]
  int Cpad, Kpad;
  int64_t ntiles;
  int64_t tiles_per_row;
  int vec_ok;  // g rows 16-byte (fp32) / 8-byte (bf16) aligned at every t % 4 == 0
};

// Thread tile: 8 filters x 8 taps of G, accumulated over time.  Warp w: filters 8w..8w+7 (g values are
// warp-broadcast LDS.128); lane l: taps 8l..8l+7 (sliding x window in registers, one LDS.128 per 4
// samples from a +4/32 padded buffer -> conflict free).  256 FMA per 9 LDS.128.
template <bool BF16>
__global__ void __launch_bounds__(BF_THREADS, 2) bwd_filter_fma_kernel(const BwdFilterArgs a) {
  extern __shared__ __align__(16) float smem[];
  float* gs = smem;                         // [2][64][BF_TT]
  float* xs = smem + 2 * 64 * BF_TT;        // [2][BF_XSP]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int C = a.C, H = (a.K - 1) / 2;
  const int c0 = blockIdx.y * kChanBlock;
  const int k0 = blockIdx.z * kTapBlock;
  const int64_t T = a.T;

  const int64_t tile_begin = a.ntiles * blockIdx.x / gridDim.x;
  const int64_t tile_end = a.ntiles * (blockIdx.x + 1) / gridDim.x;

  auto load_tile = [&](int64_t tl, int buf) {
    const int64_t b = tl / a.tiles_per_row;
    const int64_t t0 = (tl - b * a.tiles_per_row) * BF_TT;
    float* gb = gs + buf * 64 * BF_TT;
    float* xb = xs + buf * BF_XSP;
    // x chunk: raw index i <-> x[b, t0 + i - H + k0]
    const float* xr = a.x + b * a.xsb;
    for (int i = tid; i < BF_XR; i += BF_THREADS) {
      const int64_t t = t0 + i - H + k0;
      const bool ok = t >= 0 && t < T;
      cp_async4(xb + bf_pad(i), ok ? (const void*)(xr + t) : (const void*)a.x, ok ? 4 : 0);
    }
    if (!BF16) {
      const float* g = reinterpret_cast<const float*>(a.g);
#pragma unroll
      for (int r = 0; r < 64 * BF_TT / 4 / BF_THREADS; ++r) {
        const int id = tid + r * BF_THREADS;
        const int row = id >> 5, ch = id & 31;
        const int64_t t = t0 + 4 * ch;
        const int c = c0 + row;
        int valid = 0;
        if (c < C && t < T) valid = (T - t) >= 4 ? 4 : (int)(T - t);
        const float* src = g + (b * C + (c < C ? c : 0)) * T + t;
        float* dst = gb + row * BF_TT + 4 * ch;
        if (a.vec_ok) {
          cp_async16(dst, valid ? (const void*)src : (const void*)g, valid * 4);
        } else {
#pragma unroll
          for (int e = 0; e < 4; ++e)
            cp_async4(dst + e, e < valid ? (const void*)(src + e) : (const void*)g, e < valid ? 4 : 0);
        }
      }
    } else {
      const __nv_bfloat16* g = reinterpret_cast<const __nv_bfloat16*>(a.g);
#pragma unroll
      for (int r = 0; r < 64 * BF_TT / 4 / BF_THREADS; ++r) {
        const int id = tid + r * BF_THREADS;
        const int row = id >> 5, ch = id & 31;
        const int64_t t = t0 + 4 * ch;
        const int c = c0 + row;
        int valid = 0;
        if (c < C && t < T) valid = (T - t) >= 4 ? 4 : (int)(T - t);
        const __nv_bfloat16* src = g + (b * C + (c < C ? c : 0)) * T + t;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (a.vec_ok && valid == 4) {
          const uint2 raw = __ldg(reinterpret_cast<const uint2*>(src));
          const __nv_bfloat162 p0 = *reinterpret_cast<const __nv_bfloat162*>(&raw.x);
          const __nv_bfloat162 p1 = *reinterpret_cast<const __nv_bfloat162*>(&raw.y);
          v = make_float4(__low2float(p0), __high2float(p0), __low2float(p1), __high2float(p1));
        } else {
          if (valid > 0) v.x = __bfloat162float(src[0]);
          if (valid > 1) v.y = __bfloat162float(src[1]);
          if (valid > 2) v.z = __bfloat162float(src[2]);
          if (valid > 3) v.w = __bfloat162float(src[3]);
        }
        *reinterpret_cast<float4*>(gb + row * BF_TT + 4 * ch) = v;
      }
    }
    cp_async_commit();
  };

  float acc[8][8];
#pragma unroll
  for (int c = 0; c < 8; ++c)
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[c][k] = 0.f;

  float* slot = a.partials + ((size_t)blockIdx.x * a.Cpad + c0 + 8 * warp) * a.Kpad + k0 + 8 * lane;
  bool first_flush = true;
  auto flush = [&]() {
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      float4* p = reinterpret_cast<float4*>(slot + (size_t)c * a.Kpad);
      float4 v0 = make_float4(acc[c][0], acc[c][1], acc[c][2], acc[c][3]);
      float4 v1 = make_float4(acc[c][4], acc[c][5], acc[c][6], acc[c][7]);
      if (!first_flush) {
        const float4 o0 = p[0], o1 = p[1];
        v0.x += o0.x; v0.y += o0.y; v0.z += o0.z; v0.w += o0.w;
        v1.x += o1.x; v1.y += o1.y; v1.z += o1.z; v1.w += o1.w;
      }
      p[0] = v0;
      p[1] = v1;
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[c][k] = 0.f;
    }
    first_flush = false;
  };

  if (tile_begin < tile_end) load_tile(tile_begin, 0);
  int since_flush = 0;
  for (int64_t tl = tile_begin; tl < tile_end; ++tl) {
    const int buf = (int)((tl - tile_begin) & 1);
    if (tl + 1 < tile_end) {
      load_tile(tl + 1, buf ^ 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();

    const float* gb = gs + buf * 64 * BF_TT + (8 * warp) * BF_TT;
    const float* xb = xs + buf * BF_XSP;
    float xw[12];
    {
      const float4 q0 = *reinterpret_cast<const float4*>(xb + bf_pad(8 * lane));
      const float4 q1 = *reinterpret_cast<const float4*>(xb + bf_pad(8 * lane + 4));
      const float4 q2 = *reinterpret_cast<const float4*>(xb + bf_pad(8 * lane + 8));
      xw[0] = q0.x; xw[1] = q0.y; xw[2] = q0.z; xw[3] = q0.w;
      xw[4] = q1.x; xw[5] = q1.y; xw[6] = q1.z; xw[7] = q1.w;
      xw[8] = q2.x; xw[9] = q2.y; xw[10] = q2.z; xw[11] = q2.w;
    }
#pragma unroll 2
    for (int t = 0; t < BF_TT; t += 4) {
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const float4 gv = *reinterpret_cast<const float4*>(gb + c * BF_TT + t);
        const float ga[4] = {gv.x, gv.y, gv.z, gv.w};
#pragma unroll
        for (int tt = 0; tt < 4; ++tt)
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) acc[c][kk] = fmaf(ga[tt], xw[tt + kk], acc[c][kk]);
      }
      const float4 qn = *reinterpret_cast<const float4*>(xb + bf_pad(t + 8 * lane + 12));
#pragma unroll
      for (int i = 0; i < 8; ++i) xw[i] = xw[i + 4];
      xw[8] = qn.x; xw[9] = qn.y; xw[10] = qn.z; xw[11] = qn.w;
    }
    __syncthreads();  // buffer `buf` may be overwritten by the prefetch of tile tl+2
    if (++since_flush == BF_FLUSH) {
      flush();
      since_flush = 0;
    }
  }
  if (since_flush > 0 || first_flush) flush();  // CTAs without work still zero their slot
}

// =================================================================================================
// backward: grad-input
// =================================================================================================
constexpr int BI_THREADS = 256;
constexpr int BI_TT = 512;  // outputs per tile: 32 lanes x 16 consecutive samples
__host__ __device__ constexpr int bi_pad(int i) { return i + ((i >> 4) << 2); }  // +4 words per 16

struct BwdInputArgs {
  Bank bank;
  const void* g;
  float* gx;
  int64_t B, T;
  int c_begin, c_count;  // this launch reduces filters [c_begin, c_begin + c_count), c_count <= 64
  int Kq;                // flipped filter length incl. `pre` leading zero taps, multiple of 4
  int pre;               // Ha - H, Ha = H rounded up to 4: keeps the row windows 16-byte aligned
  int accumulate;        // add into gx instead of overwriting (2nd.. launches when C > 64)
  int64_t ntiles;
  int64_t tiles_per_row;
  int vec_ok;
};

// One warp per filter row of g (8 rows in flight per CTA, double buffered with cp.async); lane l owns 16
// consecutive outputs with a 20-sample sliding window in registers: 64 FMA per 2 LDS.128.  The +4/16
// padded row layout makes the lane-stride-16 window loads conflict free.  The 8 per-warp partial sums
// are reduced through shared memory at the end of the tile.
template <bool BF16>
__global__ void __launch_bounds__(BI_THREADS, 1) bwd_input_fma_kernel(const BwdInputArgs a) {
  extern __shared__ __align__(16) float smem[];
  const int Kq = a.Kq;
  const int RL = BI_TT + Kq + 16;              // raw row-window length (multiple of 4)
  const int RLP = (bi_pad(RL) + 15) / 16 * 16;  // padded
  float* wf = smem;                            // [c_count<=64][Kq] flipped, zero-prefixed filters
  float* rows = smem + kChanBlock * Kq;        // [8 warps][2][RLP]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int K = a.bank.K, H = (K - 1) / 2;
  const int Ha = H + a.pre;
  const int C = a.bank.C;
  const int64_t T = a.T;

  for (int i = tid; i < kChanBlock * Kq; i += BI_THREADS) wf[i] = 0.f;
  __syncthreads();
  for (int cl = warp; cl < a.c_count; cl += BI_THREADS / 32) {
    float* wr = wf + cl * Kq + a.pre;
    synth_filter_warp(a.bank, a.c_begin + cl, lane, [&](int k, double w) { wr[K - 1 - k] = (float)w; });
  }
  __syncthreads();

  float* myrows = rows + warp * 2 * RLP;

  for (int64_t tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
    const int64_t b = tile / a.tiles_per_row;
    const int64_t t0 = (tile - b * a.tiles_per_row) * BI_TT;

    // raw index i of a row window <-> g[b, c, t0 - Ha + i]
    auto load_row = [&](int cl, int buf) {
      float* dst = myrows + buf * RLP;
      const int c = a.c_begin + cl;
      if (!BF16) {
        const float* gr = reinterpret_cast<const float*>(a.g) + (b * C + c) * T;
        for (int m = lane; m < RL / 4; m += 32) {
          const int64_t t = t0 - Ha + 4 * m;
          float* d = dst + bi_pad(4 * m);
          if (a.vec_ok) {
            const bool ok = t >= 0 && t < T;  // T % 4 == 0 and t % 4 == 0: chunk is all-in or all-out
            cp_async16(d, ok ? (const void*)(gr + t) : a.g, ok ? 16 : 0);
          } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const bool ok = t + e >= 0 && t + e < T;
              cp_async4(d + e, ok ? (const void*)(gr + t + e) : a.g, ok ? 4 : 0);
            }
          }
        }
      } else {
        const __nv_bfloat16* gr = reinterpret_cast<const __nv_bfloat16*>(a.g) + (b * C + c) * T;
        for (int m = lane; m < RL / 4; m += 32) {
          const int64_t t = t0 - Ha + 4 * m;
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (a.vec_ok) {
            if (t >= 0 && t < T) {
              const uint2 raw = __ldg(reinterpret_cast<const uint2*>(gr + t));
              const __nv_bfloat162 p0 = *reinterpret_cast<const __nv_bfloat162*>(&raw.x);
              const __nv_bfloat162 p1 = *reinterpret_cast<const __nv_bfloat162*>(&raw.y);
              v = make_float4(__low2float(p0), __high2float(p0), __low2float(p1), __high2float(p1));
            }
          } else {
            if (t + 0 >= 0 && t + 0 < T) v.x = __bfloat162float(gr[t + 0]);
            if (t + 1 >= 0 && t + 1 < T) v.y = __bfloat162float(gr[t + 1]);
            if (t + 2 >= 0 && t + 2 < T) v.z = __bfloat162float(gr[t + 2]);
            if (t + 3 >= 0 && t + 3 < T) v.w = __bfloat162float(gr[t + 3]);
          }
          *reinterpret_cast<float4*>(dst + bi_pad(4 * m)) = v;
        }
      }
      cp_async_commit();
    };

    float acc[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) acc[j] = 0.f;

    int it = 0;
    if (warp < a.c_count) load_row(warp, 0);
    for (int cl = warp; cl < a.c_count; cl += BI_THREADS / 32, ++it) {
      const int buf = it & 1;
      if (cl + BI_THREADS / 32 < a.c_count) {
        load_row(cl + BI_THREADS / 32, buf ^ 1);
        cp_async_wait<1>();
      } else {
        cp_async_wait<0>();
      }
      __syncwarp();
      const float* row = myrows + buf * RLP;
      const float* wr = wf + cl * Kq;
      float gw[20];
#pragma unroll
      for (int q = 0; q < 5; ++q) {
        const float4 v = *reinterpret_cast<const float4*>(row + bi_pad(16 * lane + 4 * q));
        gw[4 * q + 0] = v.x; gw[4 * q + 1] = v.y; gw[4 * q + 2] = v.z; gw[4 * q + 3] = v.w;
      }
#pragma unroll 5
      for (int k4 = 0; k4 < Kq; k4 += 4) {
        const float4 w4 = *reinterpret_cast<const float4*>(wr + k4);
        const float wa[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
#pragma unroll
          for (int j = 0; j < 16; ++j) acc[j] = fmaf(wa[kk], gw[kk + j], acc[j]);
        const float4 vn = *reinterpret_cast<const float4*>(row + bi_pad(16 * lane + k4 + 20));
#pragma unroll
        for (int i = 0; i < 16; ++i) gw[i] = gw[i + 4];
        gw[16] = vn.x; gw[17] = vn.y; gw[18] = vn.z; gw[19] = vn.w;
      }
      __syncwarp();  // all lanes done with `buf` before the prefetch two rows ahead overwrites it
    }

    __syncthreads();  // every warp is done with its row buffers: reuse them for the reduction
    float* red = rows;  // [8][BI_TT]
#pragma unroll
    for (int j = 0; j < 16; ++j) red[warp * BI_TT + 16 * lane + j] = acc[j];
    __syncthreads();
#pragma unroll
    for (int r = 0; r < BI_TT / BI_THREADS; ++r) {
      const int i = tid + r * BI_THREADS;
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < BI_THREADS / 32; ++w) s += red[w * BI_TT + i];
      const int64_t t = t0 + i;
      if (t < T) {
        float* o = a.gx + b * T + t;
        *o = a.accumulate ? (*o + s) : s;
      }
    }
    __syncthreads();
  }
}

// =================================================================================================
// finalize: fixed-order fp64 reduction of the partial tiles + synthesis chain rule
// =================================================================================================
constexpr int FN_THREADS = 1024;
constexpr int FN_SPLIT = 4;  // slot quarters summed in parallel, combined in fixed order

// layout 0: FMA partial tiles [slot][Cpad][Kpad];  layout 1: tcgen05 partial tiles [slot][nblk][128][320] with
// row (c_local, h) and column rho*80 + n holding tap k = 4n + rho - 8h (sinc_tc_bwd_filter.cu);  layout 2: 16-bit
// tcgen05 partial tiles [slot][nblk][128][128], row (c_local, h), column rho*32 + n holding the single tap
// k = 8n + rho + 4h (sinc_tc16_bwd_filter.cu)
__global__ void __launch_bounds__(FN_THREADS) finalize_kernel(const Bank bank, const float* __restrict__ partials,
                                                              int slots, int Cpad, int Kpad, int layout,
                                                              float* grad_low, float* grad_band,
                                                              float* grad_filters) {
  extern __shared__ __align__(16) double dsm[];
  const int K = bank.K;
  double* Gs = dsm;                  // [K] reduced grad-filter
  double* us = dsm + K;              // [K] un-normalised taps u
  double* scratch = us + K;          // [64]
  double* part = scratch + 64;       // [FN_SPLIT][K]
  double* chi = part + FN_SPLIT * K; // [H] n_j cos(f_hi n_j) / d_j   (the chain-rule factors: independent of G)
  double* clo = chi + K;             // [H] n_j cos(f_lo n_j) / d_j
  const int c = blockIdx.x;
  const int tid = threadIdx.x;
  const bool want_params = grad_low != nullptr && grad_band != nullptr;
  const ChanCoef cc = chan_coef(bank, c);

  // The kernel is latency-bound (64 CTAs, each thread a chain of L2 loads, then fp64 sin / cos): everything that does
  // not depend on the reduced grad-filter -- the un-normalised taps and the cosine factors of the chain rule -- is
  // computed by the LAST warps while the others already have their partial-tile loads in flight, and the loads
  // themselves are issued sixteen at a time (four independent fp64 chains) instead of four.
  if (want_params && tid >= FN_THREADS - 128) finalize_factors(bank, cc, tid - (FN_THREADS - 128), 128, us, chi, clo);
  {
    const int q = tid / (FN_THREADS / FN_SPLIT), kk = tid % (FN_THREADS / FN_SPLIT);
    for (int k = kk; k < K; k += FN_THREADS / FN_SPLIT) {
      double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;   // fixed assignment of slots to chains: bitwise reproducible
      if (layout == 0) {
        const float* p = partials + (size_t)c * Kpad + k;
        const size_t stride = (size_t)Cpad * Kpad;
        int sl = q;
#pragma unroll 4
        for (; sl + 3 * FN_SPLIT < slots; sl += 4 * FN_SPLIT) {
          s0 += (double)__ldg(p + (size_t)sl * stride);
          s1 += (double)__ldg(p + (size_t)(sl + FN_SPLIT) * stride);
          s2 += (double)__ldg(p + (size_t)(sl + 2 * FN_SPLIT) * stride);
          s3 += (double)__ldg(p + (size_t)(sl + 3 * FN_SPLIT) * stride);
        }
        for (; sl < slots; sl += FN_SPLIT) s0 += (double)__ldg(p + (size_t)sl * stride);
      } else if (layout == 2) {
        const int nblk = Cpad / 64;
        const size_t stride = (size_t)nblk * 128 * 128;
        const int h = (k >> 2) & 1, rho = k & 3, n = k >> 3;
        const float* p = partials + ((size_t)(c / 64) * 128 + 2 * (c % 64) + h) * 128 + rho * 32 + n;
        int sl = q;
#pragma unroll 4
        for (; sl + 3 * FN_SPLIT < slots; sl += 4 * FN_SPLIT) {
          s0 += (double)__ldg(p + (size_t)sl * stride);
          s1 += (double)__ldg(p + (size_t)(sl + FN_SPLIT) * stride);
          s2 += (double)__ldg(p + (size_t)(sl + 2 * FN_SPLIT) * stride);
          s3 += (double)__ldg(p + (size_t)(sl + 3 * FN_SPLIT) * stride);
        }
        for (; sl < slots; sl += FN_SPLIT) s0 += (double)__ldg(p + (size_t)sl * stride);
      } else {
        const int nblk = Cpad / 64;
        const size_t stride = (size_t)nblk * 128 * 320;
        const float* p0 = partials + ((size_t)(c / 64) * 128 + 2 * (c % 64)) * 320 + (k & 3) * 80;
        const float* p1 = p0 + 320 + (k + 8) / 4;
        p0 += k / 4;
#pragma unroll 4
        for (int sl = q; sl < slots; sl += FN_SPLIT)
          s0 += (double)__ldg(p0 + (size_t)sl * stride) + (double)__ldg(p1 + (size_t)sl * stride);
      }
      part[q * K + k] = (s0 + s1) + (s2 + s3);
    }
  }
  __syncthreads();
  for (int k = tid; k < K; k += FN_THREADS) {
    double s = 0.0;
#pragma unroll
    for (int q = 0; q < FN_SPLIT; ++q) s += part[q * K + k];
    Gs[k] = s;
    if (grad_filters) grad_filters[(size_t)c * K + k] = (float)s;
  }
  if (!want_params) return;
  __syncthreads();

  finalize_chain_rule<FN_THREADS>(bank, cc, c, Gs, us, chi, clo, scratch, grad_low, grad_band);
}

// debug / test: filters alone
__global__ void __launch_bounds__(256) synth_kernel(const Bank bank, float* out) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c = blockIdx.x * 8 + warp;
  if (c < bank.C) {
    float* o = out + (size_t)c * bank.K;
    synth_filter_warp(bank, c, lane, [&](int k, double w) { o[k] = (float)w; });
  }
}

// =================================================================================================
// host-side launchers
// =================================================================================================
static size_t fwd_smem_bytes(int K) {
  const int Kp = round_up(K, 4);
  return (size_t)(Kp * 64 + FW_TT + Kp + 8) * sizeof(float);
}
static size_t bwd_input_smem_bytes(int K, int* Kq_out, int* pre_out) {
  const int H = (K - 1) / 2;
  const int pre = round_up(H, 4) - H;
  const int Kq = round_up(K + pre, 4);
  const int RL = BI_TT + Kq + 16;
  const int RLP = (bi_pad(RL) + 15) / 16 * 16;
  if (Kq_out) *Kq_out = Kq;
  if (pre_out) *pre_out = pre;
  return (size_t)(kChanBlock * Kq + (BI_THREADS / 32) * 2 * RLP) * sizeof(float);
}

bool fma_supported(int C, int K, const DeviceInfo& di) {
  (void)C;
  return fwd_smem_bytes(K) <= (size_t)di.max_smem_optin &&
         bwd_input_smem_bytes(K, nullptr, nullptr) <= (size_t)di.max_smem_optin &&
         round_up(K, 4) + FW_TT + 8 <= FW_XPT * FW_THREADS;
}

cudaError_t launch_synthesize(const Bank& bank, float* filters_out, cudaStream_t st) {
  synth_kernel<<<(bank.C + 7) / 8, 256, 0, st>>>(bank, filters_out);
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_fwd_fma(const Bank& bank, const float* x, int64_t x_stride_b, void* y, int y_dtype,
                           int64_t B, int64_t T, const DeviceInfo& di, cudaStream_t st) {
  FwdArgs a;
  a.bank = bank;
  a.x = x;
  a.xsb = x_stride_b;
  a.y = y;
  a.B = B;
  a.T = T;
  a.Kp = round_up(bank.K, 4);
  a.tiles_per_row = ceil_div64(T, FW_TT);
  a.ntiles = B * a.tiles_per_row;
  const size_t align = y_dtype == SINC_DTYPE_BF16 ? 8 : 16;
  a.vec_ok = (T % 4 == 0) && (reinterpret_cast<uintptr_t>(y) % align == 0);
  const size_t smem = fwd_smem_bytes(bank.K);
  const int cblocks = (bank.C + kChanBlock - 1) / kChanBlock;
  int64_t per_cb = di.sm_count / cblocks;
  if (per_cb < 1) per_cb = 1;
  const int gx = (int)(a.ntiles < per_cb ? a.ntiles : per_cb);
  dim3 grid(gx, cblocks, 1);
  cudaError_t e;
  if (y_dtype == SINC_DTYPE_BF16) {
    e = ensure_dynamic_smem(reinterpret_cast<const void*>(fwd_fma_kernel<true>), (int)smem);
    if (e != cudaSuccess) return e;
    fwd_fma_kernel<true><<<grid, FW_THREADS, smem, st>>>(a);
  } else {
    e = ensure_dynamic_smem(reinterpret_cast<const void*>(fwd_fma_kernel<false>), (int)smem);
    if (e != cudaSuccess) return e;
    fwd_fma_kernel<false><<<grid, FW_THREADS, smem, st>>>(a);
  }
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_bwd_input_fma(const Bank& bank, const void* gy, int gy_dtype, float* gx, int64_t B,
                                 int64_t T, const DeviceInfo& di, cudaStream_t st) {
  BwdInputArgs a;
  a.bank = bank;
  a.g = gy;
  a.gx = gx;
  a.B = B;
  a.T = T;
  const size_t smem = bwd_input_smem_bytes(bank.K, &a.Kq, &a.pre);
  a.tiles_per_row = ceil_div64(T, BI_TT);
  a.ntiles = B * a.tiles_per_row;
  const size_t align = gy_dtype == SINC_DTYPE_BF16 ? 8 : 16;
  a.vec_ok = (T % 4 == 0) && (reinterpret_cast<uintptr_t>(gy) % align == 0);
  const int grid = (int)(a.ntiles < di.sm_count ? a.ntiles : di.sm_count);
  cudaError_t e;
  if (gy_dtype == SINC_DTYPE_BF16)
    e = ensure_dynamic_smem(reinterpret_cast<const void*>(bwd_input_fma_kernel<true>), (int)smem);
  else
    e = ensure_dynamic_smem(reinterpret_cast<const void*>(bwd_input_fma_kernel<false>), (int)smem);
  if (e != cudaSuccess) return e;
  for (int cb = 0; cb < bank.C; cb += kChanBlock) {
    a.c_begin = cb;
    a.c_count = bank.C - cb < kChanBlock ? bank.C - cb : kChanBlock;
    a.accumulate = cb > 0;
    if (gy_dtype == SINC_DTYPE_BF16)
      bwd_input_fma_kernel<true><<<grid, BI_THREADS, smem, st>>>(a);
    else
      bwd_input_fma_kernel<false><<<grid, BI_THREADS, smem, st>>>(a);
    count_launch();
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
  }
  return cudaSuccess;
}

cudaError_t launch_bwd_filter_fma(const Bank& bank, const float* x, int64_t x_stride_b, const void* gy,
                                  int gy_dtype, float* ws_partials, int* slots_used, int64_t B,
                                  int64_t T, const DeviceInfo& di, cudaStream_t st) {
  BwdFilterArgs a;
  a.x = x;
  a.xsb = x_stride_b;
  a.g = gy;
  a.B = B;
  a.T = T;
  a.C = bank.C;
  a.K = bank.K;
  a.partials = ws_partials;
  const WorkspaceLayout wl = workspace_layout(bank.C, bank.K);
  a.Cpad = wl.Cpad;
  a.Kpad = wl.Kpad;
  a.tiles_per_row = ceil_div64(T, BF_TT);
  a.ntiles = B * a.tiles_per_row;
  const size_t align = gy_dtype == SINC_DTYPE_BF16 ? 8 : 16;
  a.vec_ok = (T % 4 == 0) && (reinterpret_cast<uintptr_t>(gy) % align == 0);
  const size_t smem = (size_t)(2 * 64 * BF_TT + 2 * BF_XSP) * sizeof(float);
  const int cblocks = wl.Cpad / kChanBlock, kblocks = wl.Kpad / kTapBlock;
  int64_t want = 2 * (int64_t)di.sm_count / ((int64_t)cblocks * kblocks);
  if (want < 1) want = 1;
  if (want > kMaxPartialSlots) want = kMaxPartialSlots;
  int gx = (int)(a.ntiles < want ? a.ntiles : want);
  if (gx < 1) gx = 1;
  dim3 grid(gx, cblocks, kblocks);
  cudaError_t e;
  if (gy_dtype == SINC_DTYPE_BF16) {
    e = ensure_dynamic_smem(reinterpret_cast<const void*>(bwd_filter_fma_kernel<true>), (int)smem);
    if (e != cudaSuccess) return e;
    bwd_filter_fma_kernel<true><<<grid, BF_THREADS, smem, st>>>(a);
  } else {
    e = ensure_dynamic_smem(reinterpret_cast<const void*>(bwd_filter_fma_kernel<false>), (int)smem);
    if (e != cudaSuccess) return e;
    bwd_filter_fma_kernel<false><<<grid, BF_THREADS, smem, st>>>(a);
  }
  count_launch();
  *slots_used = gx;
  return cudaGetLastError();
}

cudaError_t launch_finalize(const Bank& bank, const float* ws_partials, int slots, int Cpad, int Kpad, int layout,
                            float* grad_low, float* grad_band, float* grad_filters, cudaStream_t st) {
  const size_t smem = (size_t)((4 + FN_SPLIT) * bank.K + 64) * sizeof(double);
  if (smem > 48 * 1024) {
    cudaError_t e =
        ensure_dynamic_smem(reinterpret_cast<const void*>(finalize_kernel), (int)smem);
    if (e != cudaSuccess) return e;
  }
  finalize_kernel<<<bank.C, FN_THREADS, smem, st>>>(bank, ws_partials, slots, Cpad, Kpad, layout, grad_low, grad_band,
                                                    grad_filters);
  count_launch();
  return cudaGetLastError();
}

}  // namespace sinc
