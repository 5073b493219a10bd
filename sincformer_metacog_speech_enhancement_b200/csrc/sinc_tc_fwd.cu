// tcgen05 engine, forward:  y[b,c,t] = sum_k w[c,k] x[b,t+k-H]   (perception.py:117-118) on the 5th-gen
// tensor cores with fp32-grade accuracy (3xTF32 split), sm_100a only.
//
// Formulation (DESIGN.md "forward as a polyphase GEMM").  With t = t0 + 4j + r (r = phase 0..3):
//     Y[(c,r), j] = sum_s  W'[(c,r), s] * X[j, s],     W'[(c,r), s] = w[c, s - r],   X[j, s] = xt[4j + s]
// where xt[i] = x[b, t0 - H + i] is the waveform tile.  X is a Toeplitz matrix with row pitch 4 samples =
// 16 bytes, which is exactly the row pitch of a no-swizzle K-major core matrix: the UMMA descriptor
// (LBO = 16 B, SBO = 128 B) reads it IN PLACE from the 3 KB tile -- no im2col is ever materialised, and a
// K-step of 8 taps is a +32-byte start address.  M = 4 phases x 32 filters = 128 rows per CTA, N = 128
// (512 output samples), K = 256 >= K_taps + 3: 2 % padding instead of the 2x of an im2col.
//
// fp32 accuracy from TF32 tensor cores: w = w_hi + w_lo, x = x_hi + x_lo (each rounded to TF32), and
// y = W_hi X_hi + W_hi X_lo + W_lo X_hi (the dropped W_lo X_lo is ~2^-22 relative).
//
// On-chip budget per CTA (one CTA per SM, 32 filters; CTAs alternate between filter blocks):
//     TMEM  512 columns: D0 [0,128)  D1 [128,256)  (double-buffered fp32 accumulators), W'_hi [256,512)
//     SMEM  W'_lo 128 KB (K-major, no swizzle) | 2 x (x_hi, x_lo) tiles 12 KB | output staging 64.5 KB
// The filter operand used by 2 of the 3 MMAs lives in TMEM, so those MMAs read only the 4 KB B operand from
// shared memory (64 B/clk); only the W'_lo MMA streams its A operand from shared memory.
//
// Warp roles (192 threads): warps 0-3 epilogue (TMEM -> registers -> staging -> 1-D bulk stores of 2 KB
// rows), warp 4 MMA issuer (one thread), warp 5 waveform producer (global -> TF32 hi/lo split -> smem).
#include <cuda_bf16.h>

#include "sinc_internal.h"
#include "sinc_synth.cuh"
#include "sinc_tc.cuh"

namespace sinc {

constexpr int TF_THREADS = 224;  // warps 0-3 epilogue, 4 and 6 MMA issuers (even / odd tiles), 5 waveform producer
constexpr int TF_FB = 32;                 // filters per CTA
constexpr int TF_NJ = 128;                // MMA N: output groups of 4 samples per tile
constexpr int TF_TT = 4 * TF_NJ;          // 512 output samples per tile
constexpr int TF_KS = 256;                // GEMM K (taps + phase shift, zero padded)
constexpr int TF_XT = TF_TT + TF_KS;      // 768 waveform samples per tile
constexpr int TF_STAGE_PITCH = TF_TT + 4; // floats; +4 makes the phase-interleaved staging writes conflict free
constexpr int TF_WPITCH = 257;            // floats, setup scratch

constexpr int TF_SM_ALO = 0;
constexpr int TF_SM_X = TF_SM_ALO + 128 * TF_KS * 4;                 // 131072
constexpr int TF_SM_STAGE = TF_SM_X + 4 * TF_XT * 4;                 // + 12288
constexpr int TF_SM_BAR = TF_SM_STAGE + TF_FB * TF_STAGE_PITCH * 4;  // + 66048
constexpr int TF_SM_TOTAL = TF_SM_BAR + 128;
static_assert(2 * TF_FB * TF_WPITCH * 4 <= TF_FB * TF_STAGE_PITCH * 4, "setup scratch must fit in the staging area");

struct FwdTcArgs {
  Bank bank;
  const float* x;
  int64_t xsb;
  float* y;
  int64_t B, T;
  int nblk;            // filter blocks of 32
  int ksteps;          // ceil((K + 3) / 8) <= 32
  int64_t tiles_per_row;
  int64_t ntiles;      // B * tiles_per_row (per filter block)
};

__global__ void __launch_bounds__(TF_THREADS, 1) fwd_tc_kernel(const FwdTcArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  float* a_lo = reinterpret_cast<float*>(smem + TF_SM_ALO);
  float* xbuf = reinterpret_cast<float*>(smem + TF_SM_X);  // [buf][hi/lo][TF_XT]
  float* stage = reinterpret_cast<float*>(smem + TF_SM_STAGE);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + TF_SM_BAR);
  uint64_t* x_full = bars + 0;     // [2]
  uint64_t* x_empty = bars + 2;    // [2]
  uint64_t* acc_full = bars + 4;   // [2]
  uint64_t* acc_empty = bars + 6;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int C = a.bank.C, K = a.bank.K, H = (K - 1) / 2;
  const int per_blk = gridDim.x / a.nblk;
  const int fblk = blockIdx.x % a.nblk;
  const int slot = blockIdx.x / a.nblk;
  const int c0 = fblk * TF_FB;
  const int64_t T = a.T;

  // ---------------- setup: barriers, TMEM, filters --------------------------------------------------
  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      tc::mbar_init(&x_full[i], 1);
      tc::mbar_init(&x_empty[i], 1);
      tc::mbar_init(&acc_full[i], 1);
      tc::mbar_init(&acc_empty[i], 4);
    }
    tc::mbar_init_fence();
  }
  if (warp == 4) {
    tc::tmem_alloc(tmem_slot, 512);
    tc::tmem_relinquish();
  }
  float* w_hi = stage;                          // [32][TF_WPITCH]
  float* w_lo = stage + TF_FB * TF_WPITCH;      // [32][TF_WPITCH]
  for (int i = tid; i < 2 * TF_FB * TF_WPITCH; i += TF_THREADS) stage[i] = 0.f;
  __syncthreads();
  for (int cl = warp; cl < TF_FB; cl += TF_THREADS / 32) {  // 7 warps share the 32 filters
    const int c = c0 + cl;
    if (c < C)
      synth_filter_warp(a.bank, c, lane, [&](int k, double w) {
        const float wf = (float)w;
        const float hi = tc::to_tf32(wf);
        w_hi[cl * TF_WPITCH + k] = hi;
        w_lo[cl * TF_WPITCH + k] = tc::to_tf32(wf - hi);
      });
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = *tmem_slot;

  // W'_lo -> shared memory, K-major no-swizzle: addr(m, s) = (m/8)*8192 + (s/4)*128 + (m%8)*16 + (s%4)*4
  for (int id = tid; id < 128 * (TF_KS / 4); id += TF_THREADS) {
    const int m = id & 127, kc = id >> 7;
    const int cl = m >> 2, r = m & 3;
    float v[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int k = 4 * kc + e - r;
      v[e] = (k >= 0 && k < K) ? w_lo[cl * TF_WPITCH + k] : 0.f;
    }
    *reinterpret_cast<float4*>(reinterpret_cast<uint8_t*>(a_lo) + (m >> 3) * 8192 + kc * 128 + (m & 7) * 16) =
        make_float4(v[0], v[1], v[2], v[3]);
  }
  // W'_hi -> TMEM columns [256, 512): lane = row m, one 32-bit column per tap
  if (warp < 4) {
    const int m = tid, cl = m >> 2, r = m & 3;
    const uint32_t dst = tmem + ((uint32_t)(32 * warp) << 16) + 256;
    for (int s0 = 0; s0 < TF_KS; s0 += 8) {
      uint32_t v[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int k = s0 + e - r;
        v[e] = __float_as_uint((k >= 0 && k < K) ? w_hi[cl * TF_WPITCH + k] : 0.f);
      }
      tc::tmem_st8(dst + s0, v);
    }
    tc::tmem_st_wait();
  }
  tc::fence_async_smem();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();

  const int64_t my_tiles = slot < a.ntiles ? (a.ntiles - slot + per_blk - 1) / per_blk : 0;
  const uint32_t idesc = tc::make_idesc_tf32(128, TF_NJ, 0, 0);
  bool ok = true;

  if (warp == 5) {
    // ---------------- producer: waveform tile -> TF32 hi / lo in shared memory ----------------------
    for (int64_t it = 0; it < my_tiles; ++it) {
      const int buf = (int)(it & 1);
      const uint32_t par = (uint32_t)((it >> 1) & 1);
      const int64_t tile = slot + it * per_blk;
      const int64_t b = tile / a.tiles_per_row;
      const int64_t t0 = (tile - b * a.tiles_per_row) * TF_TT;
      const float* xr = a.x + b * a.xsb;
      float v[TF_XT / 32];
#pragma unroll
      for (int r = 0; r < TF_XT / 32; ++r) {
        const int64_t t = t0 - H + lane + 32 * r;
        v[r] = (t >= 0 && t < T) ? __ldg(xr + t) : 0.f;
      }
      ok = ok && tc::mbar_wait(&x_empty[buf], par ^ 1);
      float* xh = xbuf + (2 * buf) * TF_XT;
      float* xl = xh + TF_XT;
#pragma unroll
      for (int r = 0; r < TF_XT / 32; ++r) {
        const float hi = tc::to_tf32(v[r]);
        xh[lane + 32 * r] = hi;
        xl[lane + 32 * r] = tc::to_tf32(v[r] - hi);
      }
      tc::fence_async_smem();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&x_full[buf]);
    }
  } else if (warp == 4 || warp == 6) {
    // ---------------- MMA issuers: warp 4 takes the even tiles (accumulator 0), warp 6 the odd ones ----------
    // A single thread issues one small tcgen05.mma per ~100 cycles, more than the 64 cycles an N=128 TF32 MMA
    // occupies the tensor pipe; two issuers with independent accumulators keep the pipe fed.
    if (lane == 0) {
      const int buf = warp == 4 ? 0 : 1;
      const uint64_t bdesc = tc::make_smem_desc(16, 128);      // in-place Toeplitz view of the waveform tile
      const uint64_t adesc0 = tc::desc_at(tc::make_smem_desc(128, 8192), tc::smem_addr(a_lo));  // W'_lo
      const uint32_t d = tmem + (uint32_t)(buf * TF_NJ);
      const uint32_t xh_addr = tc::smem_addr(xbuf + (2 * buf) * TF_XT);
      const uint64_t bh0 = tc::desc_at(bdesc, xh_addr);
      const uint64_t bl0 = tc::desc_at(bdesc, xh_addr + TF_XT * 4);
      for (int64_t it = buf; it < my_tiles; it += 2) {
        const uint32_t par = (uint32_t)((it >> 1) & 1);
        ok = ok && tc::mbar_wait(&acc_empty[buf], par ^ 1);
        ok = ok && tc::mbar_wait(&x_full[buf], par);
        tc::fence_after_sync();
        uint64_t bh = bh0, bl = bl0, al = adesc0;
        uint32_t ah = tmem + 256;
        for (int q = 0; q < a.ksteps; ++q) {
          tc::mma_tf32_ts(d, ah, bh, idesc, q > 0);
          tc::mma_tf32_ts(d, ah, bl, idesc, 1);
          tc::mma_tf32_ss(d, al, bh, idesc, 1);
          bh += 2;   // +32 bytes (8 taps) in the start-address field
          bl += 2;
          al += 16;  // +256 bytes: next two K-chunks of W'_lo
          ah += 8;
        }
        tc::mma_commit(&x_empty[buf]);
        tc::mma_commit(&acc_full[buf]);
      }
    }
    __syncwarp();
  } else if (warp < 4) {
    // ---------------- epilogue: TMEM -> registers -> staging -> bulk stores -------------------------
    const int m = tid, cl = m >> 2, r = m & 3;
    const uint32_t lane_base = (uint32_t)(32 * warp) << 16;
    float* srow = stage + cl * TF_STAGE_PITCH + r;
    for (int64_t it = 0; it < my_tiles; ++it) {
      const int buf = (int)(it & 1);
      const uint32_t par = (uint32_t)((it >> 1) & 1);
      const int64_t tile = slot + it * per_blk;
      const int64_t b = tile / a.tiles_per_row;
      const int64_t t0 = (tile - b * a.tiles_per_row) * TF_TT;
      if (lane < 8) tc::bulk_wait_read<0>();   // staging rows of the previous tile have been read out
      tc::named_bar_sync(1, 128);
      ok = ok && tc::mbar_wait(&acc_full[buf], par);
      tc::fence_after_sync();
      const uint32_t d = tmem + lane_base + (uint32_t)(buf * TF_NJ);
#pragma unroll 1
      for (int ch = 0; ch < TF_NJ / 32; ++ch) {
        uint32_t v[32];
        tc::tmem_ld32(d + 32 * ch, v);
        tc::tmem_ld_wait();
#pragma unroll
        for (int jj = 0; jj < 32; ++jj) srow[4 * (32 * ch + jj)] = __uint_as_float(v[jj]);
      }
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&acc_empty[buf]);   // accumulator may be overwritten
      tc::fence_async_smem();
      tc::named_bar_sync(1, 128);
      if (lane < 8) {
        const int row = 8 * warp + lane;
        const int c = c0 + row;
        if (c < C) {
          const int64_t n = (T - t0) < TF_TT ? (T - t0) : TF_TT;
          tc::bulk_store(a.y + (b * C + c) * T + t0, stage + row * TF_STAGE_PITCH, (uint32_t)(n * 4));
        }
        tc::bulk_commit();
      }
    }
    if (lane < 8) tc::bulk_wait<0>();
  }
  (void)ok;  // mbar_wait traps on a timeout (sinc_tc.cuh): no silent failure path
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 4) tc::tmem_dealloc(tmem, 512);
}

bool tc_fwd_supported(int C, int K, int64_t T, const void* y, int y_dtype, const DeviceInfo& di) {
  (void)C;
  return K + 3 <= TF_KS && T % 4 == 0 && y_dtype == SINC_DTYPE_F32 && reinterpret_cast<uintptr_t>(y) % 16 == 0 &&
         di.max_smem_optin >= TF_SM_TOTAL + 1024;
}

cudaError_t launch_fwd_tc(const Bank& bank, const float* x, int64_t x_stride_b, void* y, int64_t B, int64_t T,
                          const DeviceInfo& di, cudaStream_t st) {
  FwdTcArgs a;
  a.bank = bank;
  a.x = x;
  a.xsb = x_stride_b;
  a.y = reinterpret_cast<float*>(y);
  a.B = B;
  a.T = T;
  a.nblk = (bank.C + TF_FB - 1) / TF_FB;
  a.ksteps = (bank.K + 3 + 7) / 8;
  a.tiles_per_row = ceil_div64(T, TF_TT);
  a.ntiles = B * a.tiles_per_row;
  int64_t per_blk = di.sm_count / a.nblk;
  if (per_blk < 1) per_blk = 1;
  if (per_blk > a.ntiles) per_blk = a.ntiles;
  const int grid = (int)(per_blk * a.nblk);
  const size_t smem = TF_SM_TOTAL + 1024;  // slack for the 1024-byte alignment of the dynamic segment
  cudaError_t e = ensure_dynamic_smem(reinterpret_cast<const void*>(fwd_tc_kernel), (int)smem);
  if (e != cudaSuccess) return e;
  fwd_tc_kernel<<<grid, TF_THREADS, smem, st>>>(a);
  count_launch();
  return cudaGetLastError();
}

}  // namespace sinc
