// tcgen05 engine, grad-filter:  G[c,k] = sum_b sum_t g[b,c,t] x[b,t+k-H]   (ConvolutionBackward0, grad-weight)
// on the tensor cores with the 3xTF32 split, sm_100a only.  Output: partial tiles for finalize_kernel.
//
// Formulation (DESIGN.md "grad-filter as a time-interleaved GEMM").  The contraction runs over time, 16 samples
// per K-step: row (c,h) of the A operand holds the 8 samples g[c, t0 + 8h .. t0 + 8h + 8), h = 0,1 -- two time
// interleaves per filter give M = 128 rows from 64 filters.  The B operand is the waveform read IN PLACE through
// a Toeplitz descriptor (row pitch 4 samples): B_rho[n, kappa] = xs_rho[4n + kappa + 16q], where the four
// arrays xs_rho[i] = x[tb + i + rho - H] are the tile shifted by rho = 0..3 samples (so every start address
// stays 16-byte aligned).  Then
//     D_rho[(c,h), n]  accumulates  G[c, k]  with  k = 4n + rho - 8h,
// i.e. the two interleaves land on tap columns that differ by a fixed shift of 2 -- N = 80 columns cover
// K <= 312 taps for both (80 % of the MMA is useful work; an MN-major TF32 operand would need the 128-byte
// swizzled layout, which neither the natural g tile nor a Toeplitz view can provide).
//
// A (= grad_y) never touches shared memory: global -> registers -> TF32 hi/lo split -> tcgen05.st into an
// 8-deep TMEM ring; every MMA is the TS form (A from TMEM), so shared memory only serves the 2.5 KB B tiles.
//     TMEM  D_0..D_3 [0,320)  |  A ring [320,448): stage s = 16 columns (8 hi + 8 lo)
//     SMEM  2 x 4 x (hi,lo) shifted waveform tiles, 54 KB
// Accumulators stay in TMEM across tiles and are flushed (read-modify-write of the CTA's own partial tile in
// the workspace) every GF_FLUSH tiles, which bounds the length of any single fp32 accumulation chain.
#include <cuda_bf16.h>

#include "sinc_internal.h"
#include "sinc_synth.cuh"
#include "sinc_tc.cuh"

namespace sinc {

constexpr int GF_THREADS = 416;  // warps 0-3, 9-12: grad_y producers; 4-7: MMA issuers (one per shift rho); 8: waveform producer
constexpr int GF_N = 80;                   // MMA N (tap columns per shift rho)
constexpr int GF_XSTAGE = 512;             // samples per tile = 32 K-steps of 16 samples
constexpr int GF_XLEN = GF_XSTAGE + 4 * GF_N + 16;  // floats per shifted waveform array
constexpr int GF_RING = 8;                 // TMEM A ring depth (K-steps)
constexpr int GF_FLUSH = 8;                // tiles between accumulator flushes (4096 samples)
constexpr int GF_DCOLS = 4 * GF_N;         // 320
constexpr int GF_SM_X = 0;
constexpr int GF_SM_BAR = GF_SM_X + 2 * 8 * GF_XLEN * 4;
constexpr int GF_SM_TOTAL = GF_SM_BAR + 256;

struct BwdFilterTcArgs {
  const float* x;
  int64_t xsb;
  const float* g;
  int64_t B, T;
  int C, K;
  float* partials;  // [gridDim.x / nblk][nblk][128][GF_DCOLS]
  int nblk;         // filter blocks of 64
  int64_t tiles_per_row;
  int64_t ntiles;
};

__global__ void __launch_bounds__(GF_THREADS, 1) bwd_filter_tc_kernel(const BwdFilterTcArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  float* xs = reinterpret_cast<float*>(smem + GF_SM_X);  // [buf][rho][hi/lo][GF_XLEN]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + GF_SM_BAR);
  uint64_t* a_full = bars;            // [GF_RING]
  uint64_t* a_empty = bars + 8;       // [GF_RING]
  uint64_t* x_full = bars + 16;       // [2]
  uint64_t* x_empty = bars + 18;      // [2]
  uint64_t* d_full = bars + 20;
  uint64_t* d_empty = bars + 21;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 22);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int C = a.C, H = (a.K - 1) / 2;
  const int per_blk = gridDim.x / a.nblk;
  const int fblk = blockIdx.x % a.nblk;
  const int slot = blockIdx.x / a.nblk;
  const int c0 = fblk * 64;
  const int64_t T = a.T;

  if (tid == 0) {
    for (int i = 0; i < GF_RING; ++i) {
      tc::mbar_init(&a_full[i], 4);
      tc::mbar_init(&a_empty[i], 4);   // one tcgen05.commit per issuer warp
    }
    for (int i = 0; i < 2; ++i) {
      tc::mbar_init(&x_full[i], 1);
      tc::mbar_init(&x_empty[i], 4);
    }
    tc::mbar_init(d_full, 4);
    tc::mbar_init(d_empty, 4);
    tc::mbar_init_fence();
  }
  if (warp == 4) {
    tc::tmem_alloc(tmem_slot, 512);
    tc::tmem_relinquish();
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = *tmem_slot;

  // contiguous range of tiles for this CTA (tile = 512 samples of one utterance)
  const int64_t tile_begin = a.ntiles * slot / per_blk;
  const int64_t tile_end = a.ntiles * (slot + 1) / per_blk;
  const int64_t my_tiles = tile_end - tile_begin;
  bool ok = true;

  if (warp < 4 || warp >= 9) {
    // ------------- grad_y producers (global -> TF32 hi/lo -> TMEM ring); group 0 also flushes --------------
    // Two groups of four warps (one warp per TMEM lane quarter each) take alternate K-steps, which doubles the
    // bytes in flight per SM: a single group left the DRAM latency exposed (ncu r01g: long-scoreboard stall on
    // the first use of the loaded data, tensor pipe 39 % active).
    const int grp = warp < 4 ? 0 : 1;
    const int quarter = warp & 3;  // tcgen05.ld/st: a warp reaches TMEM lanes 32*(warp%4) .. +31 only
    const int m = 32 * quarter + lane;  // TMEM lane = A row (c_local, h)
    const int cl = m >> 1, h = m & 1;
    const int c = c0 + cl;
    const bool row_valid = c < C;
    const uint32_t lane_base = (uint32_t)(32 * quarter) << 16;
    float* myslot = a.partials + (((size_t)slot * a.nblk + fblk) * 128 + m) * GF_DCOLS;
    int64_t kc0 = 0;  // K-steps before the current tile (ring position of its first K-step)
    bool first_flush = true;
    for (int64_t it = 0; it < my_tiles; ++it) {
      const int64_t tile = tile_begin + it;
      const int64_t b = tile / a.tiles_per_row;
      const int64_t tb = (tile - b * a.tiles_per_row) * GF_XSTAGE;
      const int64_t left = T - tb;
      const int nq = (int)((left < GF_XSTAGE ? left : GF_XSTAGE) + 15) / 16;
      const float* grow = a.g + (b * C + (row_valid ? c : 0)) * T + tb + 8 * h;
      const int qfirst = (int)((kc0 + grp) & 1);  // this group's first K-step in the tile: parity of kc == grp
      const int nmine = nq > qfirst ? (nq - qfirst + 1) / 2 : 0;
      // software pipeline: 4 of this group's K-steps (2 x LDG.128 each) in flight ahead of the conversion
      float4 cur[8], nxt[8];
      auto load4 = [&](float4 (&dst)[8], int i0) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int q = qfirst + 2 * (i0 + u);
          const int64_t t = tb + 16 * q + 8 * h;
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const bool in = row_valid && (i0 + u) < nmine && (t + 4 * j) < T;
            dst[2 * u + j] = in ? __ldg(reinterpret_cast<const float4*>(grow + 16 * q + 4 * j))
                                : make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
      };
      load4(cur, 0);
      for (int i0 = 0; i0 < nmine; i0 += 4) {
        if (i0 + 4 < nmine) load4(nxt, i0 + 4);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (i0 + u < nmine) {
            const int64_t kc = kc0 + qfirst + 2 * (i0 + u);
            const int s = (int)(kc % GF_RING);
            const uint32_t par = (uint32_t)((kc / GF_RING) & 1);
            ok = ok && tc::mbar_wait(&a_empty[s], par ^ 1);
            tc::fence_after_sync();
            const float v[8] = {cur[2 * u].x, cur[2 * u].y, cur[2 * u].z, cur[2 * u].w,
                                cur[2 * u + 1].x, cur[2 * u + 1].y, cur[2 * u + 1].z, cur[2 * u + 1].w};
            uint32_t hi[8], lo[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {  // truncation split: hi = top 19 bits, lo = exact remainder (the MMA
              hi[e] = __float_as_uint(v[e]) & 0xffffe000u;  // ignores the low 13 mantissa bits of lo)
              lo[e] = __float_as_uint(v[e] - __uint_as_float(hi[e]));
            }
            const uint32_t dst = tmem + lane_base + GF_DCOLS + 16 * s;
            tc::tmem_st8(dst, hi);
            tc::tmem_st8(dst + 8, lo);
            tc::tmem_st_wait();
            tc::fence_before_sync();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&a_full[s]);
          }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) cur[i] = nxt[i];
      }
      kc0 += nq;
      // flush the TMEM accumulators into this CTA's partial tile (group 0)
      const bool flush_now = ((it + 1) % GF_FLUSH == 0) || (it + 1 == my_tiles);
      if (flush_now && grp == 0) {
        const uint32_t fpar = (uint32_t)((it / GF_FLUSH) & 1);
        ok = ok && tc::mbar_wait(d_full, fpar);
        tc::fence_after_sync();
#pragma unroll 1
        for (int ch = 0; ch < GF_DCOLS / 32; ++ch) {
          uint32_t v[32];
          tc::tmem_ld32(tmem + lane_base + 32 * ch, v);
          tc::tmem_ld_wait();
          float4* p = reinterpret_cast<float4*>(myslot + 32 * ch);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float4 o = make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]),
                                   __uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3]));
            if (!first_flush) {
              const float4 q = p[i];
              o.x += q.x; o.y += q.y; o.z += q.z; o.w += q.w;
            }
            p[i] = o;
          }
        }
        first_flush = false;
        tc::fence_before_sync();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(d_empty);
      }
    }
    if (my_tiles == 0 && grp == 0) {  // idle CTA: its partial tile must still read as zero
      for (int i = 0; i < GF_DCOLS; i += 4) *reinterpret_cast<float4*>(myslot + i) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  } else if (warp < 8) {
    // ------------- MMA issuers: warp 4+rho owns accumulator D_rho ---------------------------------------
    // One thread can issue a small tcgen05.mma only every ~100 cycles (ncu r01g: the single issuer warp was
    // busy the whole time while the tensor pipe idled 61 %); the four shifts rho have independent accumulators,
    // so four warps issue them concurrently and each keeps its own in-order stream.
    if (lane == 0) {
      const int rho = warp - 4;
      const uint32_t idesc = tc::make_idesc_tf32(128, GF_N, 0, 0);
      const uint64_t bdesc = tc::make_smem_desc(16, 128);  // in-place Toeplitz view, row pitch 4 samples
      const uint32_t d = tmem + GF_N * rho;
      int64_t kc = 0;
      uint32_t acc = 0;  // 0: the next K-step overwrites the accumulator
      for (int64_t it = 0; it < my_tiles; ++it) {
        const int64_t tile = tile_begin + it;
        const int64_t b = tile / a.tiles_per_row;
        const int64_t tb = (tile - b * a.tiles_per_row) * GF_XSTAGE;
        const int64_t left = T - tb;
        const int nq = (int)((left < GF_XSTAGE ? left : GF_XSTAGE) + 15) / 16;
        const int xb = (int)(it & 1);
        ok = ok && tc::mbar_wait(&x_full[xb], (uint32_t)((it >> 1) & 1));
        if (it % GF_FLUSH == 0 && it > 0) {
          // accumulators were handed to the flusher after tile it-1: wait until they have been read out
          ok = ok && tc::mbar_wait(d_empty, (uint32_t)(((it / GF_FLUSH) - 1) & 1));
          acc = 0;
        }
        const uint32_t xbase = tc::smem_addr(xs + (size_t)xb * 8 * GF_XLEN);
        uint64_t bh = tc::desc_at(bdesc, xbase + (2 * rho) * GF_XLEN * 4);
        uint64_t bl = tc::desc_at(bdesc, xbase + (2 * rho + 1) * GF_XLEN * 4);
        for (int q = 0; q < nq; ++q) {
          const int s = (int)(kc % GF_RING);
          ok = ok && tc::mbar_wait(&a_full[s], (uint32_t)((kc / GF_RING) & 1));
          tc::fence_after_sync();
          const uint32_t ah = tmem + GF_DCOLS + 16 * s, al = ah + 8;
          tc::mma_tf32_ts(d, ah, bh, idesc, acc);
          tc::mma_tf32_ts(d, ah, bl, idesc, 1);
          tc::mma_tf32_ts(d, al, bh, idesc, 1);
          acc = 1;
          tc::mma_commit(&a_empty[s]);
          bh += 4;  // +64 bytes (16 samples) in the start-address field
          bl += 4;
          ++kc;
        }
        tc::mma_commit(&x_empty[xb]);
        if ((it + 1) % GF_FLUSH == 0 || it + 1 == my_tiles) tc::mma_commit(d_full);
      }
    }
    __syncwarp();
  } else if (warp == 8) {
    // ------------- waveform producer: four shifted TF32 hi/lo copies of the tile -----------------------
    for (int64_t it = 0; it < my_tiles; ++it) {
      const int64_t tile = tile_begin + it;
      const int64_t b = tile / a.tiles_per_row;
      const int64_t tb = (tile - b * a.tiles_per_row) * GF_XSTAGE;
      const int xb = (int)(it & 1);
      const float* xr = a.x + b * a.xsb;
      constexpr int NV = (GF_XLEN + 3 + 31) / 32;
      float v[NV];
#pragma unroll
      for (int r = 0; r < NV; ++r) {
        const int64_t t = tb - H + lane + 32 * r;
        v[r] = (t >= 0 && t < T) ? __ldg(xr + t) : 0.f;
      }
      ok = ok && tc::mbar_wait(&x_empty[xb], (uint32_t)(((it >> 1) & 1) ^ 1));
      float* base = xs + (size_t)xb * 8 * GF_XLEN;
#pragma unroll
      for (int r = 0; r < NV; ++r) {
        const int ip = lane + 32 * r;  // xs_rho[i] = v(i + rho)  ->  element ip goes to index ip - rho
        const float hi = tc::to_tf32(v[r]);
        const float lo = tc::to_tf32(v[r] - hi);
#pragma unroll
        for (int rho = 0; rho < 4; ++rho) {
          const int i = ip - rho;
          if (i >= 0 && i < GF_XLEN) {
            base[(2 * rho) * GF_XLEN + i] = hi;
            base[(2 * rho + 1) * GF_XLEN + i] = lo;
          }
        }
      }
      tc::fence_async_smem();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&x_full[xb]);
    }
  }
  (void)ok;  // mbar_wait traps on a timeout (sinc_tc.cuh): no silent failure path
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 4) tc::tmem_dealloc(tmem, 512);
}

bool tc_bwd_filter_supported(int C, int K, int64_t T, const void* g, int g_dtype, const DeviceInfo& di) {
  (void)C;
  return K <= 4 * GF_N - 8 && T % 4 == 0 && g_dtype == SINC_DTYPE_F32 && reinterpret_cast<uintptr_t>(g) % 16 == 0 &&
         di.max_smem_optin >= GF_SM_TOTAL + 1024;
}

size_t tc_bwd_filter_partial_bytes(int C) {
  const int nblk = (C + 63) / 64;
  return (size_t)kMaxPartialSlots * nblk * 128 * GF_DCOLS * sizeof(float);
}

cudaError_t launch_bwd_filter_tc(const Bank& bank, const float* x, int64_t x_stride_b, const void* gy, float* ws_partials,
                                 int* slots_used, int64_t B, int64_t T, const DeviceInfo& di, cudaStream_t st) {
  BwdFilterTcArgs a;
  a.x = x;
  a.xsb = x_stride_b;
  a.g = reinterpret_cast<const float*>(gy);
  a.B = B;
  a.T = T;
  a.C = bank.C;
  a.K = bank.K;
  a.partials = ws_partials;
  a.nblk = (bank.C + 63) / 64;
  a.tiles_per_row = ceil_div64(T, GF_XSTAGE);
  a.ntiles = B * a.tiles_per_row;
  int64_t per_blk = di.sm_count / a.nblk;
  if (per_blk < 1) per_blk = 1;
  if (per_blk > a.ntiles) per_blk = a.ntiles;
  if (per_blk > kMaxPartialSlots) per_blk = kMaxPartialSlots;
  const int grid = (int)(per_blk * a.nblk);
  const size_t smem = GF_SM_TOTAL + 1024;
  cudaError_t e = ensure_dynamic_smem(reinterpret_cast<const void*>(bwd_filter_tc_kernel), (int)smem);
  if (e != cudaSuccess) return e;
  bwd_filter_tc_kernel<<<grid, GF_THREADS, smem, st>>>(a);
  count_launch();
  *slots_used = (int)per_blk;
  return cudaGetLastError();
}

}  // namespace sinc
