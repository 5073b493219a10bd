// tcgen05 engine, grad-input:  gx[b,t] = sum_c sum_k w[c,k] g[b,c,t-k+H]   (ConvolutionBackward0, grad-input)
// as GEMM over the filters + col2im, 3xTF32 split, sm_100a only.
//
// The output has ONE channel, so no polyphase row structure exists (any GEMM whose B operand is a Toeplitz view
// of g can only have as many useful A rows as phases: 4).  The tensor-core formulation is the classic dgrad:
//     P[t', k] = sum_c g[c, t'] * w[c, k]          GEMM: M = 128 positions t', N = taps (2 halves of 128), K = C
//     gx[t]    = sum_k P[t - k + H, k]             col2im: P[t', k] is added to output t' + k - H
// A = g^T never touches shared memory: each producer thread owns one position t' (TMEM lane), reads its 64
// channel values with coalesced 4-byte loads, splits them into TF32 hi/lo and writes them with tcgen05.st into a
// 2-stage TMEM ring.  B = the filter bank (hi and lo, K-major, no swizzle, 2 x 64 KB) is built once per CTA by
// the fused fp64 synthesis prologue.  Accumulators: D[:, 0:128) taps 0..127, D[:, 128:256) taps 128..255, one
// issuer warp and four epilogue warps per half.
//
// col2im without atomics or bank conflicts (per epilogue warp, lanes = 32 consecutive positions):
//   phase 1  for a chunk of 32 taps, lane l stores P[t'_l, k0 + j] to S[l][l + j]  (skewed band in a 32 x 64 buffer
//            that is zero outside the band; row pitch 64 makes every store instruction one wavefront)
//   phase 2  lane j sums columns j and 32 + j over the 32 rows (immediate-offset LDS, one wavefront each);
//            the upper 32 column sums overlap the next chunk and are carried in a register
// and writes the sums to the warp's private window; the eight windows and the carry from earlier stages are
// combined in a fixed order, the 128 oldest outputs are complete and written, the rest is carried.  Bitwise
// reproducible.  A CTA owns a contiguous range of 128-sample output blocks and recomputes the 2H-column halo
// at the start of each run instead of exchanging partial sums with its neighbour.
#include <cuda_bf16.h>

#include "sinc_internal.h"
#include "sinc_synth.cuh"
#include "sinc_tc.cuh"

namespace sinc {

constexpr int GI_THREADS = 448;  // warps 0-7 epilogue (half = w/4, quarter = w%4), 8-9 MMA issuers, 10-13 producers
constexpr int GI_NS = 128;       // positions (TMEM lanes) per stage
constexpr int GI_WIN = 384;      // output window of one stage: 128 + 255 rounded up
constexpr int GI_SM_W = 0;                               // filter bank: hi 64 KB, lo 64 KB
constexpr int GI_SM_R = GI_SM_W + 2 * 65536;             // 8 x R ring [64][32] fp32 = 64 KB
constexpr int GI_SM_OUT = GI_SM_R + 8 * 64 * 32 * 4;     // 8 x window [384] fp32 = 12 KB
constexpr int GI_SM_CARRY = GI_SM_OUT + 8 * GI_WIN * 4;  // carry [256] fp32
constexpr int GI_SM_BAR = GI_SM_CARRY + 256 * 4;
constexpr int GI_SM_TOTAL = GI_SM_BAR + 128;

struct BwdInputTcArgs {
  Bank bank;
  const float* g;
  float* gx;
  int64_t B, T;
  int64_t blocks_per_row;  // ceil(T / 128)
  int64_t nblocks;         // B * blocks_per_row
};

__global__ void __launch_bounds__(GI_THREADS, 1) bwd_input_tc_kernel(const BwdInputTcArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* wsm = smem + GI_SM_W;
  float* rring = reinterpret_cast<float*>(smem + GI_SM_R);
  float* outw = reinterpret_cast<float*>(smem + GI_SM_OUT);
  float* carry = reinterpret_cast<float*>(smem + GI_SM_CARRY);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + GI_SM_BAR);
  uint64_t* a_full = bars;        // [2]  producers (4 warps) -> issuers
  uint64_t* a_empty = bars + 2;   // [2]  issuers (2 commits) -> producers
  uint64_t* d_full = bars + 4;    // [2]  per tap half: issuer -> epilogue
  uint64_t* d_empty = bars + 6;   // [2]  per tap half: epilogue (4 warps) -> issuer
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int C = a.bank.C, K = a.bank.K, H = (K - 1) / 2;
  const int64_t T = a.T;

  // ---------------- setup ------------------------------------------------------------------------------------
  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      tc::mbar_init(&a_full[i], 4);
      tc::mbar_init(&a_empty[i], 2);
      tc::mbar_init(&d_full[i], 1);
      tc::mbar_init(&d_empty[i], 4);
    }
    tc::mbar_init_fence();
  }
  if (warp == 8) {
    tc::tmem_alloc(tmem_slot, 512);
    tc::tmem_relinquish();
  }
  for (int i = tid; i < (GI_SM_BAR - GI_SM_W) / 16; i += GI_THREADS)
    reinterpret_cast<float4*>(smem)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  __syncthreads();
  // filter bank, K-major no-swizzle: addr(k, c) = (k/8)*2048 + (c/4)*128 + (k%8)*16 + (c%4)*4
  for (int c = warp; c < C; c += GI_THREADS / 32) {
    synth_filter_warp(a.bank, c, lane, [&](int k, double w) {
      const float wf = (float)w;
      const float hi = tc::to_tf32(wf);
      const int off = (k >> 3) * 2048 + (c >> 2) * 128 + (k & 7) * 16 + (c & 3) * 4;
      *reinterpret_cast<float*>(wsm + off) = hi;
      *reinterpret_cast<float*>(wsm + 65536 + off) = tc::to_tf32(wf - hi);
    });
  }
  tc::fence_async_smem();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = *tmem_slot;

  // ---------------- work: a contiguous range of 128-sample output blocks, cut into per-utterance runs ----------
  const int64_t v0 = a.nblocks * blockIdx.x / gridDim.x;
  const int64_t v1 = a.nblocks * (blockIdx.x + 1) / gridDim.x;
  const int ksteps = (C + 7) / 8;
  bool ok = true;
  int64_t stage = 0;  // global stage counter of this CTA (ring / barrier phases)

  // every role walks the same sequence of (run, stage) pairs
  int64_t v = v0;
  while (v < v1) {
    const int64_t b = v / a.blocks_per_row;
    const int64_t blk0 = v - b * a.blocks_per_row;
    int64_t blk1 = blk0 + (v1 - v);
    if (blk1 > a.blocks_per_row) blk1 = a.blocks_per_row;
    const int64_t s_out = blk0 * GI_NS;                                  // owned outputs [s_out, e_out)
    const int64_t e_out = blk1 * GI_NS < T ? blk1 * GI_NS : T;
    const int nst = (int)((e_out - s_out + 2 * H + GI_NS - 1) / GI_NS);  // stages incl. the 2H-column halo
    v += blk1 - blk0;

    if (warp >= 10) {
      // ------------- producers: g^T -> TF32 hi/lo -> TMEM ring ---------------------------------------------------
      // The 64 channel values of the NEXT stage are loaded into registers before this stage's ring slot is
      // waited for, so the global-load latency overlaps the MMAs of the previous stage.
      const int quarter = warp & 3;
      const uint32_t lane_base = (uint32_t)(32 * quarter) << 16;
      float vv[64];
      auto load_stage = [&](int i) {
        const int64_t tp = s_out - H + (int64_t)GI_NS * i + 32 * quarter + lane;  // this thread's position t'
        const bool in = tp >= 0 && tp < T;
        const float* gp = a.g + (b * C) * T + (in ? tp : 0);
#pragma unroll
        for (int e = 0; e < 64; ++e) vv[e] = (in && e < C) ? __ldg(gp + (int64_t)e * T) : 0.f;
      };
      load_stage(0);
      for (int i = 0; i < nst; ++i) {
        const int64_t sg = stage + i;
        const int sa = (int)(sg & 1);
        ok = ok && tc::mbar_wait(&a_empty[sa], (uint32_t)(((sg >> 1) & 1) ^ 1));
        tc::fence_after_sync();
        const uint32_t dst = tmem + lane_base + 256 + 128 * sa;
#pragma unroll
        for (int cb = 0; cb < 64; cb += 8) {
          uint32_t hi[8], lo[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {  // truncation split: hi = top 19 bits, lo = exact remainder (the MMA
            hi[e] = __float_as_uint(vv[cb + e]) & 0xffffe000u;  // ignores the low 13 mantissa bits of lo)
            lo[e] = __float_as_uint(vv[cb + e] - __uint_as_float(hi[e]));
          }
          tc::tmem_st8(dst + cb, hi);
          tc::tmem_st8(dst + 64 + cb, lo);
        }
        if (i + 1 < nst) load_stage(i + 1);
        tc::tmem_st_wait();
        tc::fence_before_sync();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&a_full[sa]);
      }
    } else if (warp >= 8) {
      // ------------- MMA issuers: warp 8 taps 0..127, warp 9 taps 128..255 --------------------------------------
      if (lane == 0) {
        const int hf = warp - 8;
        const uint32_t idesc = tc::make_idesc_tf32(128, 128, 0, 0);
        const uint64_t wdesc = tc::make_smem_desc(128, 2048);
        const uint64_t bh0 = tc::desc_at(wdesc, tc::smem_addr(wsm) + hf * 32768);
        const uint64_t bl0 = tc::desc_at(wdesc, tc::smem_addr(wsm) + 65536 + hf * 32768);
        const uint32_t d = tmem + 128 * hf;
        for (int i = 0; i < nst; ++i) {
          const int64_t sg = stage + i;
          const int sa = (int)(sg & 1);
          ok = ok && tc::mbar_wait(&d_empty[hf], (uint32_t)((sg & 1) ^ 1));
          ok = ok && tc::mbar_wait(&a_full[sa], (uint32_t)((sg >> 1) & 1));
          tc::fence_after_sync();
          uint32_t ah = tmem + 256 + 128 * sa;
          uint64_t bh = bh0, bl = bl0;
          for (int j = 0; j < ksteps; ++j) {
            tc::mma_tf32_ts(d, ah, bh, idesc, j > 0);
            tc::mma_tf32_ts(d, ah, bl, idesc, 1);
            tc::mma_tf32_ts(d, ah + 64, bh, idesc, 1);
            ah += 8;
            bh += 16;  // +256 bytes: next 8 filters (two 128-byte K chunks)
            bl += 16;
          }
          tc::mma_commit(&a_empty[sa]);
          tc::mma_commit(&d_full[hf]);
        }
      }
      __syncwarp();
    } else {
      // ------------- epilogue: TMEM -> col2im -> window -> gx -------------------------------------------------
      const int hf = warp >> 2, quarter = warp & 3;
      const uint32_t lane_base = (uint32_t)(32 * quarter) << 16;
      float* S = rring + warp * 64 * 32;          // [32 rows][64] skew buffer, zero outside the band
      float* win = outw + warp * GI_WIN;
      const int obase = 32 * quarter + 128 * hf;  // window index of (lane 0, tap 128*hf)
      tc::named_bar_sync(2, 256);                 // previous run's carry writes are done
      carry[tid] = 0.f;                           // new run: nothing carried
      for (int i = 0; i < nst; ++i) {
        const int64_t sg = stage + i;
        const int64_t c0 = s_out - H + (int64_t)GI_NS * i;  // first position of the stage
        ok = ok && tc::mbar_wait(&d_full[hf], (uint32_t)(sg & 1));
        tc::fence_after_sync();
        float tail = 0.f;  // column sums of the upper 32 rows of the previous chunk (they overlap the next chunk)
#pragma unroll 1
        for (int ch = 0; ch < 4; ++ch) {
          uint32_t vv[32];
          tc::tmem_ld32(tmem + lane_base + 128 * hf + 32 * ch, vv);
          tc::tmem_ld_wait();
          if (ch == 3) {  // all 128 columns of this half have left TMEM: release the accumulator
            tc::fence_before_sync();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&d_empty[hf]);
          }
          // phase 1: lane l writes its 32 values to S[l][l + j] (skewed band; cells outside the band stay zero)
          float* srow = S + lane * 65;
#pragma unroll
          for (int j = 0; j < 32; ++j) srow[j] = __uint_as_float(vv[j]);
          __syncwarp();
          // phase 2: column sums; lane owns columns tau = lane and tau = 32 + lane of the 64-wide band
          float s0 = 0.f, s1 = 0.f;
          const float* col = S + lane;
#pragma unroll
          for (int h2 = 0; h2 < 2; ++h2) {  // 32 independent loads in flight, then two pairwise trees
            float t0[16], t1[16];
#pragma unroll
            for (int l2 = 0; l2 < 16; ++l2) {
              t0[l2] = col[(16 * h2 + l2) * 64];
              t1[l2] = col[(16 * h2 + l2) * 64 + 32];
            }
#pragma unroll
            for (int w2 = 8; w2 > 0; w2 >>= 1)
#pragma unroll
              for (int l2 = 0; l2 < w2; ++l2) {
                t0[l2] += t0[l2 + w2];
                t1[l2] += t1[l2 + w2];
              }
            s0 += t0[0];
            s1 += t1[0];
          }
          win[obase + 32 * ch + lane] = s0 + tail;  // every window cell of this warp is written once per stage
          tail = s1;
          __syncwarp();
        }
        win[obase + 128 + lane] = tail;
        // combine the eight windows with the carry in a fixed order; the first 128 outputs are complete
        tc::named_bar_sync(2, 256);
        float tot[2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          const int o = tid + 256 * r;
          float acc = 0.f;
          if (o < GI_WIN) {
            acc = o < 256 ? carry[o] : 0.f;
#pragma unroll
            for (int w = 0; w < 8; ++w) {
              const int tau = o - 32 * (w & 3) - 128 * (w >> 2);  // window cell o <- warp w, row tau in [0,160)
              if (tau >= 0 && tau < 160) acc += outw[w * GI_WIN + o];
            }
          }
          tot[r] = acc;
        }
        tc::named_bar_sync(2, 256);  // everyone has read the old carry and the windows
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          const int o = tid + 256 * r;
          if (o < GI_NS) {
            const int64_t t = c0 - H + o;
            if (t >= s_out && t < e_out) a.gx[b * T + t] = tot[r];
          } else if (o < GI_WIN) {
            carry[o - GI_NS] = tot[r];
          }
        }
      }
    }
    stage += nst;
  }
  (void)ok;  // mbar_wait traps on a timeout (sinc_tc.cuh): no silent failure path
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 8) tc::tmem_dealloc(tmem, 512);
}

bool tc_bwd_input_supported(int C, int K, const void* g, int g_dtype, const DeviceInfo& di) {
  return C <= 64 && K <= 256 && g_dtype == SINC_DTYPE_F32 && reinterpret_cast<uintptr_t>(g) % 4 == 0 &&
         di.max_smem_optin >= GI_SM_TOTAL + 1024;
}

cudaError_t launch_bwd_input_tc(const Bank& bank, const void* gy, float* gx, int64_t B, int64_t T, const DeviceInfo& di,
                                cudaStream_t st) {
  BwdInputTcArgs a;
  a.bank = bank;
  a.g = reinterpret_cast<const float*>(gy);
  a.gx = gx;
  a.B = B;
  a.T = T;
  a.blocks_per_row = ceil_div64(T, GI_NS);
  a.nblocks = B * a.blocks_per_row;
  const int grid = (int)(a.nblocks < di.sm_count ? a.nblocks : di.sm_count);
  const size_t smem = GI_SM_TOTAL + 1024;
  cudaError_t e = ensure_dynamic_smem(reinterpret_cast<const void*>(bwd_input_tc_kernel), (int)smem);
  if (e != cudaSuccess) return e;
  bwd_input_tc_kernel<<<grid, GI_THREADS, smem, st>>>(a);
  count_launch();
  return cudaGetLastError();
}

}  // namespace sinc
