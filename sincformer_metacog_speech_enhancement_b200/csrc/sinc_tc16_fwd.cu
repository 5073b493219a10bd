// tcgen05 engine, forward, 16-bit split operands:  y[b,c,t] = sum_k w[c,k] x[b,t+k-H]   (perception.py:117-118)
// on kind::f16 tensor cores with fp32-grade accuracy: scaled fp16 hi/lo split, sm_100a only.
//
// Precision.  fp16 has the 11-bit significand of TF32 at twice the tensor-core rate; its 5-bit exponent is handled
// by exact power-of-two scaling.  Per waveform tile sx = 2^e puts max|x| in [2^13, 2^14); per filter sw_c does the
// same for max_k|w[c,k]|.  Then  x*sx = x_hi + x_lo,  w*sw = w_hi + w_lo  (each rounded to fp16, 22 bits together;
// elements more than 2^-17 below their tile/filter maximum lose relative -- not max-norm -- precision) and
//     y = (W_hi X_hi + W_hi X_lo + W_lo X_hi) / (sx * sw_c)          fp32 accumulation in TMEM,
// the dropped W_lo X_lo term is 2^-22 relative.  Against the fp64 oracle: ~1e-7 max-norm relative (gate 1e-5).
//
// Formulation.  With t = t0 + 8j + r (r = phase 0..7):
//     Y[(c,r), j] = sum_s  W'[(c,r), s] * X[j, s],     W'[(c,r), s] = w[c, s - r],   X[j, s] = xt[8j + s]
// where xt[i] = x[b, t0 - H + i] is the waveform tile.  X is a Toeplitz matrix with a row pitch of 8 fp16 samples
// = 16 bytes = the row pitch of a no-swizzle K-major core matrix, so the UMMA descriptor (LBO = 16, SBO = 128)
// reads it IN PLACE from the 2.3 KB tile; a K-step of 16 taps is a +32-byte start address.  M = 8 phases x 16
// filters = 128 rows per CTA, N = 112 (896 output samples), K = 272 >= K_taps + 7.
//
// On-chip budget per CTA (one CTA per SM, 16 filters; CTAs alternate between the filter blocks):
//     TMEM  D0 [0,112)  D1 [128,240)  W'_hi [240,376)  W'_lo [376,512)   (16-bit operands: two taps per column)
//     SMEM  2 x (x_hi, x_lo) tiles 9 KB | 3 x output staging 57 KB (the bulk stores of two tiles may still be draining)
// Every MMA takes A from TMEM; shared memory only serves the 3.5 KB B operand (64 B/clk).
//
// Warp roles (352 threads): warps 0-7 epilogue (two per TMEM lane quarter, 56 columns each: TMEM -> scale ->
// phase-interleaved staging -> 1-D bulk stores of 3.5 KB rows; plain global stores straight from the registers were
// measured and are slower: main loop 158k -> 182k cycles), warp 8 MMA issuer (warp-uniform issue loop), warps 9 and 10
// waveform producers (global -> max -> scale -> fp16 hi/lo -> smem; one per tile buffer).
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <cstdio>
#include <cstdlib>
#include <type_traits>

#include "sinc_internal.h"
#include "sinc_synth.cuh"
#include "sinc_tc.cuh"

namespace sinc {

constexpr int U_THREADS = 352;  // warps 0-7 epilogue (lane quarter = w%4, column half = w/4), 8 MMA issuer, 9-10 waveform producers
constexpr int U_FB = 16;                  // filters per CTA
constexpr int U_NJ = 112;                 // MMA N: output groups of 8 samples per tile
constexpr int U_TT = 8 * U_NJ;            // 896 output samples per tile
constexpr int U_KS = 272;                 // GEMM K (taps + phase shift, zero padded): 17 K-steps
constexpr int U_XT = U_TT + U_KS;         // 1168 waveform samples per tile
constexpr int U_PITCH = U_TT + 8;         // floats; +8 makes the phase-interleaved staging writes conflict free
constexpr int U_WPITCH = 272;             // floats, setup scratch (one filter per row)
constexpr int U_TM_D1 = 128, U_TM_WHI = 240, U_TM_WLO = 376;
// Prologue scratch (inside the output staging buffers, which are idle until the main loop starts), per filter: the scaled
// taps split into fp16 hi / lo ONCE and packed in pairs, in two alignments -- word j of the even copy holds taps
// (2j - 8, 2j - 7), of the odd copy (2j - 9, 2j - 8); taps outside [0, K) are zero -- so that row (filter, phase r) of W'
// is a run of consecutive words of one copy.  Offsets (32-bit words) chosen so that the 32 lanes of a warp (4 filters x 8
// phases) read 32 distinct banks: odd copy at +164 (4 mod 32), lo copies at +328, filter pitch 648 (8 mod 32).
constexpr int U_PLEN = 144, U_P_ODD = 164, U_P_LO = 328, U_PPITCH = 648;
constexpr int U_SCR_P = U_FB * U_WPITCH;                 // float offset of the packed area within the scratch
constexpr int U_SCR_N = U_SCR_P + U_FB * U_PPITCH;       // n_ [<= 144]
constexpr int U_SCR_WIN = U_SCR_N + 144;                 // window [<= 272]
constexpr int U_SCR_CC = U_SCR_WIN + U_KS;                // per-filter scalars (ChanCoef) [16]
constexpr int U_SCR_END = U_SCR_CC + U_FB * 8;
static_assert(sizeof(ChanCoef) <= 32 && (U_SCR_CC * 4) % 8 == 0, "ChanCoef slots");
static_assert(U_P_LO + U_P_ODD + U_PLEN <= U_PPITCH && U_P_ODD >= U_PLEN && U_P_LO >= U_P_ODD + U_PLEN, "packed layout");
static_assert((U_SCR_P * 4) % 16 == 0 && (U_FB * U_PPITCH) % 4 == 0, "the packed area is zeroed with 16-byte stores");

constexpr int U_SM_X = 0;                                     // [buf 2][hi/lo 2][U_XT] fp16
constexpr int U_SM_STAGE = U_SM_X + ((2 * 2 * U_XT * 2 + 127) / 128) * 128;
constexpr int U_NSTAGE = 3;               // output staging buffers: the bulk stores of two tiles may still be draining
constexpr int U_SM_SCALE = U_SM_STAGE + U_NSTAGE * U_FB * U_PITCH * 4;   // x inverse scales [4], filter inverse scales [16]
constexpr int U_SM_BAR = U_SM_SCALE + 256;   // scale area: xinv [4], winv [16], GroupNorm warp sums [8][2]
constexpr int U_SM_TOTAL = U_SM_BAR + 128;
static_assert(U_SCR_END * 4 <= U_NSTAGE * U_FB * U_PITCH * 4, "setup scratch must fit in the staging buffers");

struct FwdTc16Args {
  Bank bank;
  const float* x;
  int64_t xsb;
  void* y;             // fp32 or bf16 (bf16-output mode: fp32 accumulators, rounded once in the epilogue)
  int64_t B, T;
  int nblk;            // filter blocks of 16
  int ksteps;          // ceil((K + 7) / 16) <= 17
  long long* phase;    // diagnostic time stamps (CTA 0), may be null
  int64_t tiles_per_row;
  int64_t ntiles;      // B * tiles_per_row (per filter block)
  float2* gn_partial;  // optional (next row N1): per (utterance, group of 8 filters, tile) sums of y and y^2 for the
                       // GroupNorm that follows the layer (agents/perception.py:233), [B][C/8][tiles_per_row]
#ifdef SINC_ABLATION
  int ablate;          // ablation build only: bit 0 epilogue skips TMEM load + staging writes, 1 no stores to y, 2 no MMAs,
                       // 3 producers skip loads + conversion
#endif
};
#ifdef SINC_ABLATION
#define FABL(bit) ((a.ablate >> (bit)) & 1)
#else
#define FABL(bit) 0
#endif

// power-of-two scale that puts `m` (>= 0) into [2^13, 2^14): returns the exponent, clamped so that both the
// scale and its inverse are normal floats; m = 0 or non-finite -> 0
__device__ __forceinline__ int u_scale_exp(float m) {
  const uint32_t bits = __float_as_uint(m);
  const int e = (int)((bits >> 23) & 0xffu) - 127;
  if (m == 0.f || e == 128) return 0;
  int se = 13 - e;
  se = se > 100 ? 100 : se;
  se = se < -100 ? -100 : se;
  return se;
}
__device__ __forceinline__ float u_pow2(int e) { return __uint_as_float((uint32_t)(127 + e) << 23); }

// v (already scaled) -> fp16 hi (RN) and fp16 lo = RN(v - hi), as raw 16-bit patterns
__device__ __forceinline__ void u_split_f16(float v, uint16_t& hi, uint16_t& lo) {
  const __half h = __float2half_rn(v);
  const __half l = __float2half_rn(v - __half2float(h));
  hi = __half_as_ushort(h);
  lo = __half_as_ushort(l);
}

// One filter of the forward prologue, whole warp: synthesis (same arithmetic and order as synth_filter_warp in sinc_synth.cuh:
// bit-identical taps) -> power-of-two scale -> fp16 hi / lo split -> the two packed copies described at U_PLEN.  Written for
// a small code footprint -- real loops, one expansion of cosf / sinf: a prologue runs its code exactly once per warp, and
// what that costs is mostly the instruction fetch of code that is in no cache (ncu r02z: 63 % `no_inst` stalls in the
// 1356 straight-line instructions of the previous version; 40k cycles of prologue in a 200k-cycle kernel).  n_ and the
// window are read from shared-memory copies: a global load per loop iteration was ~700 cycles of exposed latency.
static __device__ __noinline__ void u_synth_pack(int K, const ChanCoef cc, int lane, float* row, uint32_t* P, const float* n_s,
                                                 const float* win_s, float* winv_out) {
  const int H = (K - 1) / 2;
  const float fs = (float)(0.5 * (cc.f_hi + cc.f_lo));
  const float fd = (float)(0.5 * (cc.f_hi - cc.f_lo));
  double sumabs = 0.0;
  float mxu = 0.f;
#pragma unroll 1
  for (int j = lane; j < H; j += 32) {
    const float nj = n_s[j];
    const float left = __fdividef(2.0f * cosf(fs * nj) * sinf(fd * nj), nj * 0.5f + 1e-8f);
    const float l = left * win_s[j], rr = left * win_s[K - 1 - j];
    sumabs += (double)fabsf(l) + (double)fabsf(rr);
    mxu = fmaxf(mxu, fmaxf(fabsf(l), fabsf(rr)));
    row[j] = l;
    row[K - 1 - j] = rr;
  }
  const float uc = (float)(2.0 * (cc.f_hi - cc.f_lo)) * win_s[H];
  if (lane == 0) {
    sumabs += (double)fabsf(uc);
    mxu = fmaxf(mxu, fabsf(uc));
    row[H] = uc;
  }
  const double rS = 1.0 / (warp_sum(sumabs) + 1e-8);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mxu = fmaxf(mxu, __shfl_xor_sync(0xffffffffu, mxu, o));
  // max_k |w_k| with w_k = (float)((double)u_k * rS): rounding is monotonic, so it is the image of max_k |u_k|
  const int se = u_scale_exp((float)((double)mxu * rS));
  if (lane == 0) *winv_out = u_pow2(-se);
  const float sw = u_pow2(se);
  __syncwarp();
#pragma unroll 1
  for (int j = lane; j < U_PLEN; j += 32) {
    const int k = 2 * j - 9;  // taps k, k+1, k+2: the odd copy's word j is (k, k+1), the even copy's (k+1, k+2)
    uint16_t h[3], l[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const int kk = k + i;
      const float w = (kk >= 0 && kk < K) ? (float)((double)row[kk] * rS) * sw : 0.f;  // the scaling is exact
      u_split_f16(w, h[i], l[i]);
    }
    P[j] = (uint32_t)h[1] | ((uint32_t)h[2] << 16);
    P[U_P_LO + j] = (uint32_t)l[1] | ((uint32_t)l[2] << 16);
    P[U_P_ODD + j] = (uint32_t)h[0] | ((uint32_t)h[1] << 16);
    P[U_P_ODD + U_P_LO + j] = (uint32_t)l[0] | ((uint32_t)l[1] << 16);
  }
}

// RG ("ragged"): T is not a multiple of 4 (fp32) / 8 (bf16) or y is not 16-byte aligned, so the rows of y do not start
// on 16-byte boundaries and bulk stores cannot write them.  The staged tile then goes out with plain coalesced stores
// from all eight epilogue warps (measured ~15 % slower than the bulk path for this kernel -- against 9x for the FMA
// engine, which used to take every such length).
template <bool OB, bool GN, bool RG>
__global__ void __launch_bounds__(U_THREADS, 1) fwd_tc16_kernel(const FwdTc16Args a) {
  constexpr int ESZ = OB ? 2 : 4;                 // bytes per output element
  constexpr int PITCHB = U_PITCH * ESZ;           // staging row pitch in bytes (32 resp. 16 mod 128: conflict free)
  extern __shared__ __align__(1024) uint8_t smem[];
  uint16_t* xbuf = reinterpret_cast<uint16_t*>(smem + U_SM_X);  // [buf][hi/lo][U_XT]
  float* stage = reinterpret_cast<float*>(smem + U_SM_STAGE);   // [2][U_FB][U_PITCH]
  float* xinv = reinterpret_cast<float*>(smem + U_SM_SCALE);    // [4]  1 / sx of tile it & 3
  float* winv = xinv + 4;                                       // [16] 1 / sw_c
  float* gnred = winv + 16;                                     // [8 warps][2] GroupNorm partial sums of a tile
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + U_SM_BAR);
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
  const int c0 = fblk * U_FB;
  const int64_t T = a.T;

  // ---------------- setup: barriers, TMEM, filters --------------------------------------------------
  if (blockIdx.x == 0 && threadIdx.x == 0 && a.phase) {
    a.phase[0] = clock64();
    a.phase[1] = (long long)tc::global_timer_ns();
  }
  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      tc::mbar_init(&x_full[i], 1);
      tc::mbar_init(&x_empty[i], 1);
      tc::mbar_init(&acc_full[i], 1);
      tc::mbar_init(&acc_empty[i], 8);
    }
    tc::mbar_init_fence();
  }
  if (warp == 4) {
    tc::tmem_alloc(tmem_slot, 512);
    tc::tmem_relinquish();
  }
  float* wsc = stage;                                            // scratch: raw taps [16][U_WPITCH]
  uint32_t* pk = reinterpret_cast<uint32_t*>(stage + U_SCR_P);   // packed fp16 hi / lo copies [16][U_PPITCH]
  float* n_s = stage + U_SCR_N;
  float* win_s = stage + U_SCR_WIN;
  for (int i = tid; i < U_FB * U_PPITCH / 4; i += U_THREADS)     // taps outside [0, K) and filters beyond C read as zero
    reinterpret_cast<uint4*>(pk)[i] = make_uint4(0u, 0u, 0u, 0u);
  tc::griddep_wait();     // everything above touched shared memory and TMEM only
  tc::griddep_launch();
#ifdef SINC_ABLATION
  const long long dbg_c0 = clock64();
  long long dbg_c1 = 0, dbg_c2 = 0;
#endif
  // the per-filter scalars (parameter loads + an fp64 division: ~1k cycles of latency) for all 16 filters at once, next to
  // the loads of n_ and the window, instead of at the head of every filter's synthesis
  ChanCoef* cc_s = reinterpret_cast<ChanCoef*>(stage + U_SCR_CC);
  if (tid >= U_THREADS - U_FB && c0 + (tid - (U_THREADS - U_FB)) < C)
    cc_s[tid - (U_THREADS - U_FB)] = chan_coef(a.bank, c0 + (tid - (U_THREADS - U_FB)));
  for (int i = tid; i < H; i += U_THREADS) n_s[i] = a.bank.n[i];
  for (int i = tid; i < K; i += U_THREADS) win_s[i] = a.bank.window[i];
  __syncthreads();
  if (warp < U_FB / 2) {   // warps 0-7: two filters each, one after the other through ONE small loop body
#pragma unroll 1
    for (int f = 0; f < 2; ++f) {
      const int cl = 2 * warp + f;
      if (c0 + cl < C)
        u_synth_pack(K, cc_s[cl], lane, wsc + cl * U_WPITCH, pk + cl * U_PPITCH, n_s, win_s, &winv[cl]);
      else if (lane == 0)
        winv[cl] = 1.f;
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
#ifdef SINC_ABLATION
  dbg_c1 = clock64();
#endif
  const uint32_t tmem = *tmem_slot;

  // W' -> TMEM: lane = row m = 8*cl + r, column = tap pair (s, s+1), W'[m, s] = w[cl, s - r]: eight consecutive words of the
  // copy whose alignment matches the parity of r (two warps per TMEM lane quarter, alternate 16-tap column groups)
  if (warp < 8) {
    const int m = tid & 127, cl = m >> 3, r = m & 7;
    const uint32_t base = tmem + ((uint32_t)(32 * (warp & 3)) << 16);
    const uint32_t* pr = pk + cl * U_PPITCH + ((r & 1) ? U_P_ODD + (9 - r) / 2 : (8 - r) / 2);
#pragma unroll 1
    for (int s0 = 16 * (warp >> 2); s0 < U_KS; s0 += 32) {
      uint32_t hi[8], lo[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        hi[e] = pr[s0 / 2 + e];
        lo[e] = pr[U_P_LO + s0 / 2 + e];
      }
      tc::tmem_st8(base + U_TM_WHI + s0 / 2, hi);
      tc::tmem_st8(base + U_TM_WLO + s0 / 2, lo);
    }
    tc::tmem_st_wait();
  }
  tc::fence_before_sync();
  __syncthreads();  // the setup scratch becomes staging buffer 0
  tc::fence_after_sync();
#ifdef SINC_ABLATION
  dbg_c2 = clock64();
  if (FABL(5) && tid == 0 && (blockIdx.x == 0 || blockIdx.x == 77))
    printf("fwd prologue cta %d: zero+synth %lld cycles, W' build %lld cycles\n", (int)blockIdx.x, dbg_c1 - dbg_c0, dbg_c2 - dbg_c1);
#endif

  const int64_t my_tiles = slot < a.ntiles ? (a.ntiles - slot + per_blk - 1) / per_blk : 0;
  bool ok = true;
  if (blockIdx.x == 0 && threadIdx.x == 0 && a.phase) {
    a.phase[2] = clock64();
    a.phase[3] = (long long)tc::global_timer_ns();
  }

  if (warp == 9 || warp == 10) {
    // ---------------- producers: waveform tile -> max -> scale -> fp16 hi / lo in shared memory ---------------
    // TWO warps, one per tile buffer (warp 9 the even tiles, warp 10 the odd ones); each issues the loads of its NEXT
    // tile (two tiles ahead) before it converts the current one.  With a single producer looking one tile ahead the
    // issuer polled x_full 38 times per tile (ncu r02c): the waveform is read cold from DRAM while the kernel streams
    // 3 TB/s of output into the same memory system, and one tile of MMAs (1.8 us) does not cover that latency.
    constexpr int NV = (U_XT + 31) / 32;
    const int pb = warp - 9;                       // this producer's buffer = tile parity
    float v[NV], vn[NV];
    auto load_tile = [&](float (&dst)[NV], int64_t it) {
      const int64_t tile = slot + it * per_blk;
      const int64_t b = tile / a.tiles_per_row;
      const int64_t t0 = (tile - b * a.tiles_per_row) * U_TT;
      const float* xr = a.x + b * a.xsb;
#pragma unroll
      for (int r = 0; r < NV; ++r) {
        const int i = lane + 32 * r;
        const int64_t t = t0 - H + i;
        dst[r] = (i < U_XT && t >= 0 && t < T) ? __ldg(xr + t) : 0.f;
      }
    };
    if (pb < my_tiles) load_tile(v, pb);
    for (int64_t it = pb; it < my_tiles; it += 2) {
      const int buf = pb;
      const uint32_t par = (uint32_t)((it >> 1) & 1);
      if (it + 2 < my_tiles && !FABL(3)) load_tile(vn, it + 2);
      float mx = 0.f;
#pragma unroll
      for (int r = 0; r < NV; ++r) mx = fmaxf(mx, fabsf(v[r]));
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      const int se = u_scale_exp(mx);
      const float sx = u_pow2(se);
      ok = ok && tc::mbar_wait(&x_empty[buf], par ^ 1);
      if (lane == 0) xinv[it & 3] = u_pow2(-se);
      uint16_t* xh = xbuf + (2 * buf) * U_XT;
      uint16_t* xl = xh + U_XT;
      if (!FABL(3))
#pragma unroll
      for (int r = 0; r < NV; ++r) {
        const int i = lane + 32 * r;
        if (i < U_XT) {
          uint16_t h, l;
          u_split_f16(v[r] * sx, h, l);
          xh[i] = h;
          xl[i] = l;
        }
      }
      tc::fence_async_smem();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&x_full[buf]);
#pragma unroll
      for (int r = 0; r < NV; ++r) v[r] = vn[r];
    }
  } else if (warp == 8) {
    // ---------------- MMA issuer: all tiles, accumulators alternate -----------------------------------------
    // All lanes run the loop (warp-uniform operands stay in uniform registers); the elected lane issues.
    // (tools/mma_rate.py: an N = 112 TS-form MMA costs 66 cycles whatever the accumulator rotation, so one
    // in-order stream loses nothing against two issuers and keeps the MMAs of a tile back to back.)
    const int ntl = tc::uniform((int)my_tiles);
    const int nks = tc::uniform(a.ksteps);
    const uint32_t tm = tc::uniform((int)tmem);
    const uint32_t idesc = tc::make_idesc_f16(128, U_NJ, tc::kFmtF16, tc::kFmtF16);
    const uint64_t bdesc = tc::make_smem_desc(16, 128);      // in-place Toeplitz view of the waveform tile
    const uint32_t x0 = tc::uniform((int)tc::smem_addr(xbuf));
    for (int it = 0; it < ntl; ++it) {
      const int buf = it & 1;
      const uint32_t par = (uint32_t)((it >> 1) & 1);
      const uint32_t d = tm + (uint32_t)(buf * U_TM_D1);
      const uint32_t xh_addr = x0 + (2 * buf) * U_XT * 2;
      ok = ok && tc::mbar_wait(&acc_empty[buf], par ^ 1);
      ok = ok && tc::mbar_wait(&x_full[buf], par);
      tc::fence_after_sync();
      if (tc::elect_one()) {
        uint32_t bh = xh_addr, bl = xh_addr + U_XT * 2;
        uint32_t ah = tm + U_TM_WHI, al = tm + U_TM_WLO;
        for (int q = 0; q < (FABL(2) ? 0 : nks); ++q) {
          const uint64_t dh = tc::desc_at(bdesc, bh), dl = tc::desc_at(bdesc, bl);
          tc::mma_f16_ts(d, ah, dh, idesc, q > 0);
          tc::mma_f16_ts(d, ah, dl, idesc, 1);
          tc::mma_f16_ts(d, al, dh, idesc, 1);
          bh += 32;  // 16 taps
          bl += 32;
          ah += 8;
          al += 8;
        }
        tc::mma_commit(&x_empty[buf]);
        tc::mma_commit(&acc_full[buf]);
      }
      __syncwarp();
    }
  } else if (warp < 8) {
    // ---------------- epilogue: TMEM -> scale -> staging -> bulk stores -------------------------------------
    // Two warps per TMEM lane quarter, 56 columns each: the per-tile chain (TMEM load, 56 scaled stores) of a
    // single warp per quarter was longer than the tile's MMAs (ncu r02d: tensor pipe 65 % active, epilogue warps
    // never waiting on acc_full).
    const int m = tid & 127, half = tid >> 7, cl = m >> 3, r = m & 7;
    const int col0 = (U_NJ / 2) * half;
    const uint32_t lane_base = (uint32_t)(32 * (warp & 3)) << 16;
    const float wi = winv[cl];
    for (int64_t it = 0; it < my_tiles; ++it) {
      const int buf = (int)(it & 1);
      const uint32_t par = (uint32_t)((it >> 1) & 1);
      const int64_t tile = slot + it * per_blk;
      const int64_t b = tile / a.tiles_per_row;
      const int64_t t0 = (tile - b * a.tiles_per_row) * U_TT;
      uint8_t* sbuf = reinterpret_cast<uint8_t*>(stage) + (int)(it % U_NSTAGE) * (U_FB * U_PITCH * 4);
      uint8_t* srow = sbuf + cl * PITCHB + (8 * col0 + r) * ESZ;
      if constexpr (!RG)
        if (warp == 0 && lane < U_FB) tc::bulk_wait_read<U_NSTAGE - 1>();  // stores that read this buffer U_NSTAGE tiles ago
      tc::named_bar_sync(1, 256);
      ok = ok && tc::mbar_wait(&acc_full[buf], par);
      tc::fence_after_sync();
      const float sc = xinv[it & 3] * wi;      // read before acc_empty is released (the slot is reused 4 tiles later)
      const int64_t nvalid = T - t0 - r;       // column j is a real output iff 8j < nvalid
      const uint32_t d = tmem + lane_base + (uint32_t)(buf * U_TM_D1) + col0;
      uint32_t v[64];                          // 56 columns are this warp's; the load granule is 32
      if (!FABL(0)) {
        tc::tmem_ld32(d, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
        tc::tmem_ld32(d + 32, *reinterpret_cast<uint32_t(*)[32]>(&v[32]));
        tc::tmem_ld_wait();
      }
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&acc_empty[buf]);   // accumulator may be overwritten
      float gs[4] = {0.f, 0.f, 0.f, 0.f}, gq[4] = {0.f, 0.f, 0.f, 0.f};  // GroupNorm sums, 4 independent chains
      if (!FABL(0))
#pragma unroll
      for (int jj = 0; jj < U_NJ / 2; ++jj) {
        const float yv = __uint_as_float(v[jj]) * sc;
        if constexpr (GN) {
          const float yg = 8 * (col0 + jj) < nvalid ? yv : 0.f;
          gs[jj & 3] += yg;
          gq[jj & 3] = fmaf(yg, yg, gq[jj & 3]);
        }
        if constexpr (OB)
          *reinterpret_cast<__nv_bfloat16*>(srow + 8 * jj * ESZ) = __float2bfloat16_rn(yv);
        else
          *reinterpret_cast<float*>(srow + 8 * jj * ESZ) = yv;
      }
      if constexpr (GN) {                      // warp = 4 filters x 8 phases x 56 columns; a group of 8 filters = 4 warps
        float s = (gs[0] + gs[1]) + (gs[2] + gs[3]), q = (gq[0] + gq[1]) + (gq[2] + gq[3]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          s += __shfl_xor_sync(0xffffffffu, s, o);
          q += __shfl_xor_sync(0xffffffffu, q, o);
        }
        if (lane == 0) {
          gnred[2 * warp] = s;
          gnred[2 * warp + 1] = q;
        }
      }
      tc::fence_async_smem();
      tc::named_bar_sync(1, 256);
      if constexpr (GN) {
        if ((tid & 63) == 0 && tid < 128) {    // warps 0 and 2: groups 0 and 1 of this CTA's 16 filters
          const int gl = tid >> 6;
          if (c0 + 8 * gl < C) {
            const int64_t g = c0 / 8 + gl;
            const int w0 = 2 * gl;             // warps w0, w0+1 (columns 0..55) and w0+4, w0+5 (columns 56..111)
            a.gn_partial[(b * (C / 8) + g) * a.tiles_per_row + t0 / U_TT] =
                make_float2((gnred[2 * w0] + gnred[2 * w0 + 2]) + (gnred[2 * w0 + 8] + gnred[2 * w0 + 10]),
                            (gnred[2 * w0 + 1] + gnred[2 * w0 + 3]) + (gnred[2 * w0 + 9] + gnred[2 * w0 + 11]));
          }
        }
      }
      if constexpr (RG) {
        // plain stores: warp w writes rows 2w and 2w + 1, a 128-byte (fp32) line per instruction; the staging buffer is
        // free again when the loop is done (the barrier at the top of the next tile orders the reuse)
        const int n = (int)((T - t0) < U_TT ? (T - t0) : U_TT);
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
          const int row = 2 * warp + rr;
          const int c = c0 + row;
          if (c < C) {
            using elem = typename std::conditional<OB, uint16_t, uint32_t>::type;
            const elem* src = reinterpret_cast<const elem*>(sbuf + row * PITCHB);
            elem* dst = reinterpret_cast<elem*>(a.y) + ((b * C + c) * T + t0);
            for (int i = lane; i < n; i += 32) dst[i] = src[i];
          }
        }
      } else if (warp == 0 && lane < U_FB) {
        const int row = lane;
        const int c = c0 + row;
        if (c < C && !FABL(1)) {
          const int64_t n = (T - t0) < U_TT ? (T - t0) : U_TT;
          tc::bulk_store(reinterpret_cast<uint8_t*>(a.y) + ((b * C + c) * T + t0) * ESZ, sbuf + row * PITCHB,
                         (uint32_t)(n * ESZ));
        }
        tc::bulk_commit();
      }
    }
    if constexpr (!RG)
      if (warp == 0 && lane < U_FB) tc::bulk_wait<0>();
  }
  (void)ok;  // mbar_wait traps on a timeout (sinc_tc.cuh): no silent failure path
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 4) tc::tmem_dealloc(tmem, 512);
  if (blockIdx.x == 0 && threadIdx.x == 0 && a.phase) {
    a.phase[4] = clock64();
    a.phase[5] = (long long)tc::global_timer_ns();
  }
}

bool tc16_fwd_supported(int C, int K, int64_t T, const void* y, int y_dtype, const DeviceInfo& di) {
  (void)C;
  (void)T;   // any length: rows that are not 16-byte aligned take the plain-store epilogue (template flag RG)
  (void)y;
  return K + 7 <= U_KS && (y_dtype == SINC_DTYPE_F32 || y_dtype == SINC_DTYPE_BF16) && di.max_smem_optin >= U_SM_TOTAL + 1024;
}
// bulk stores need every row of y to start on a 16-byte boundary and to be a multiple of 16 bytes long
static bool fwd_rows_aligned(int64_t T, const void* y, int y_dtype) {
  return (y_dtype == SINC_DTYPE_F32 ? T % 4 == 0 : T % 8 == 0) && reinterpret_cast<uintptr_t>(y) % 16 == 0;
}

int64_t tc16_fwd_tiles_per_row(int64_t T) { return ceil_div64(T, U_TT); }

cudaError_t launch_fwd_tc16(const Bank& bank, const float* x, int64_t x_stride_b, void* y, int y_dtype, int64_t B,
                            int64_t T, const DeviceInfo& di, cudaStream_t st, float* gn_partial, long long* phase) {
  FwdTc16Args a;
  a.bank = bank;
  a.x = x;
  a.xsb = x_stride_b;
  a.y = y;
  a.B = B;
  a.T = T;
  a.nblk = (bank.C + U_FB - 1) / U_FB;
  a.ksteps = (bank.K + 7 + 15) / 16;
  a.tiles_per_row = ceil_div64(T, U_TT);
  a.ntiles = B * a.tiles_per_row;
  a.phase = phase;  // sinc_options.phase_stamps (diagnostic, normally null)
  a.gn_partial = reinterpret_cast<float2*>(gn_partial);
#ifdef SINC_ABLATION
  a.ablate = getenv("SINC_ABLATE_FW") ? atoi(getenv("SINC_ABLATE_FW")) : 0;
#endif
  int64_t per_blk = di.sm_count / a.nblk;
  if (per_blk < 1) per_blk = 1;
  if (per_blk > a.ntiles) per_blk = a.ntiles;
  const int grid = (int)(per_blk * a.nblk);
  const size_t smem = U_SM_TOTAL + 1024;  // slack for the 1024-byte alignment of the dynamic segment
  const bool ragged = !fwd_rows_aligned(T, y, y_dtype);
  auto kern = y_dtype == SINC_DTYPE_BF16
                  ? (ragged ? fwd_tc16_kernel<true, false, true> : fwd_tc16_kernel<true, false, false>)
                  : (gn_partial ? (ragged ? fwd_tc16_kernel<false, true, true> : fwd_tc16_kernel<false, true, false>)
                                : (ragged ? fwd_tc16_kernel<false, false, true> : fwd_tc16_kernel<false, false, false>));
  if (gn_partial && y_dtype != SINC_DTYPE_F32) return cudaErrorInvalidValue;
  cudaError_t e = ensure_dynamic_smem(reinterpret_cast<const void*>(kern), (int)smem);
  if (e != cudaSuccess) return e;
  e = launch_pdl(kern, grid, U_THREADS, smem, st, a);
  if (e != cudaSuccess) return e;
  count_launch();
  return cudaGetLastError();
}

}  // namespace sinc
