// tcgen05 engine, grad-filter, 16-bit split operands:
//     G[c,k] = sum_b sum_t g[b,c,t] x[b,t+k-H]                          (ConvolutionBackward0, grad-weight)
// bf16 hi/mid split (three kind::f16 MMAs per product, fp32 accumulate), sm_100a only.
// Output: one partial tile per CTA (layout 2 of sinc_finalize.cuh); for one filter block (C <= 64) the kernel reduces the
// tiles and applies the synthesis chain rule in its own tail (see "fused finalize" at the end of the kernel), otherwise
// finalize_kernel does.
//
// Precision: 16 significant bits per operand, ~5e-6 max-norm relative on G against the fp64 oracle (gate 1e-4);
// the 3xTF32 kernel (sinc_tc_bwd_filter.cu) stays available as SINC_ENGINE_TC_TF32.
//
// Formulation.  The contraction runs over time, 16 samples per K-step.  The B operand is the waveform read IN
// PLACE through a Toeplitz descriptor (a K-major no-swizzle operand with LBO = 16, SBO = 128 is a matrix with a
// row pitch of 16 bytes = 8 bf16 samples):  B_rho[n, kappa] = xs_rho[8n + kappa + 16q],  xs_rho[i] = x[tb-H+i+rho],
// four copies shifted by rho = 0..3 samples so that every start address stays 16-byte aligned.  Row (c,h) of the
// A operand holds the 16 samples g[c, tb + 16q - 4h + kappa], h = 0,1 -- two time interleaves per filter give
// M = 128 rows from 64 filters -- and then
//     D_rho[(c,h), n]  accumulates  G[c, k]  with  k = 8n + rho + 4h :
// the interleave offset of 4 samples makes the two rows of a filter produce COMPLEMENTARY taps (k mod 8 in 0..3
// and 4..7), so four shifts x N = 32 columns cover K <= 256 taps with no redundant MMA work (the 3xTF32 kernel
// computes every tap in both interleaves: 78 % useful).  Each g sample is converted twice (once per interleave).
//
// A (= grad_y) is staged as raw fp32 rows by 1-D bulk copies (2 stages x 64 rows x 256 samples = 128 KB in flight: 1 KB
// per copy -- 512-byte copies are capped at 5.2 TB/s by the issue rate of cp.async.bulk), read back row-per-lane, split
// into bf16 hi and NEGATED mid (tc::split_bf16x2_negmid: four instructions per pair; the MMA that consumes the mid piece
// negates A in its instruction descriptor) and written with tcgen05.st into a TMEM ring of 4 slots x 4 K-steps (one
// producer/issuer handshake per slot); every MMA is the TS form (A from TMEM), the B tiles are 1 KB each.
//     TMEM  two accumulator sets [0,128) / [128,256): D_rho = 32 columns  |  A ring [256,512): 16 columns per K-step
//     SMEM  grad_y staging ring 132 KB  |  2 x 4 x (hi,mid) shifted waveform tiles, 26 KB  |  finalize scratch 13 KB
// The accumulators are flushed every F2_FLUSH tiles (read-modify-write of the CTA's own partial tile in the
// workspace), which bounds the length of any fp32 accumulation chain; see the issuer for why the flush does not
// overlap the next window's MMAs.
#include <cuda_bf16.h>

#include <cstdlib>
#include <type_traits>

#include <atomic>
#include <cstdio>
#include <chrono>

#include "sinc_finalize.cuh"
#include "sinc_internal.h"
#include "sinc_synth.cuh"
#include "sinc_tc.cuh"

namespace sinc {

constexpr int F2_THREADS = 512;  // warps 0-7 grad_y converters (2 groups), 8-9 MMA issuers, 10 waveform producer,
                                 // 11 grad_y bulk-copy issuer, 12-15 flush
constexpr int F2_N = 32;                    // MMA N (tap columns per shift rho)
constexpr int F2_TILE = 512;                // samples per tile = 32 K-steps of 16 samples
constexpr int F2_XLEN = 832;                // bf16 elements per shifted waveform copy: >= 8*31 + 16*36
constexpr int F2_KPS = 4;                   // K-steps per A ring slot (one producer/issuer handshake per slot)
constexpr int F2_RING = 4;                  // TMEM A ring depth (slots of F2_KPS K-steps = 16*F2_KPS columns)
static_assert(8 % F2_KPS == 0 && (F2_RING & (F2_RING - 1)) == 0 && F2_RING * F2_KPS * 16 <= 256, "A ring geometry");
// chunks (ring slots) per staging stage: must be even -- chunk ch of a tile belongs to converter group ch & 1
#ifndef F2_FLUSH_TILES
#define F2_FLUSH_TILES 8
#endif
constexpr int F2_FLUSH = F2_FLUSH_TILES;    // tiles per accumulation window (4096 samples = 256 K-steps).  The tensor core's fp32
                                            // accumulation truncates, so a chain of n accumulations is biased by ~n * 2^-24: this
                                            // is most of the 1.3e-5 error of G (5.7e-6 with 1024-sample windows, where the 16-bit
                                            // operand split is what is left) and the reason two shardings of one batch differ by
                                            // ~1.4e-5 in the parameter gradients.  Shorter windows were measured and cost more than
                                            // they are worth: a flush is ~6 us with the issuers waiting for it, ~2.6 us overlapped
                                            // (profiles/r02_grad_filter_window_length.log: 2 tiles: +30 % / +13 % of the kernel).
constexpr int F2_DCOLS = 4 * F2_N;          // 128
constexpr int F2_GSAMP = 256;               // samples per grad_y staging stage (16 K-steps = four ring slots): one bulk copy moves
                                            // 1 KB of a row.  With 512-byte copies the single issuing thread caps the feed at
                                            // 5.2 TB/s chip-wide whatever the depth (one UBLKCP per ~24 cycles); 1 KB copies with
                                            // 128 KB in flight stream at 6.5 TB/s (tools/feed_probe.py, profiles/r02_feed_probe.log)
constexpr int F2_SPS = F2_GSAMP / 16;       // K-steps per staging stage
constexpr int F2_GPITCH = F2_GSAMP * 4 + 32;  // bytes per staged row: 32 mod 128, so the row-per-lane LDS.128 of a
                                              // quarter-warp (4 filters x 2 interleaves) hit 8 distinct bank groups
constexpr int F2_GSTAGES = 2;               // staging ring depth: 128 KB of grad_y in flight per SM
constexpr int F2_GSTAGE_BYTES = 64 * F2_GPITCH;
constexpr int F2_SM_G = 0;                  // [F2_GSTAGES][64 rows][F2_GPITCH]
constexpr int F2_SM_X = F2_SM_G + F2_GSTAGES * F2_GSTAGE_BYTES;  // [buf 2][rho 4][hi/mid 2][F2_XLEN] bf16
constexpr int F2_SM_BAR = F2_SM_X + 2 * 8 * F2_XLEN * 2;
constexpr int F2_SM_FIN = F2_SM_BAR + 512;  // fused finalize (doubles): us[256] chi[256] clo[256] Gs[256] part[2][256] scratch[64] ChanCoef
constexpr int F2_SM_TOTAL = F2_SM_FIN + (6 * 256 + 64 + 4) * 8;   // + the filter's ChanCoef
static_assert(28 + F2_GSTAGES <= 32, "barrier table: tmem_slot lives at bars + 32");
static_assert((F2_XLEN * 2) % 16 == 0, "waveform copies must stay 16-byte aligned");

// grad_y element type: fp32 (split into bf16 hi + mid) or bf16 (bf16-output mode: already exact, no mid piece and
// one MMA fewer per product).  `vec` = 4 consecutive samples of a staged row.
template <bool GB>
struct F2G;
template <>
struct F2G<false> {
  using vec = float4;
  static constexpr int esz = 4;
  static constexpr int pitch = F2_GPITCH;
  __device__ static vec zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
};
template <>
struct F2G<true> {
  using vec = uint2;
  static constexpr int esz = 2;
  static constexpr int pitch = F2_GSAMP * 2 + 16;  // 16 mod 128: the LDS.64 of a half-warp (8 filters x 2 interleaves)
                                                   // cover 16 distinct 8-byte slots
  __device__ static vec zero() { return make_uint2(0u, 0u); }
};

struct BwdFilterTc16Args {
  const float* x;
  int64_t xsb;
  const void* g;
  int64_t B, T;
  int C, K;
  float* partials;  // [gridDim.x / nblk][nblk][128][F2_DCOLS]
  int nblk;         // filter blocks of 64
  int64_t tiles_per_row;
  int64_t ntiles;
  long long* phase;  // diagnostic time stamps (CTA 0), may be null
  Bank bank;         // only read when tab != nullptr
  uint8_t* tab;      // filter table for the grad-input kernel (sinc_synth.cuh), or nullptr
  // Fused finalize (one filter block, C <= 64): after its last flush every CTA announces its partial tile on a tagged
  // counter in the workspace; CTA c < C waits for all of them, reduces column c of every tile in fp64 in a fixed order
  // and applies the synthesis chain rule (sinc_finalize.cuh) -- no second launch, no second ramp-up.
  unsigned long long* fin_sync;  // {tag : 40 | arrivals : 24}, or nullptr: the caller launches finalize_kernel instead
  unsigned long long fin_tag;    // unique per launch (host counter): stale contents of the word never match it
  float* grad_low;
  float* grad_band;
  float* grad_filters;           // debug output [C][K], may be null
#ifdef SINC_ABLATION
  int ablate;        // experiment build only (tools/gf_overlap_repro.py): bit 3 = the issuers do NOT wait for the previous
                     // window's flush (the overlap under investigation), bit 4 = the flush warps release the accumulator
                     // set only after their global-memory read-modify-write has completed (membar)
#endif
};
#ifdef SINC_ABLATION
#define GABL(bit) ((a.ablate >> (bit)) & 1)
#else
#define GABL(bit) 0
#endif

__device__ __forceinline__ void f2_split_bf16x2(float v0, float v1, uint32_t& hi, uint32_t& mid) {
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(v1), "f"(v0));
  const float h0 = __uint_as_float(hi << 16), h1 = __uint_as_float(hi & 0xffff0000u);
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(mid) : "f"(v1 - h1), "f"(v0 - h0));
}

// K-steps of a tile: the h = 1 interleave lags by 4 samples, so the last tile of an utterance needs the step that
// covers its final 4 samples.
__device__ __forceinline__ int f2_ksteps(int64_t left, bool last) {
  const int span = (int)(left < F2_TILE ? left : F2_TILE) + (last ? 4 : 0);
  return (span + 15) / 16;
}

template <bool GB>
__global__ void __launch_bounds__(F2_THREADS, 1) bwd_filter_tc16_kernel(const BwdFilterTc16Args a) {
  using G = F2G<GB>;
  using gvec = typename G::vec;
  constexpr int GPITCH = G::pitch;
  extern __shared__ __align__(1024) uint8_t smem[];
  if (blockIdx.x == 0 && threadIdx.x == 0 && a.phase) {
    a.phase[0] = clock64();
    a.phase[1] = (long long)tc::global_timer_ns();
  }
#ifdef SINC_ABLATION
  const uint64_t dbg_start = tc::global_timer_ns();
#endif
  uint16_t* xs = reinterpret_cast<uint16_t*>(smem + F2_SM_X);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + F2_SM_BAR);
  uint64_t* a_full = bars;            // [F2_RING]  producers (4 warps of one group) -> issuers
  uint64_t* a_empty = bars + 8;       // [F2_RING]  issuers (2 commits) -> producers
  uint64_t* x_full = bars + 16;       // [2]
  uint64_t* x_empty = bars + 18;      // [2]  issuers (2 commits)
  uint64_t* d_full = bars + 20;       // [2]  issuers (2 commits) -> flush warps
  uint64_t* d_empty = bars + 22;      // [2]  flush warps (4) -> issuers
  uint64_t* g_full = bars + 24;       // [F2_GSTAGES]  bulk copies (expect_tx) -> converters
  uint64_t* g_empty = bars + 28;      // [F2_GSTAGES]  converters (8 warps) -> copy issuer
  uint64_t* fin_bar = bars + 12;      // fused finalize: bulk copies of the partial tiles -> the whole CTA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 32);  // after g_empty[F2_GSTAGES]
  uint8_t* gst = smem + F2_SM_G;
  double* fin_us = reinterpret_cast<double*>(smem + F2_SM_FIN);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int C = a.C, H = (a.K - 1) / 2;
  const int per_blk = gridDim.x / a.nblk;
  const int fblk = blockIdx.x % a.nblk;
  const int slot = blockIdx.x / a.nblk;
  const int c0 = fblk * 64;
  const int64_t T = a.T;

  if (tid == 0) {
    for (int i = 0; i < F2_RING; ++i) {
      tc::mbar_init(&a_full[i], 4);
      tc::mbar_init(&a_empty[i], 2);
    }
    for (int i = 0; i < 2; ++i) {
      tc::mbar_init(&x_full[i], 1);
      tc::mbar_init(&x_empty[i], 2);
      tc::mbar_init(&d_full[i], 2);
      tc::mbar_init(&d_empty[i], 4);
    }
    for (int i = 0; i < F2_GSTAGES; ++i) {
      tc::mbar_init(&g_full[i], 1);
      tc::mbar_init(&g_empty[i], 8);
    }
    tc::mbar_init(fin_bar, 1);
    tc::mbar_init_fence();
  }
  if (warp == 8) {
    tc::tmem_alloc(tmem_slot, 512);
    tc::tmem_relinquish();
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  tc::griddep_wait();     // everything above touched shared memory and TMEM only
  tc::griddep_launch();

  // contiguous range of tiles for this CTA (tile = 512 samples of one utterance)
  const int64_t tile_begin = a.ntiles * slot / per_blk;
  const int64_t tile_end = a.ntiles * (slot + 1) / per_blk;
  const int64_t my_tiles = tile_end - tile_begin;
  bool ok = true;
  if (blockIdx.x == 0 && threadIdx.x == 0 && a.phase) {
    a.phase[2] = clock64();
    a.phase[3] = (long long)tc::global_timer_ns();
  }

  if (warp < 8) {
    // ------------- grad_y converters (staged fp32 rows -> bf16 hi/mid -> TMEM ring) --------------------------
    // Two groups of four warps (one warp per TMEM lane quarter each) take alternate K-steps.  The rows come from
    // the bulk-copy ring: the register pipeline of the first version only covered ~800 cycles of load latency
    // (ncu r01t: converters stalled on first use, issuers idle on a_full); 4 stages x 32 KB of asynchronous
    // copies are in flight per SM instead, and no register holds data that has not arrived yet.
    const int grp = warp >> 2;
    const int quarter = warp & 3;   // tcgen05.ld/st: a warp reaches TMEM lanes 32*(warp%4) .. +31 only
    const int m = 32 * quarter + lane;  // TMEM lane = A row (c_local, h)
    const int cl = m >> 1, h = m & 1;
    const int c = c0 + cl;
    const bool row_valid = c < C;
    const uint32_t lane_base = (uint32_t)(32 * quarter) << 16;
    const gvec zero4 = G::zero();
    gvec halo = zero4;  // the 4 samples before the current stage (h = 1 rows start 4 samples early)
    const bool all_rows = __all_sync(0xffffffffu, row_valid);
    const uint8_t* gbytes = reinterpret_cast<const uint8_t*>(a.g);
    int64_t cc0 = 0;      // ring slots (chunks of F2_KPS K-steps) before the current tile
    int64_t gs = 0;       // staging stages consumed so far
    auto emit = [&](const gvec (&v)[4 * F2_KPS], int64_t cc) {  // F2_KPS K-steps of this row -> one TMEM ring slot
      const int s = (int)(cc & (F2_RING - 1));
      const uint32_t par = (uint32_t)((cc / F2_RING) & 1);
      uint32_t hi[8 * F2_KPS], mid[8 * F2_KPS];
#pragma unroll
      for (int j = 0; j < 4 * F2_KPS; ++j) {
        if constexpr (GB) {  // bf16 pairs are already packed in memory order (even sample in the low half)
          hi[2 * j] = v[j].x;
          hi[2 * j + 1] = v[j].y;
          mid[2 * j] = mid[2 * j + 1] = 0u;
        } else {
          // the second piece is stored NEGATED (sinc_tc.cuh: 4 instead of 6 instructions per pair); its MMA negates A back
          tc::split_bf16x2_negmid(v[j].x, v[j].y, hi[2 * j], mid[2 * j]);
          tc::split_bf16x2_negmid(v[j].z, v[j].w, hi[2 * j + 1], mid[2 * j + 1]);
        }
      }
      ok = ok && tc::mbar_wait(&a_empty[s], par ^ 1);
      tc::fence_after_sync();
      const uint32_t dst = tmem + lane_base + 2 * F2_DCOLS + 16 * F2_KPS * s;
#pragma unroll
      for (int kk = 0; kk < F2_KPS; ++kk) {  // per K-step: 8 columns hi, 8 columns mid
        uint32_t th[8], tm8[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          th[e] = hi[8 * kk + e];
          tm8[e] = mid[8 * kk + e];
        }
        tc::tmem_st8(dst + 16 * kk, th);
        if constexpr (!GB) tc::tmem_st8(dst + 16 * kk + 8, tm8);
      }
      tc::tmem_st_wait();
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&a_full[s]);
    };
    for (int64_t it = 0; it < my_tiles; ++it) {
      const int64_t tile = tile_begin + it;
      const int64_t b = tile / a.tiles_per_row;
      const int64_t ti = tile - b * a.tiles_per_row;
      const int64_t tb = ti * F2_TILE;
      const int64_t left = T - tb;
      const int span = (int)(left < F2_TILE ? left : F2_TILE);
      const int nq = f2_ksteps(left, ti + 1 == a.tiles_per_row);
      const int nstg = (span + F2_GSAMP - 1) / F2_GSAMP;
      if (ti == 0) {
        halo = zero4;  // utterance start: t < 0
      } else if (it == 0) {  // this CTA starts in the middle of an utterance: fetch the 4 samples before its first tile
        halo = (h && row_valid) ? __ldg(reinterpret_cast<const gvec*>(gbytes + ((b * C + c) * T + tb - 4) * G::esz)) : zero4;
      }
      for (int st = 0; st < nstg; ++st) {
        const int gb = (int)(gs & (F2_GSTAGES - 1));
        ok = ok && tc::mbar_wait(&g_full[gb], (uint32_t)((gs / F2_GSTAGES) & 1));
        const uint8_t* rowp = gst + gb * F2_GSTAGE_BYTES + cl * GPITCH;
        const int64_t ts = tb + (int64_t)F2_GSAMP * st;
#pragma unroll 1
        for (int u = 0; u < F2_SPS / F2_KPS / 2; ++u) {
          const int pch = grp + 2 * u;                 // chunk of the stage; chunk ch of the tile belongs to group ch & 1
          const int ch = (F2_SPS / F2_KPS) * st + pch;
          if (F2_KPS * ch < nq) {
            gvec v[4 * F2_KPS];
            // Interior stage of a full filter block (the common case): no per-element range checks -- the 64-bit
            // compares and selects of the general path were a third of the converters' instructions (ncu r04c),
            // and the converters are co-critical with the MMAs.  (t < 0 never needs a check: the halo is zero at
            // the start of an utterance.)
            const bool interior = all_rows && ts + F2_GSAMP <= T;
            auto fetch = [&](int i) -> gvec {
              const int off = 16 * F2_KPS * pch - 4 * h + 4 * i;  // sample offset within the stage (-4 only for pch = 0, h = 1)
              if (pch == 0 && i == 0) return h ? halo : *reinterpret_cast<const gvec*>(rowp);
              return *reinterpret_cast<const gvec*>(rowp + G::esz * off);
            };
            if (interior) {
#pragma unroll
              for (int i = 0; i < 4 * F2_KPS; ++i) v[i] = fetch(i);
            } else {
#pragma unroll
              for (int i = 0; i < 4 * F2_KPS; ++i) {
                const int64_t tj = ts + (16 * F2_KPS * pch - 4 * h + 4 * i);
                const gvec w = fetch(i);
                v[i] = (row_valid && tj >= 0 && tj < T) ? w : zero4;
              }
            }
            emit(v, cc0 + ch);
          }
        }
        if (h) {  // carry the last 4 samples of the row into the next stage
          const gvec w = *reinterpret_cast<const gvec*>(rowp + G::esz * (F2_GSAMP - 4));
          halo = (row_valid && ts + F2_GSAMP - 4 < T) ? w : zero4;
        }
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&g_empty[gb]);
        ++gs;
      }
      if (nq > F2_SPS * nstg && grp == 0) {  // the K-step after the last stage: only the carried samples are non-zero
        gvec v[4 * F2_KPS];
#pragma unroll
        for (int i = 0; i < 4 * F2_KPS; ++i) v[i] = zero4;
        if (h) v[0] = halo;
        emit(v, cc0 + (F2_SPS / F2_KPS) * nstg);
      }
      cc0 += (nq + F2_KPS - 1) / F2_KPS;
    }
  } else if (warp < 10) {
    // ------------- MMA issuers: warp 8 owns the shifts rho = 0,1, warp 9 rho = 2,3 -------------------------------
    // All lanes run the loop (warp-uniform control flow and operands); only the elected lane issues.
    // The issue loop is instruction-bound (an N = 32 MMA occupies the tensor pipe for 16 cycles; tools/
    // mma_rate_pattern.py reaches that rate from a lean instruction stream): everything the loop needs is kept
    // provably warp-uniform -- derived from kernel parameters, blockIdx and compile-time constants, the issuer's
    // identity is a template constant -- so that the compiler holds it in uniform registers, and the shared-memory
    // descriptors are advanced by additions instead of being rebuilt from addresses.
    auto issue = [&](auto wtag) {
      constexpr int w = decltype(wtag)::value;
      const int ntl = (int)my_tiles;
      const int tpr = (int)a.tiles_per_row;
      const int tile0 = (int)(tile_begin % a.tiles_per_row);
      const uint32_t idesc = tc::make_idesc_f16(128, F2_N, tc::kFmtBF16, tc::kFmtBF16);
      const uint32_t idesc_na = idesc | tc::kIdescNegA;    // the mid piece of grad_y sits in TMEM negated (see the converters)
      const uint64_t bdesc = tc::make_smem_desc(16, 128);  // in-place Toeplitz view, row pitch 8 samples
      const uint32_t xs0 = tc::smem_addr(xs);
      const uint32_t tm = tc::uniform((int)tmem);
      int cc = 0;          // ring slots consumed so far
      int ti = tile0;      // tile index within its utterance
      uint32_t acc = 0;    // 0: the next K-step overwrites the accumulators
      for (int it = 0; it < ntl; ++it) {
        const int64_t left = T - (int64_t)ti * F2_TILE;
        const int nq = f2_ksteps(left, ti + 1 == tpr);
        if (++ti == tpr) ti = 0;
        const int xb = it & 1;
        const int win = it / F2_FLUSH;
        const int dbuf = win & 1;
        ok = ok && tc::mbar_wait(&x_full[xb], (uint32_t)((it >> 1) & 1));
        if (it % F2_FLUSH == 0) {
          // this accumulator set was handed to the flush warps two windows ago: wait until it has been read out
          ok = ok && tc::mbar_wait(&d_empty[dbuf], (uint32_t)(((win >> 1) & 1) ^ 1));
          // ... and until the PREVIOUS window's set has been read out as well.  Letting the flush of one set overlap
          // the MMAs into the other produced rare, run-to-run varying errors in rows of the last lane quarter
          // (tools/gf_debug.py, r02k: 3e-3 on a few filters with > 8 tiles per CTA; never without the overlap).
          // The root cause is not understood yet, so the issuer stalls for the ~3 us a flush takes (once per
          // F2_FLUSH tiles, ~3 % of the kernel) -- the two sets still let the flush overlap the converters.
          if (win >= 1 && !GABL(3)) ok = ok && tc::mbar_wait(&d_empty[dbuf ^ 1], (uint32_t)(((win - 1) >> 1) & 1));
          acc = 0;
        }
        const uint32_t d0 = tm + F2_DCOLS * dbuf + F2_N * (2 * w);
        uint32_t b0 = xs0 + xb * (8 * F2_XLEN * 2) + (2 * (2 * w)) * F2_XLEN * 2;  // rho = 2w: hi copy, mid copy follows
        const int nch = (nq + F2_KPS - 1) / F2_KPS;  // K-steps past nq carry zeros in A: harmless
        for (int ch = 0; ch < nch; ++ch) {
          const int s = cc & (F2_RING - 1);
          ok = ok && tc::mbar_wait(&a_full[s], (uint32_t)((cc / F2_RING) & 1));
          tc::fence_after_sync();
          const uint64_t dsc0 = tc::desc_at(bdesc, b0);
          if (tc::elect_one()) {
#pragma unroll
            for (int kk = 0; kk < F2_KPS; ++kk) {
              const uint32_t ah = tm + 2 * F2_DCOLS + 16 * F2_KPS * s + 16 * kk, am = ah + 8;
              // consecutive MMAs into the same accumulator serialise on its TMEM read-modify-write: alternate the
              // two shifts this issuer owns so that dependent instructions are never adjacent
              // (descriptor start-address field = bytes >> 4: a byte offset is an addition to the low word)
              const uint64_t bh0 = dsc0 + (uint64_t)((32 * kk) >> 4), bm0 = bh0 + (uint64_t)((F2_XLEN * 2) >> 4);
              const uint64_t bh1 = bh0 + (uint64_t)((2 * F2_XLEN * 2) >> 4);
              const uint64_t bm1 = bh0 + (uint64_t)((3 * F2_XLEN * 2) >> 4);
              const uint32_t acc0 = kk == 0 ? acc : 1u;
              tc::mma_f16_ts(d0, ah, bh0, idesc, acc0);
              tc::mma_f16_ts(d0 + F2_N, ah, bh1, idesc, acc0);
              tc::mma_f16_ts(d0, ah, bm0, idesc, 1);
              tc::mma_f16_ts(d0 + F2_N, ah, bm1, idesc, 1);
              if constexpr (!GB) {  // bf16 grad_y has no mid piece
                tc::mma_f16_ts(d0, am, bh0, idesc_na, 1);
                tc::mma_f16_ts(d0 + F2_N, am, bh1, idesc_na, 1);
              }
            }
            tc::mma_commit(&a_empty[s]);
          }
          __syncwarp();
          acc = 1;
          b0 += 32 * F2_KPS;  // 16 samples per K-step
          ++cc;
        }
        if (tc::elect_one()) {
          tc::mma_commit(&x_empty[xb]);
          if ((it + 1) % F2_FLUSH == 0 || it + 1 == ntl) tc::mma_commit(&d_full[dbuf]);
        }
        __syncwarp();
      }
    };
    if (warp == 8)
      issue(std::integral_constant<int, 0>{});
    else
      issue(std::integral_constant<int, 1>{});
  } else if (warp == 10) {
    // ------------- waveform producer: four shifted bf16 hi/mid copies of the tile ---------------------------
    // xs_rho[i] = x[tb - H + i + rho].  Lane handles element pairs: even shifts store the pair (i, i+1) at word
    // (i - rho)/2, odd shifts the pair (i+1, i+2) -- every store is one aligned 32-bit word.
    for (int64_t it = 0; it < my_tiles; ++it) {
      const int64_t tile = tile_begin + it;
      const int64_t b = tile / a.tiles_per_row;
      const int64_t tb = (tile - b * a.tiles_per_row) * F2_TILE;
      const int xb = (int)(it & 1);
      const float* xr = a.x + b * a.xsb;
      constexpr int NP = (F2_XLEN / 2 + 2 + 31) / 32;  // pair slots per lane (covers indices up to F2_XLEN + 3)
      float v0[NP], v1[NP], v2[NP];
#pragma unroll
      for (int r = 0; r < NP; ++r) {
        const int i = 2 * (lane + 32 * r);
        const int64_t t = tb - H + i;
        v0[r] = (t >= 0 && t < T) ? __ldg(xr + t) : 0.f;
        v1[r] = (t + 1 >= 0 && t + 1 < T) ? __ldg(xr + t + 1) : 0.f;
        v2[r] = (t + 2 >= 0 && t + 2 < T) ? __ldg(xr + t + 2) : 0.f;
      }
      ok = ok && tc::mbar_wait(&x_empty[xb], (uint32_t)(((it >> 1) & 1) ^ 1));
      uint32_t* base = reinterpret_cast<uint32_t*>(xs + (size_t)xb * 8 * F2_XLEN);
#pragma unroll
      for (int r = 0; r < NP; ++r) {
        const int i = 2 * (lane + 32 * r);
        uint32_t he, me, ho, mo;
        f2_split_bf16x2(v0[r], v1[r], he, me);  // pair (i, i+1)
        f2_split_bf16x2(v1[r], v2[r], ho, mo);  // pair (i+1, i+2)
#pragma unroll
        for (int rho = 0; rho < 4; ++rho) {
          // even rho: word index (i - rho)/2 holds (i, i+1);  odd rho: word (i + 1 - rho)/2 holds (i+1, i+2)
          const int wi = (rho & 1) ? (i + 1 - rho) / 2 : (i - rho) / 2;
          const bool okw = ((rho & 1) ? (i + 1 - rho) : (i - rho)) >= 0 && wi < F2_XLEN / 2;
          if (okw) {
            base[(2 * rho) * (F2_XLEN / 2) + wi] = (rho & 1) ? ho : he;
            base[(2 * rho + 1) * (F2_XLEN / 2) + wi] = (rho & 1) ? mo : me;
          }
        }
      }
      tc::fence_async_smem();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&x_full[xb]);
    }
  } else if (warp == 11) {
    // ------------- grad_y bulk-copy issuer: one 512-byte row copy per filter and stage (TMA engine, 1-D) --------
    // cp.async.bulk takes uniform operands: the loop over rows is warp-uniform and one elected lane issues (a
    // per-lane loop costs an ELECT/R2UR round trip per copy -- ncu r01x: 3000 cycles per stage, converters starved).
    {
      const int ntl = tc::uniform((int)my_tiles);
      const int tpr = tc::uniform((int)a.tiles_per_row);
      int ti = tc::uniform((int)(tile_begin % a.tiles_per_row));
      int64_t brow = tile_begin / a.tiles_per_row;  // utterance of the current tile (uniform: from blockIdx only)
      const int nrows = (C - c0) < 64 ? (C - c0) : 64;
      const uint32_t gst0 = tc::uniform((int)tc::smem_addr(gst));
      int gs = 0;
      for (int it = 0; it < ntl; ++it) {
        const int64_t tb = (int64_t)ti * F2_TILE;
        const int64_t left = T - tb;
        const int span = (int)(left < F2_TILE ? left : F2_TILE);
        const int nstg = (span + F2_GSAMP - 1) / F2_GSAMP;
        const uint8_t* src_tile = reinterpret_cast<const uint8_t*>(a.g) + ((brow * C + c0) * T + tb) * G::esz;
        for (int st = 0; st < nstg; ++st) {
          const int gb = gs & (F2_GSTAGES - 1);
          ok = ok && tc::mbar_wait(&g_empty[gb], (uint32_t)(((gs / F2_GSTAGES) & 1) ^ 1));
          const int nsamp = span - F2_GSAMP * st < F2_GSAMP ? span - F2_GSAMP * st : F2_GSAMP;
          const uint32_t bytes = (uint32_t)nsamp * G::esz;  // T % 4 == 0 (fp32) / T % 8 == 0 (bf16): a multiple of 16
          if (tc::elect_one()) {
            tc::mbar_expect_tx(&g_full[gb], (uint32_t)nrows * bytes);
            const uint8_t* src = src_tile + (int64_t)F2_GSAMP * st * G::esz;
            uint32_t dst = gst0 + gb * F2_GSTAGE_BYTES;
            const uint32_t bar = tc::smem_addr(&g_full[gb]);
#pragma unroll 4
            for (int r = 0; r < nrows; ++r) {
              asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                           "l"(src), "r"(bytes), "r"(bar)
                           : "memory");
              src += T * G::esz;
              dst += GPITCH;
            }
          }
          __syncwarp();
          ++gs;
        }
        if (++ti == tpr) {
          ti = 0;
          ++brow;
        }
      }
    }
  } else if (warp >= 12) {
    // ------------- flush warps: accumulator set -> the CTA's partial tile (fixed order, no atomics) ------------
    // They have nothing to do until the first window is complete: first they leave the synthesised filters in the
    // workspace for the grad-input kernel that follows on this stream (sinc_synth.cuh, "filter table").
    if (a.tab) {
      if (blockIdx.x == 0)
        for (int i = tid - 384; i < kTabHdrWords; i += 128)
          reinterpret_cast<uint32_t*>(a.tab)[i] = tab_header_word(a.bank, i);
      for (int c = 4 * blockIdx.x + (warp - 12); c < 64; c += 4 * gridDim.x) tab_write_row(a.bank, a.tab, c, lane);
    }
    if (a.fin_sync && (int)blockIdx.x < C) {  // the G-independent factors of this CTA's filter (read in the tail, below)
      const ChanCoef cc = chan_coef(a.bank, blockIdx.x);
      if (tid == 384) *reinterpret_cast<ChanCoef*>(fin_us + 6 * 256 + 64) = cc;
      finalize_factors(a.bank, cc, tid - 384, 128, fin_us, fin_us + 256, fin_us + 512);
    }
    const int quarter = warp & 3;
    const int m = 32 * quarter + lane;
    const uint32_t lane_base = (uint32_t)(32 * quarter) << 16;
    float* myrow = a.partials + (((size_t)slot * a.nblk + fblk) * 128 + m) * F2_DCOLS;
    const int64_t nwin = (my_tiles + F2_FLUSH - 1) / F2_FLUSH;
    for (int64_t win = 0; win < nwin; ++win) {
      const int dbuf = (int)(win & 1);
      ok = ok && tc::mbar_wait(&d_full[dbuf], (uint32_t)((win >> 1) & 1));
      tc::fence_after_sync();
#pragma unroll 1
      for (int ch = 0; ch < F2_DCOLS / 32; ++ch) {
        uint32_t v[32];
        tc::tmem_ld32(tmem + lane_base + F2_DCOLS * dbuf + 32 * ch, v);
        tc::tmem_ld_wait();
        float4* p = reinterpret_cast<float4*>(myrow + 32 * ch);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float4 o = make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]),
                                 __uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3]));
          if (win > 0) {
            const float4 q = p[i];
            o.x += q.x; o.y += q.y; o.z += q.z; o.w += q.w;
          }
          p[i] = o;
        }
      }
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&d_empty[dbuf]);
    }
    if (my_tiles == 0) {  // idle CTA: its partial tile must still read as zero
      for (int i = 0; i < F2_DCOLS; i += 4) *reinterpret_cast<float4*>(myrow + i) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  (void)ok;  // mbar_wait traps on a timeout (sinc_tc.cuh): no silent failure path
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 8) tc::tmem_dealloc(tmem, 512);
  if (blockIdx.x == 0 && threadIdx.x == 0 && a.phase) {
    a.phase[4] = clock64();
    a.phase[5] = (long long)tc::global_timer_ns();
  }
  if (!a.fin_sync) return;
#ifdef SINC_ABLATION
  const uint64_t dbg_t0 = tc::global_timer_ns();
  uint64_t dbg_t1 = 0, dbg_t2 = 0, dbg_ta = 0, dbg_tb = 0, dbg_tc = 0;
  if (tid == 0 && GABL(6)) {
    uint32_t smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    printf("cta %d sm %u start %llu mainend %llu\n", (int)blockIdx.x, smid, (unsigned long long)dbg_start, (unsigned long long)dbg_t0);
  }
#endif

  // ------------- fused finalize ------------------------------------------------------------------------------------
  // The partial tile of this CTA is complete (the __syncthreads above orders the flush warps' stores before thread 0's
  // fence).  Arrival is a CAS on {tag, count}: whatever the word held before this launch -- zero, or the garbage of a
  // fresh workspace -- is replaced by {tag, 1} by the first CTA to arrive.  Finalising CTAs count themselves a second
  // time once they are past the wait, and the last of them writes zero: a replay of the same launch (CUDA graph, same
  // tag) starts from a word that does not carry its tag.
  const unsigned long long cnt_mask = (1ull << 24) - 1;
  const unsigned long long nfin = (unsigned long long)(C < (int)gridDim.x ? C : (int)gridDim.x);
  if (tid == 0) {
    __threadfence();
    // Fast path: the word already carries this launch's tag -> one fire-and-forget add.  (A CAS per arrival turns 148
    // CTAs finishing within microseconds of each other into a retry storm on one L2 line: measured 35 us.)  Only the
    // first arrivals of a launch find a foreign word and race to replace it; the losers see the tag and add.
    unsigned long long old = *reinterpret_cast<volatile unsigned long long*>(a.fin_sync);
    for (;;) {
      if ((old >> 24) == a.fin_tag) {
        atomicAdd(a.fin_sync, 1ull);
        break;
      }
      const unsigned long long prev = atomicCAS(a.fin_sync, old, (a.fin_tag << 24) | 1ull);
      if (prev == old) break;
      old = prev;
    }
  }
  if ((int)blockIdx.x >= C) return;   // CTAs beyond the filter count are done: their SMs go to the next kernel
  if (tid == 0) {
    const uint64_t t0 = tc::global_timer_ns();
    for (;;) {
      unsigned long long v;
      asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(a.fin_sync) : "memory");
      if ((v >> 24) == a.fin_tag && (v & cnt_mask) >= (unsigned long long)gridDim.x) break;
      __nanosleep(64);
      if (tc::global_timer_ns() - t0 > tc::kWaitLimitNs) __trap();   // loud, like every other wait of this kernel
    }
    __threadfence();
#ifdef SINC_ABLATION
    dbg_t1 = tc::global_timer_ns();
#endif
  }
  __syncthreads();
  {
    double* us = fin_us;
    double* chi = fin_us + 256;
    double* clo = fin_us + 512;
    double* Gs = fin_us + 768;
    double* part = fin_us + 1024;   // [2][256]
    double* scratch = fin_us + 1536;
    const int K = a.K, slots = gridDim.x;
    uint32_t fin_parity = 0;
    for (int c = blockIdx.x; c < C; c += gridDim.x) {
      // the filter's scalars were computed next to its factors at the start of the kernel (parameter loads and an fp64
      // division are ~1.5k cycles of latency that would sit on the critical path here)
      ChanCoef cc = *reinterpret_cast<const ChanCoef*>(fin_us + 6 * 256 + 64);
      if (c != (int)blockIdx.x) {   // grids smaller than the filter count: later filters get their factors here
        cc = chan_coef(a.bank, c);
        __syncthreads();
        if (tid >= 384) finalize_factors(a.bank, cc, tid - 384, 128, us, chi, clo);
      }
      // The filter's two rows of every partial tile (1 KB, contiguous) come into shared memory with one bulk copy per
      // slot, all in flight at once: ONE L2 round trip instead of a chain of dependent load batches (measured 1.1 - 1.7 us
      // per batch of 16 - 32 loads per thread, 5.5 - 8 us in total).  The staging ring and the waveform buffers are idle now.
      // Thread (q, e) then sums element e over the slots q, q + 2, ... in eight fixed fp64 chains: bitwise reproducible.
      // tile layout (see the flush warps): row (c, h), column rho * 32 + n holds tap k = 8n + rho + 4h
      const int q = tid >> 8, e = tid & 255;
      const int k = 8 * (e & 31) + ((e >> 5) & 3) + 4 * (e >> 7);
      const float* src = a.partials + (size_t)(2 * c) * F2_DCOLS;
      const size_t stride = (size_t)128 * F2_DCOLS;
      const bool staged = slots * 1024 <= F2_SM_BAR;
      if (staged) {
        // generic -> async proxy: the partial tiles were written with ordinary stores (made visible by the arrival protocol
        // above), the copies below read them through the async proxy; shared memory changes hands as well.  The copies
        // may complete before thread 0 has posted the byte count: the transaction count of an mbarrier may run negative.
        if (tid == 0) tc::mbar_expect_tx(fin_bar, (uint32_t)slots * 1024u);
        if (tc::elect_one()) {
          asm volatile("fence.proxy.async.global;" ::: "memory");
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          const uint32_t bar = tc::smem_addr(fin_bar);
          for (int sl = warp; sl < slots; sl += F2_THREADS / 32)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             tc::smem_addr(smem) + (uint32_t)sl * 1024u),
                         "l"(src + (size_t)sl * stride), "r"(1024u), "r"(bar)
                         : "memory");
        }
#ifdef SINC_ABLATION
        if (tid == 0) dbg_ta = tc::global_timer_ns();
#endif
        ok = ok && tc::mbar_wait(fin_bar, fin_parity);
        fin_parity ^= 1u;
#ifdef SINC_ABLATION
        if (tid == 0) dbg_tb = tc::global_timer_ns();
#endif
      }
      if (k < K) {
        double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0, s4 = 0.0, s5 = 0.0, s6 = 0.0, s7 = 0.0;
        if (staged) {
          const float* p = reinterpret_cast<const float*>(smem) + e;
#pragma unroll 1
          for (int sl = q; sl < slots; sl += 16) {
            float v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = sl + 2 * i < slots ? p[(sl + 2 * i) * 256] : 0.f;
            s0 += (double)v[0];
            s1 += (double)v[1];
            s2 += (double)v[2];
            s3 += (double)v[3];
            s4 += (double)v[4];
            s5 += (double)v[5];
            s6 += (double)v[6];
            s7 += (double)v[7];
          }
        } else {   // more partial tiles than shared memory holds (not on a 148-SM part): L2 loads (ld.global.cg), never .nc
          const float* p = src + e;
#pragma unroll 1
          for (int sl = q; sl < slots; sl += 16) {
            float v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = sl + 2 * i < slots ? __ldcg(p + (size_t)(sl + 2 * i) * stride) : 0.f;
            s0 += (double)v[0];
            s1 += (double)v[1];
            s2 += (double)v[2];
            s3 += (double)v[3];
            s4 += (double)v[4];
            s5 += (double)v[5];
            s6 += (double)v[6];
            s7 += (double)v[7];
          }
        }
        part[q * 256 + k] = ((s0 + s1) + (s2 + s3)) + ((s4 + s5) + (s6 + s7));
      }
#ifdef SINC_ABLATION
      if (tid == 0) dbg_tc = tc::global_timer_ns();
#endif
      __syncthreads();
      if (tid < K) {
        const double s = part[tid] + part[256 + tid];
        Gs[tid] = s;
        if (a.grad_filters) a.grad_filters[(size_t)c * K + tid] = (float)s;
      }
      __syncthreads();
#ifdef SINC_ABLATION
      if (tid == 0) dbg_t2 = tc::global_timer_ns();
#endif
      finalize_chain_rule<F2_THREADS>(a.bank, cc, c, Gs, us, chi, clo, scratch, a.grad_low, a.grad_band);
    }
  }
  if (tid == 0) {   // past the wait and done: count this CTA a second time; the last one leaves the word at zero (off the critical path)
    const unsigned long long before = atomicAdd(a.fin_sync, 1ull);
    if ((before & cnt_mask) == (unsigned long long)gridDim.x + nfin - 1) atomicExch(a.fin_sync, 0ull);
  }
  if (blockIdx.x == 0 && tid == 0 && a.phase) a.phase[7] = (long long)tc::global_timer_ns();   // end of the fused finalize
#ifdef SINC_ABLATION
  if (tid == 0 && GABL(5) && (blockIdx.x == 0 || blockIdx.x == 63))
    printf("gf tail cta %d: wait %llu ns, reduce %llu ns (issue %llu, land %llu, sum %llu, combine %llu), chain rule %llu ns\n", (int)blockIdx.x,
           (unsigned long long)(dbg_t1 - dbg_t0), (unsigned long long)(dbg_t2 - dbg_t1),
           (unsigned long long)(dbg_ta - dbg_t1), (unsigned long long)(dbg_tb - dbg_ta), (unsigned long long)(dbg_tc - dbg_tb),
           (unsigned long long)(dbg_t2 - dbg_tc),
           (unsigned long long)(tc::global_timer_ns() - dbg_t2));
#endif
}

bool tc16_bwd_filter_supported(int C, int K, int64_t T, const void* g, int g_dtype, const DeviceInfo& di) {
  (void)C;
  const bool shape_ok = g_dtype == SINC_DTYPE_F32 ? T % 4 == 0 : (g_dtype == SINC_DTYPE_BF16 && T % 8 == 0);
  return K <= 256 && shape_ok && reinterpret_cast<uintptr_t>(g) % 16 == 0 && di.max_smem_optin >= F2_SM_TOTAL + 1024;
}

size_t tc16_bwd_filter_partial_bytes(int C) {
  const int nblk = (C + 63) / 64;
  return (size_t)kMaxPartialSlots * nblk * 128 * F2_DCOLS * sizeof(float);
}

cudaError_t launch_bwd_filter_tc16(const Bank& bank, const float* x, int64_t x_stride_b, const void* gy, int gy_dtype,
                                   float* ws_partials, int* slots_used, int64_t B, int64_t T, const DeviceInfo& di,
                                   cudaStream_t st, uint8_t* filter_table, long long* phase, int reserved_sms,
                                   const FusedFinalize* fin, bool* fused_out) {
  BwdFilterTc16Args a;
  a.x = x;
  a.xsb = x_stride_b;
  a.g = gy;
  a.B = B;
  a.T = T;
  a.C = bank.C;
  a.K = bank.K;
  a.partials = ws_partials;
  a.nblk = (bank.C + 63) / 64;
  a.tiles_per_row = ceil_div64(T, F2_TILE);
  a.ntiles = B * a.tiles_per_row;
  a.phase = phase ? phase + 12 : nullptr;  // sinc_options.phase_stamps (diagnostic, normally null)
  a.bank = bank;
  a.tab = (bank.C <= 64 && bank.K <= 256) ? filter_table : nullptr;
  a.fin_sync = nullptr;
  a.fin_tag = 0;
  a.grad_low = a.grad_band = a.grad_filters = nullptr;
  if (fin && fin->sync_word && a.nblk == 1) {
    // launch tags: a process-wide counter started from the clock, 40 bits -- never repeats within a process, and neither
    // zero-filled nor small-integer garbage in a fresh workspace looks like one
    static std::atomic<unsigned long long> next_tag{
        (unsigned long long)std::chrono::steady_clock::now().time_since_epoch().count() * 0x9E3779B97F4A7C15ull};
    a.fin_sync = fin->sync_word;
    a.fin_tag = (next_tag.fetch_add(1) & ((1ull << 39) - 1)) | (1ull << 39);   // never zero
    a.grad_low = fin->grad_low;
    a.grad_band = fin->grad_band;
    a.grad_filters = fin->grad_filters;
  }
#ifdef SINC_ABLATION
  a.ablate = getenv("SINC_ABLATE_GF") ? atoi(getenv("SINC_ABLATE_GF")) : 0;
#endif
  int64_t per_blk = (di.sm_count - reserved_sms) / a.nblk;   // reserved_sms: SMs left to a kernel running next to this one
  if (per_blk < 1) per_blk = 1;
  if (per_blk > a.ntiles) per_blk = a.ntiles;
  if (per_blk > kMaxPartialSlots) per_blk = kMaxPartialSlots;
  const int grid = (int)(per_blk * a.nblk);
  const size_t smem = F2_SM_TOTAL + 1024;
  auto kern = gy_dtype == SINC_DTYPE_BF16 ? bwd_filter_tc16_kernel<true> : bwd_filter_tc16_kernel<false>;
  cudaError_t e = ensure_dynamic_smem(reinterpret_cast<const void*>(kern), (int)smem);
  if (e != cudaSuccess) return e;
  e = launch_pdl(kern, grid, F2_THREADS, smem, st, a);
  if (e != cudaSuccess) return e;
  count_launch();
  *slots_used = (int)per_blk;
  if (fused_out) *fused_out = a.fin_sync != nullptr;
  return cudaGetLastError();
}

}  // namespace sinc
