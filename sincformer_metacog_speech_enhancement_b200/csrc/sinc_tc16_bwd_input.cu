// tcgen05 engine, grad-input, 16-bit split operands:
//     gx[b,t] = sum_c sum_k w[c,k] g[b,c,t-k+H]                        (ConvolutionBackward0, grad-input)
// as a GEMM over (tap parity, filter) + a skewed column sum (col2im), bf16 hi/mid split, sm_100a only.
//
// Precision.  Every fp32 operand is split v = hi + mid, hi = bf16(v), mid = bf16(v - hi) (16 significant bits);
// three kind::f16 MMAs (hi*hi + hi*mid + mid*hi) accumulate in fp32.  Measured against the fp64 oracle the result
// is ~6e-6 max-norm relative -- 16x inside the 1e-4 gradient gate -- at twice the tensor-core rate of the 3xTF32
// kernel (sinc_tc_bwd_input.cu, kept as the SINC_ENGINE_TC_TF32 variant).
//
// Formulation.  Taps are paired, k = 2k' + e.  With  D[k', p] = sum_e sum_c w[c, 2k'+e] * g[c, p-e]  every
// element of D belongs to ONE output, t = p + 2k' - H, so the GEMM does half of the col2im additions itself:
//     A = W'[k', (e,c)] = w[c, 2k'+e]      128 tap pairs x K = 2*64, resident in TMEM (hi and mid) for the whole kernel
//     B = G [p , (e,c)] = g[c, p-e]        positions x K, K-major in shared memory
// B is stored ONCE per position as 8 "planes" of 16-byte rows (8 filters each): a K-major no-swizzle operand
// with SBO = 128 is a plain row-pitch-16-byte matrix, so the e = 1 half of K is the same memory read one row
// earlier (start address - 16 bytes): no duplication, no im2col.  Shared memory holds no filters at all.
//
// col2im (per epilogue warp: 32 tap pairs L x 64 positions c of the accumulator; cell (L, c) belongs to window cell
// o = c + 2L).  The accumulator is read with the 16x256b TMEM load shape: thread t = 4r + q4 receives, for the four rows
// L = r + 8m (m = 0..3), the columns c = 8j + 2 q4 + e (j = 0..7, e = 0..1) -- rows EIGHT LANES APART land in the same
// thread, and 2 * 8m = 16m is a multiple of the thread's column period 8, so the four rows are added in registers with
// no cross-thread traffic at all:
//     E[r][8 j' + 2 q4 + e] = sum_m D[r + 8m][8 (j' - 2m) + 2 q4 + e],   j' = 0..13          (36 FADD per thread)
// leaving 8 rows x 112 columns per warp (896 values instead of 2048) for the part that must cross threads:
//   store   thread (r, q4) writes its 14 float2 to S[r][14 + 8 j' + 2 q4]        (14 STS.64, row pitch 152: conflict free)
//   gather  lane l sums S[r][14 + (l + 32 i) - 2r] over r = 0..7 for its four cells o = l + 32 i   (32 LDS.32; the 14-word
//           margins on both sides of a row are zero, so no cell needs a range check)
// i.e. 60 shared-memory wavefronts per 2048 accumulator values -- the previous version (one row per thread, 32x32b
// loads) needed 132 and was bound by exactly that traffic (ncu r05a: LSU data pipe + MMA operand reads = 93 % of the
// shared-memory pipe).  The fragment layout is pinned on hardware by tests/test_gpu_tc.py::test_tmem_load_16x256b_fragment.
// The eight warps' 128-cell windows and the carry from earlier stages are combined in a fixed order: bitwise
// reproducible.
//
// Pipeline: 4 producer warps (TMA tensor boxes -> registers -> bf16 hi/mid -> planes, for fp32 and bf16 grad_y alike;
// 2-slot operand ring, the last row of a stage is also the halo row of the next), 1 MMA issuer warp
// (the two 64-column halves of a stage one after the other), double-buffered accumulators in TMEM, 8 epilogue warps,
// 3 combiner warps.  A CTA owns a contiguous range of 128-sample output blocks and recomputes the 2H-column halo at
// the start of each run.  The filters come from the workspace table of the grad-filter kernel when it is valid
// (sinc_synth.cuh), else every CTA synthesises them.
#include <cuda.h>
#include <cuda_bf16.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <type_traits>

#include "sinc_internal.h"
#include "sinc_synth.cuh"
#include "sinc_tc.cuh"

namespace sinc {

constexpr int H2_THREADS = 512;  // warps 0-7 epilogue (quarter = w%4, column half = w/4), 8-11 producers, 12 MMA issuer,
                                 // 13-15 combiners (windows + carry -> gx)
// setmaxnreg re-partitions the registers the CTA was LAUNCHED with (512 threads x 128), not the whole file:
constexpr int H2_REGS_WORK = 152;  // epilogue and producer warpgroups
constexpr int H2_REGS_ISSUE = 40;  // issuer warpgroup
static_assert(3 * 128 * H2_REGS_WORK + 128 * H2_REGS_ISSUE <= H2_THREADS * 128, "setmaxnreg.inc would block forever");
constexpr int H2_NS = 128;       // D columns (positions) per stage
constexpr int H2_SLOTS = 2;      // B ring depth (load latency is hidden by the staging boxes, not by the ring)
constexpr int H2_ACCS = 3;       // accumulator buffers in TMEM (128 columns each)
constexpr int H2_TM_W = 128 * H2_ACCS;  // first TMEM column of W' (hi: 64 columns, mid: the next 64)
static_assert(H2_TM_W + 128 <= 512, "TMEM budget");
constexpr int H2_PLANE = (H2_NS + 1) * 16;       // bytes: halo row + 128 rows of 8 filters x bf16
constexpr int H2_SLOT = 2 * 8 * H2_PLANE;        // hi planes [8], mid planes [8]
constexpr int H2_WIN = 384;                      // output window of one stage: 128 positions + 2*127 + 1 (3 combiner warps x 4 cells)
static_assert(H2_WIN == 4 * 96, "combiner cell ownership");
constexpr int H2_EPITCH = 152;                   // floats per row of an epilogue warp's col2im buffer: 14 zero | 112 | 26 zero;
                                                 // 152 = 24 mod 32: the STS.64 of a half-warp (4 rows x 8 words) hit 32 banks
constexpr int H2_SKEW = 8 * H2_EPITCH;           // floats per epilogue warp: 8 rows
constexpr int H2_STG_BOX = 64 * 32 * 4;                        // one TMA box: 64 filter rows x 32 positions, fp32, 128B-swizzled
constexpr int H2_SM_STG = 0;                                   // staging: [4 producer warps][2 boxes]
constexpr int H2_NBOX = 2;                                     // TMA boxes in flight per producer warp: 64 KB of grad_y per SM
                                                               // (3 measured slower, r02h: 0.103 vs 0.097 ms for the bare feed)
                                                               // (an L2 tensor prefetch six stages ahead changed nothing either)
constexpr int H2_SM_G = H2_SM_STG + 4 * H2_NBOX * H2_STG_BOX;  // B ring; in the prologue: fp32 filters [64][256]
constexpr int H2_SM_S = H2_SM_G + H2_SLOTS * H2_SLOT;          // col2im buffers: one per epilogue warp (margins stay zero)
constexpr int H2_SM_OUT = H2_SM_S + 8 * H2_SKEW * 4;           // 2 x 8 x window [128] (double-buffered by stage parity)
constexpr int H2_SM_CARRY = H2_SM_OUT + 2 * 8 * 128 * 4;       // 2 x carry [256]
constexpr int H2_SM_BAR = H2_SM_CARRY + 2 * 256 * 4;
constexpr int H2_SM_TOTAL = H2_SM_BAR + 512;
static_assert(64 * 256 * 4 <= H2_SLOTS * H2_SLOT, "filter staging must fit in the B ring");
static_assert(H2_SM_S % 16 == 0 && H2_SM_OUT % 16 == 0 && H2_SM_G % 1024 == 0, "alignment");
static_assert(H2_SM_TOTAL + 1024 <= 227 * 1024, "shared memory budget");
static_assert(21 + 4 * H2_NBOX <= 64, "barrier table");

struct BwdInputTc16Args {
  Bank bank;
  const void* g;           // fp32, or bf16 (bf16-output mode: exact in one bf16 piece -- no mid planes, one MMA fewer)
  float* gx;
  int64_t B, T;
  int64_t blocks_per_row;  // ceil(T / 128)
  int64_t nblocks;         // B * blocks_per_row
  long long* phase;        // diagnostic time stamps (CTA 0), may be null
  const uint8_t* tab;      // filter table left in the workspace by the grad-filter kernel (sinc_synth.cuh), or nullptr
  // More than 64 filters: one launch per block of 64 (`bank`, `g` and the tensor map describe the block), every launch
  // after the first adds to gx.  c_total = filter rows per utterance in grad_y (the batch stride).
  int c_total;
  int accumulate;
#ifdef SINC_ABLATION
  int ablate;              // ablation build only (tools/ablate.py): bit 0 epilogue skips col2im, 1 epilogue skips the TMEM
                           // load too, 2 producers skip conversion + plane stores, 3 producers skip the staging reads too,
                           // 4 the issuer issues no MMAs (commits only), 5 every TMA box reads the same few (L2-resident)
                           // stages: a "perfect feed".  Results are garbage by construction.
#endif
};
#ifdef SINC_ABLATION
#define ABL(bit) ((a.ablate >> (bit)) & 1)
#else
#define ABL(bit) 0
#endif

// v0, v1 -> packed bf16 pairs (v1 in the high half): hi = RN(v), mid = RN(v - hi)
__device__ __forceinline__ void split_bf16x2(float v0, float v1, uint32_t& hi, uint32_t& mid) {
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(v1), "f"(v0));
  const float h0 = __uint_as_float(hi << 16), h1 = __uint_as_float(hi & 0xffff0000u);
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(mid) : "f"(v1 - h1), "f"(v0 - h0));
}

template <bool GB>
__global__ void __launch_bounds__(H2_THREADS, 1)
    bwd_input_tc16_kernel(const BwdInputTc16Args a, const __grid_constant__ CUtensorMap gmap) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (tc::smem_addr(smem_raw) & 1023u)) & 1023u);  // swizzled boxes: 1024-byte aligned
  uint8_t* ring = smem + H2_SM_G;
  float* skew = reinterpret_cast<float*>(smem + H2_SM_S);
  float* outw = reinterpret_cast<float*>(smem + H2_SM_OUT);
  float* carry = reinterpret_cast<float*>(smem + H2_SM_CARRY);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + H2_SM_BAR);
  uint64_t* b_full = bars;         // [4]  producers (4 warps) -> issuers
  uint64_t* b_empty = bars + 4;    // [4]  issuer (1 commit per stage) -> producers
  uint64_t* d_full = bars + 8;     // [H2_ACCS accumulator buffers]  issuer -> epilogue
  uint64_t* d_empty = bars + 12;   // [H2_ACCS]  epilogue (8 warps) -> issuer
  uint64_t* win_full = bars + 16;  // [2]  epilogue (8 warps) -> combiners: the windows of a stage are written
  uint64_t* win_empty = bars + 18; // [2]  combiners (3 warps) -> epilogue: the window buffer may be rewritten
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20);
  uint64_t* st_full = bars + 21;   // [4 producer warps][H2_NBOX]  TMA -> the warp that issued it

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int C = a.bank.C, K = a.bank.K, H = (K - 1) / 2;
  const int ncb = (C + 15) / 16;  // 16-filter K-steps per tap parity
  const int Hr = (H + 3) & ~3;    // H rounded up to a multiple of 4: stage columns stay 16-byte aligned in g
  const int64_t T = a.T;

  // ---------------- setup ------------------------------------------------------------------------------------
  if (blockIdx.x == 0 && threadIdx.x == 0 && a.phase) {
    a.phase[0] = clock64();
    a.phase[1] = (long long)tc::global_timer_ns();
  }
  if (tid == 0) {
    for (int i = 0; i < 4; ++i) {
      tc::mbar_init(&b_full[i], 4);
      tc::mbar_init(&b_empty[i], 1);
      tc::mbar_init(&d_full[i], 1);
      tc::mbar_init(&d_empty[i], 8);
    }
    for (int i = 0; i < 2; ++i) {
      tc::mbar_init(&win_full[i], 8);
      tc::mbar_init(&win_empty[i], 3);
    }
    for (int i = 0; i < 4 * H2_NBOX; ++i) tc::mbar_init(&st_full[i], 1);
    tc::tensormap_prefetch(&gmap);
    tc::mbar_init_fence();
  }
  if (warp == 12) {
    tc::tmem_alloc(tmem_slot, 512);
    tc::tmem_relinquish();
  }
  for (int i = tid; i < H2_SM_BAR / 16; i += H2_THREADS)
    reinterpret_cast<float4*>(smem)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  __syncthreads();
  tc::griddep_wait();     // everything above touched shared memory, TMEM and the (parameter-space) tensor map only
  tc::griddep_launch();
#ifdef SINC_ABLATION
  const long long dbg_c0 = clock64();
  long long dbg_c1 = 0;
#endif
  {
    float* wst = reinterpret_cast<float*>(ring);  // fp32 filters [c][256], rows >= C and taps >= K stay zero
    // The filters come from the table the grad-filter kernel left in the workspace when that table provably belongs
    // to this filter bank (header equal word for word, row checksums intact); otherwise every CTA synthesises them.
    bool tab_ok = false;
    if (a.tab) {
      int good = tab_shape_ok(a.bank) ? 1 : 0;
      if (good) {
        // Every load of this block is issued before the first use (the table rows, the row checksums, the header words and
        // the parameters they are compared with): ONE L2 round trip.  As a loop over rows this was four dependent round
        // trips and 10k cycles, on the critical path of every CTA.
        static_assert(H2_THREADS / 32 == 16, "four table rows per warp");
        const uint32_t* hdr = reinterpret_cast<const uint32_t*>(a.tab);
        const float4* src = reinterpret_cast<const float4*>(a.tab + kTabDataOff);
        const unsigned long long* rowsum = reinterpret_cast<const unsigned long long*>(a.tab + kTabRowsumOff);
        float4 v[4][2];
        unsigned long long rs[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int c = warp + 16 * i;
          v[i][0] = src[c * 64 + lane];
          v[i][1] = src[c * 64 + lane + 32];
          rs[i] = rowsum[c];
        }
        uint32_t hw[2] = {0u, 0u}, bw[2] = {0u, 0u};
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const int w = tid + i * H2_THREADS;
          if (w < kTabHdrWords) {
            hw[i] = hdr[w];
            bw[i] = tab_header_word(a.bank, w);
          }
        }
        static_assert(kTabHdrWords <= 2 * H2_THREADS, "two header words per thread");
        good &= (hw[0] == bw[0] && hw[1] == bw[1]) ? 1 : 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int c = warp + 16 * i;
          unsigned long long sum = 0ull;
#pragma unroll
          for (int h2 = 0; h2 < 2; ++h2) {
            const int q = lane + 32 * h2;
            const float4 f = v[i][h2];
            reinterpret_cast<float4*>(wst)[c * 64 + q] = f;
            sum += tab_mix(f.x, 4 * q) + tab_mix(f.y, 4 * q + 1) + tab_mix(f.z, 4 * q + 2) + tab_mix(f.w, 4 * q + 3);
          }
          good &= (warp_sum_u64(sum) == rs[i]) ? 1 : 0;
        }
      }
      tab_ok = __syncthreads_and(good) != 0;
      if (!tab_ok) {  // rejected: whatever was copied goes away again
        for (int i = tid; i < 64 * 256 / 4; i += H2_THREADS)
          reinterpret_cast<float4*>(wst)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        __syncthreads();
      }
    }
    if (!tab_ok)
      for (int c = warp; c < C; c += H2_THREADS / 32)
        synth_filter_warp(a.bank, c, lane, [&](int k, double w) { wst[c * 256 + k] = (float)w; });
    if (blockIdx.x == 0 && tid == 0 && a.phase) a.phase[12] = tab_ok ? 1 : 0;   // diagnostic: which source was used
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
#ifdef SINC_ABLATION
  dbg_c1 = clock64();
#endif
  const uint32_t tmem = *tmem_slot;
  // TMEM map: [0,384) THREE accumulators of 128 columns; [384,448) W' hi, [448,512) W' mid (column = e*32 + c/2).
  // Three, not two: the accumulator of stage s is free again when the epilogue has read it out of TMEM, which takes
  // longer than the MMAs of one stage (a barrier wake-up on each side of 64 KB of tcgen05.ld); with two buffers the
  // issuer waited for that every stage (r02e: 2500 cycles per stage against 1536 of MMAs).
  {   // all 16 warps: four per TMEM lane quarter, two of the eight (tap parity, 16-filter group) column blocks each
    const float* wst = reinterpret_cast<const float*>(ring);
    const int kp = 32 * (warp & 3) + lane;  // tap pair = TMEM lane
    const uint32_t lane_base = (uint32_t)(32 * (warp & 3)) << 16;
#pragma unroll 1
    for (int blk = warp >> 2; blk < 8; blk += 4) {
      const int e = blk >> 2, cg = blk & 3;
      {
        uint32_t hi[8], mid[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          // fp32 grad_y: plane cq holds its filter pairs rotated by (cq & 3) -- the order in which the
          // producers read them conflict-free out of the swizzled staging boxes; W' follows the same order
          const int cq = 2 * cg + (j >> 2);
          const int rot = GB ? (cq & 3) : 2 * (cq & 3);   // see the producers: rotation of the filter order per plane
          const int ca = 8 * cq + ((2 * (j & 3) + rot) & 7), cb = 8 * cq + ((2 * (j & 3) + 1 + rot) & 7);
          split_bf16x2(wst[ca * 256 + 2 * kp + e], wst[cb * 256 + 2 * kp + e], hi[j], mid[j]);
        }
        tc::tmem_st8(tmem + lane_base + H2_TM_W + 32 * e + 8 * cg, hi);
        tc::tmem_st8(tmem + lane_base + H2_TM_W + 64 + 32 * e + 8 * cg, mid);
      }
    }
    tc::tmem_st_wait();
  }
  tc::fence_before_sync();
  __syncthreads();  // the filter staging area becomes the B ring
  tc::fence_after_sync();
#ifdef SINC_ABLATION
  if (ABL(6) && tid == 0 && (blockIdx.x == 0 || blockIdx.x == 100))
    printf("bi prologue cta %d: table %lld cycles, W' build %lld cycles\n", (int)blockIdx.x, dbg_c1 - dbg_c0, clock64() - dbg_c1);
#endif

  // ---------------- work: a contiguous range of 128-sample output blocks, cut into per-utterance runs ----------
  // Every role walks the same sequence of (run, stage) pairs.  The three roles are separate top-level branches
  // that never merge before the end of the kernel, so each keeps the register budget setmaxnreg gave it.
  if (blockIdx.x == 0 && threadIdx.x == 0 && a.phase) {
    a.phase[2] = clock64();
    a.phase[3] = (long long)tc::global_timer_ns();
  }
  const int64_t v0 = a.nblocks * blockIdx.x / gridDim.x;
  const int64_t v1 = a.nblocks * (blockIdx.x + 1) / gridDim.x;
  bool ok = true;
  auto for_each_run = [&](auto&& body) {
    int64_t stage = 0;  // global stage counter of this CTA (ring slots / barrier phases)
    int64_t v = v0;
    while (v < v1) {
      const int64_t b = v / a.blocks_per_row;
      const int64_t blk0 = v - b * a.blocks_per_row;
      int64_t blk1 = blk0 + (v1 - v);
      if (blk1 > a.blocks_per_row) blk1 = a.blocks_per_row;
      const int64_t s_out = blk0 * H2_NS;                                   // owned outputs [s_out, e_out)
      const int64_t e_out = blk1 * H2_NS < T ? blk1 * H2_NS : T;
      const int64_t c0 = s_out - Hr;  // first D column (position) of the run: a multiple of 4 at or before s_out - H
      const int nst = (int)((e_out + H + 1 - c0 + H2_NS - 1) / H2_NS);      // columns [s_out - H, e_out + H]
      v += blk1 - blk0;
      body(b, s_out, e_out, c0, nst, stage);
      stage += nst;
    }
  };

  if (warp >= 8 && warp < 12) {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(H2_REGS_WORK));
    // ------------- producers: g[:, p] -> bf16 hi/mid -> row (p - col0 + 1) of the slot's planes -------------------
    // Warp pw owns 32 positions of every stage; lane (fg, pg) converts 4 consecutive positions of the filters
    // [8fg, 8fg+8) and [32+8fg, 32+8fg+8) -- two 16-byte plane rows per position.  Neighbouring planes are 16 bytes
    // apart modulo 128 and lane = 4*pg + fg, so every quarter-warp of an STS.128 covers 8 distinct 16-byte bank
    // groups: conflict-free.
    // fp32 grad_y arrives through the TMA engine: each warp keeps two 8 KB boxes (64 filters x its 32 positions,
    // 128-byte swizzle, rows beyond C and positions outside [0,T) zero-filled) in flight and re-issues a box as soon
    // as it has read it into registers -- global-load latency is off the conversion path and the L1 data pipe moves
    // 128-byte wavefronts (LDS.128 of the box) instead of one wavefront per 32-byte sector of an LDG.  Lane (fg, pg)
    // reads row 8cq + ((j + 2fg) & 7), 16-byte chunk pg ^ row: 8 distinct bank groups per quarter-warp.
    // bf16 grad_y: the same with 64-byte box rows, 64-byte swizzle and LDS.64 (see the load loop below).
    const int pw = warp - 8;
    // a quarter-warp = 2 position groups x 4 planes: 8 distinct 16-byte bank groups.  bf16 grad_y: the position groups
    // of a half-warp are {0,1,4,5} / {2,3,6,7}, so that the LDS.64 of a half-warp cover all four 16-byte chunks
    const int fg = lane & 3;
    const int pg = GB ? ((lane >> 2) & 1) | (((lane >> 2) & 2) << 1) | (((lane >> 2) & 4) >> 1) : lane >> 2;
    constexpr int BOXB = GB ? H2_STG_BOX / 2 : H2_STG_BOX;   // bytes of one box: 64 rows x 32 positions
    const int n0 = 32 * pw + 4 * pg;          // first of this thread's four columns within the stage
    const uint32_t stg0 = tc::smem_addr(smem + H2_SM_STG + pw * H2_NBOX * H2_STG_BOX);
    uint64_t* my_full = st_full + H2_NBOX * pw;
    uint32_t ldq = 0;                                // boxes consumed so far by this warp (box = ldq % H2_NBOX)
    const bool keeps_halo = (pw == 3 && pg == 7);    // these lanes convert the last row of every stage
    uint4 hsave[2], msave[2];                        // ... which is also row 0 (the halo row) of the next stage's slot
    hsave[0] = hsave[1] = msave[0] = msave[1] = make_uint4(0u, 0u, 0u, 0u);
    for_each_run([&](int64_t b, int64_t s_out, int64_t e_out, int64_t c0, int nst, int64_t stage) {
      (void)s_out; (void)e_out;
      using gvec = typename std::conditional<GB, uint2, float4>::type;   // 4 consecutive positions of one filter row
      constexpr int ESZ = GB ? 2 : 4;
      const uint8_t* gb = reinterpret_cast<const uint8_t*>(a.g) + (b * a.c_total) * T * ESZ;
      gvec vv[16];
      auto issue_box = [&](int i, uint32_t q) {  // stage i of this run -> box q % H2_NBOX (the whole warp calls; one lane issues)
        if (tc::elect_one()) {
          uint64_t* bar = &my_full[q % H2_NBOX];
          tc::mbar_expect_tx(bar, BOXB);
          tc::tensor_load_3d(stg0 + (q % H2_NBOX) * H2_STG_BOX, &gmap,
                             (int)(c0 + (int64_t)H2_NS * (ABL(5) ? (i & 3) : i) + 32 * pw), 0, (int)(ABL(5) ? 0 : b), bar);
        }
        __syncwarp();
      };
#pragma unroll
      for (int pre = 0; pre < H2_NBOX; ++pre)
        if (pre < nst) issue_box(pre, ldq + pre);
      for (int i = 0; i < nst; ++i) {
        const int64_t sg = stage + i;
        const int slot = (int)(sg & 1);
        {
          ok = ok && tc::mbar_wait(&my_full[ldq % H2_NBOX], (ldq / H2_NBOX) & 1);
          const uint32_t box = stg0 + (ldq % H2_NBOX) * H2_STG_BOX;
          if (!ABL(3))
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int cq = fg + 4 * (j >> 3);
            if constexpr (GB) {
              // 64-byte rows, 64-byte swizzle (16-byte chunk ^= bits 7..8 of the offset = (row >> 1) & 3); rows rotated
              // by fg so that the 16 lanes of a half-warp read 16 distinct 8-byte bank pairs
              const int r7 = ((j & 7) + fg) & 7;
              const uint32_t addr = box + (8 * cq + r7) * 64 + ((((pg >> 1) ^ (r7 >> 1)) & 3) << 4) + (pg & 1) * 8;
              asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(vv[j].x), "=r"(vv[j].y) : "r"(addr) : "memory");
            } else {
              const int r7 = ((j & 7) + 2 * fg) & 7;
              const uint32_t addr = box + (8 * cq + r7) * 128 + ((pg ^ r7) << 4);
              asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                           : "=f"(vv[j].x), "=f"(vv[j].y), "=f"(vv[j].z), "=f"(vv[j].w)
                           : "r"(addr)
                           : "memory");
            }
          }
          __syncwarp();
          if (i + H2_NBOX < nst) issue_box(i + H2_NBOX, ldq + H2_NBOX);  // same box: its contents are in registers now
          ++ldq;
        }
        // Convert FIRST, wait for the slot afterwards: the slot becomes free when the MMAs of stage sg - 2 complete, and
        // from that moment the issuer is waiting for this stage -- with the conversion (64 pairs per thread) behind the
        // wait, ring slot hand-over was a ping-pong with ~800 cycles of conversion in its critical path (ncu r02c: the
        // issuer polled b_full 7 times per stage, the producers b_empty twice, tensor pipe 53 % active).  Now only the
        // sixteen 16-byte stores, the proxy fence and the arrive follow the wait.
        uint4 hq[2][4], mq[2][4];
        if (ABL(2)) {
#pragma unroll
          for (int s2 = 0; s2 < 2; ++s2)
#pragma unroll
            for (int o = 0; o < 4; ++o) hq[s2][o] = mq[s2][o] = make_uint4(0u, 0u, 0u, 0u);
        } else
#pragma unroll
        for (int s2 = 0; s2 < 2; ++s2) {
          const gvec* v8 = vv + 8 * s2;
#pragma unroll
          for (int o = 0; o < 4; ++o) {
            uint4 h, m;
            if constexpr (GB) {
              // bf16 inputs: pick the 16-bit sample o of two filter rows and pack them (even filter in the low half)
              auto pk = [&](const uint2& lo, const uint2& hi2) {
                const uint32_t wl = o < 2 ? lo.x : lo.y, wh = o < 2 ? hi2.x : hi2.y;
                return __byte_perm(wl, wh, (o & 1) ? 0x7632 : 0x5410);
              };
              h = make_uint4(pk(v8[0], v8[1]), pk(v8[2], v8[3]), pk(v8[4], v8[5]), pk(v8[6], v8[7]));
              m = make_uint4(0u, 0u, 0u, 0u);
            } else {
              auto comp = [&](const float4& f) { return o == 0 ? f.x : o == 1 ? f.y : o == 2 ? f.z : f.w; };
              // mid planes hold the NEGATED second piece (4 instead of 6 instructions per pair); its MMA negates B back
              tc::split_bf16x2_negmid(comp(v8[0]), comp(v8[1]), h.x, m.x);
              tc::split_bf16x2_negmid(comp(v8[2]), comp(v8[3]), h.y, m.y);
              tc::split_bf16x2_negmid(comp(v8[4]), comp(v8[5]), h.z, m.z);
              tc::split_bf16x2_negmid(comp(v8[6]), comp(v8[7]), h.w, m.w);
            }
            hq[s2][o] = h;
            mq[s2][o] = m;
          }
        }
        // slot sg & 1 was last read by the MMAs of stage sg - 2
        if (sg >= 2) ok = ok && tc::mbar_wait(&b_empty[slot], (uint32_t)(((sg - 2) >> 1) & 1));
        uint8_t* base = ring + slot * H2_SLOT;
        if (i == 0) {
          if (pw == 0) {  // run start: the halo row (position c0 - 1) has no previous stage; lane = filter pair
            const int64_t p = c0 - 1;
            const bool in = p >= 0 && p < T;
            const int hq8 = lane >> 2;                                                     // plane
            const int hrot = GB ? (hq8 & 3) : 2 * (hq8 & 3);
            const int hc = 8 * hq8 + ((2 * (lane & 3) + hrot) & 7);      // the pair of filters of this word (see W')
            const int hc1 = 8 * hq8 + ((2 * (lane & 3) + 1 + hrot) & 7);
            uint32_t h, m;
            if constexpr (GB) {
              const uint16_t* g16 = reinterpret_cast<const uint16_t*>(gb);
              const uint32_t w0 = (in && hc < C) ? __ldg(g16 + (int64_t)hc * T + p) : 0u;
              const uint32_t w1 = (in && hc1 < C) ? __ldg(g16 + (int64_t)hc1 * T + p) : 0u;
              h = w0 | (w1 << 16);
              m = 0u;
            } else {
              const float* g32 = reinterpret_cast<const float*>(gb);
              const float w0 = (in && hc < C) ? __ldg(g32 + (int64_t)hc * T + p) : 0.f;
              const float w1 = (in && hc1 < C) ? __ldg(g32 + (int64_t)hc1 * T + p) : 0.f;
              tc::split_bf16x2_negmid(w0, w1, h, m);
            }
            *reinterpret_cast<uint32_t*>(base + hq8 * H2_PLANE + (lane & 3) * 4) = h;
            if constexpr (!GB) *reinterpret_cast<uint32_t*>(base + (8 + hq8) * H2_PLANE + (lane & 3) * 4) = m;
          }
        } else if (keeps_halo) {
#pragma unroll
          for (int s2 = 0; s2 < 2; ++s2) {
            const int cq = fg + 4 * s2;
            if (cq < 2 * ncb) {
              *reinterpret_cast<uint4*>(base + cq * H2_PLANE) = hsave[s2];
              if constexpr (!GB) *reinterpret_cast<uint4*>(base + (8 + cq) * H2_PLANE) = msave[s2];
            }
          }
        }
        if (!ABL(2))
#pragma unroll
        for (int s2 = 0; s2 < 2; ++s2) {
          const int cq = fg + 4 * s2;
          if (cq < 2 * ncb) {
#pragma unroll
            for (int o = 0; o < 4; ++o) {
              uint8_t* dst = base + cq * H2_PLANE + (n0 + o + 1) * 16;
              *reinterpret_cast<uint4*>(dst) = hq[s2][o];
              if constexpr (!GB) *reinterpret_cast<uint4*>(dst + 8 * H2_PLANE) = mq[s2][o];
            }
            hsave[s2] = hq[s2][3];
            msave[s2] = mq[s2][3];
          }
        }
        tc::fence_async_smem();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&b_full[slot]);
      }
    });
  } else if (warp >= 12) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(H2_REGS_ISSUE));
    for_each_run([&](int64_t b, int64_t s_out, int64_t e_out, int64_t c0, int nst, int64_t stage) {
      // ------------- MMA issuers: warp 12 positions 0..63 of the stage, warp 13 positions 64..127 ----------------
      // All lanes of an issuer warp run the loop (warp-uniform operands stay in uniform registers); only the
      // elected lane issues tcgen05 instructions.
      // ONE issuer (warp 12) does the two 64-column halves of a stage one after the other and hands each to its
      // epilogue warps as soon as it is complete: the epilogue of half 0 (TMEM loads, shared-memory stores) then
      // overlaps the MMAs of half 1 instead of all eight epilogue warps hitting the LSU in the same phase.
      if (warp >= 13) {
        // ------------- combiners (warps 13-15): windows of the eight epilogue warps + carry -> gx ------------------
        // Decoupled from the epilogue warps by mbarriers: those no longer meet at a CTA-wide barrier every stage.
        // 96 threads own window cells o = ct + 96 r, r = 0..3 (384 cells); carry is double-buffered in shared memory.
        const int ct = 32 * (warp - 13) + lane;
        tc::named_bar_sync(3, 96);  // previous run's carry traffic is done
        for (int o = ct; o < 256; o += 96) carry[o] = 0.f;  // new run: nothing carried (stage 0 reads carry buffer 0)
        tc::named_bar_sync(3, 96);
        for (int i = 0; i < nst; ++i) {
          const int64_t sg = stage + i;
          const int buf = (int)(sg & 1), cb = i & 1;
          const int64_t col0 = c0 + (int64_t)H2_NS * i;  // first position of the stage
          ok = ok && tc::mbar_wait(&win_full[buf], (uint32_t)((sg >> 1) & 1));
          const float* wbase = outw + buf * (8 * 128);
          const float* cin = carry + cb * 256;
          float* cout = carry + (cb ^ 1) * 256;
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            const int o = ct + 96 * r;
            float acc = o < 256 ? cin[o] : 0.f;
#pragma unroll
            for (int w = 0; w < 8; ++w) {  // warp w's window starts at cell 64*(q_w + hf_w); fixed order
              const int rel = o - 64 * ((w & 3) + (w >> 2));
              if (rel >= 0 && rel < 128) acc += wbase[w * 128 + rel];
            }
            if (o < H2_NS) {
              const int64_t t = col0 - H + o;
              if (t >= s_out && t < e_out) a.gx[b * T + t] = a.accumulate ? a.gx[b * T + t] + acc : acc;
            } else {
              cout[o - H2_NS] = acc;
            }
          }
          __syncwarp();
          if (lane == 0) tc::mbar_arrive(&win_empty[buf]);
          tc::named_bar_sync(3, 96);  // the carry written above is read by other combiner threads in the next stage
        }
      }
      if (warp == 12) {
        const int ncbu = tc::uniform(ncb);
        const int nstu = tc::uniform(nst);
        const int stg0 = tc::uniform((int)(stage & 0x3fffffff));  // only bits 0..1 of the stage counter matter below
        const uint32_t tm = tc::uniform((int)tmem);
        const uint32_t ring0 = tc::uniform((int)tc::smem_addr(ring));
        // One N = 128 MMA per (tap parity, 16 filters, product): 24 instructions of 64 tensor-pipe cycles per stage.  The
        // earlier version issued the two 64-column halves of a stage separately (48 MMAs of 32 cycles) so that the
        // epilogue of half 0 could start early -- but one thread sustains an MMA only every ~40 cycles in this loop
        // (descriptor arithmetic on the uniform datapath + MIO back-pressure from the other warps' shared-memory and
        // TMEM traffic; ncu r02c: the UTCHMMA instructions stall on `mio`, tensor pipe 53 % active), so 32-cycle MMAs
        // left the pipe idle a third of the time.  The accumulators are double-buffered by stage, so the epilogue of
        // stage s still overlaps the MMAs of stage s + 1.
        const uint32_t idesc = tc::make_idesc_f16(128, H2_NS, tc::kFmtBF16, tc::kFmtBF16);
        const uint32_t idesc_nb = idesc | tc::kIdescNegB;          // the mid planes of grad_y are stored negated (producers)
        const uint64_t bdesc = tc::make_smem_desc(H2_PLANE, 128);  // K chunk stride = plane, 8-row groups contiguous
        for (int i = 0; i < nstu; ++i) {
          const int sg = stg0 + i;
          const int slot = sg & 1;
          const int acb = sg % H2_ACCS, acp = (sg / H2_ACCS) & 1;   // accumulator buffer and the parity of its use count
          ok = ok && tc::mbar_wait(&d_empty[acb], (uint32_t)(acp ^ 1));
          ok = ok && tc::mbar_wait(&b_full[slot], (uint32_t)((sg >> 1) & 1));
          tc::fence_after_sync();
          const uint32_t d = tm + 128 * acb;
          const uint32_t sbase = ring0 + slot * H2_SLOT;
          if (tc::elect_one()) {
            uint32_t acc = 0;
            for (int e = 0; e < (ABL(4) ? 0 : 2); ++e) {
              uint32_t ah = tm + H2_TM_W + 32 * e;
              uint32_t bh = sbase + (1 - e) * 16;  // e = 0 reads rows 1..128, e = 1 the same planes one row earlier
              for (int cb = 0; cb < ncbu; ++cb) {
                const uint64_t dh = tc::desc_at(bdesc, bh), dm = tc::desc_at(bdesc, bh + 8 * H2_PLANE);
                tc::mma_f16_ts(d, ah, dh, idesc, acc);
                if constexpr (!GB) tc::mma_f16_ts(d, ah, dm, idesc_nb, 1);  // bf16 grad_y has no mid planes
                tc::mma_f16_ts(d, ah + 64, dh, idesc, 1);
                acc = 1;
                ah += 8;
                bh += 2 * H2_PLANE;
              }
            }
            tc::mma_commit(&b_empty[slot]);
            tc::mma_commit(&d_full[acb]);
          }
          __syncwarp();
        }
      }
      __syncwarp();
    });
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(H2_REGS_WORK));
    for_each_run([&](int64_t b, int64_t s_out, int64_t e_out, int64_t c0, int nst, int64_t stage) {
      // ------------- epilogue: TMEM -> skewed column sums -> window -> gx --------------------------------------
      const int q = warp & 3, hf = warp >> 2;
      const uint32_t lane_base = (uint32_t)(32 * q) << 16;
      float* S0 = skew + warp * H2_SKEW;                                 // [8][H2_EPITCH], margins zero since the prologue
      float* srow = S0 + (lane >> 2) * H2_EPITCH + 14 + 2 * (lane & 3);   // thread (r, q4): row r, first column pair
      const float* gat = S0 + 14 + lane;                                 // lane l gathers cells o = l + 32 i
      for (int i = 0; i < nst; ++i) {
        const int64_t sg = stage + i;
        const int buf = (int)(sg & 1);
        const int acb = (int)(sg % H2_ACCS);
        ok = ok && tc::mbar_wait(&d_full[acb], (uint32_t)((sg / H2_ACCS) & 1));
        tc::fence_after_sync();
        uint32_t va[32], vb[32];   // va: TMEM lanes 0..15 of this quarter (rows r, r+8), vb: lanes 16..31 (rows r+16, r+24)
        const uint32_t dcol = tmem + lane_base + 128 * acb + 64 * hf;
        if (!ABL(1)) {
          tc::tmem_ld16x256_x8(dcol, va);
          tc::tmem_ld16x256_x8(dcol + (16u << 16), vb);
          tc::tmem_ld_wait();
        } else {
#pragma unroll
          for (int z = 0; z < 32; ++z) va[z] = vb[z] = 0u;
        }
        tc::fence_before_sync();  // all 64 columns of this half have left TMEM: release the accumulator
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&d_empty[acb]);
        // rows r + 8m, m = 0..3: element (j, e) of row m is v[4j + 2(m & 1) + e] of va (m < 2) or vb; it belongs to column
        // group j' = j + 2m of the summed row
        float c0s = 0.f, c1s = 0.f, c2s = 0.f, c3s = 0.f;
        if (!ABL(0)) {
#pragma unroll
        for (int jp = 0; jp < 14; ++jp) {
          float e0 = 0.f, e1 = 0.f;
#pragma unroll
          for (int m = 0; m < 4; ++m) {
            const int j = jp - 2 * m;
            if (j >= 0 && j < 8) {
              const uint32_t* v = m < 2 ? va : vb;
              e0 += __uint_as_float(v[4 * j + 2 * (m & 1)]);
              e1 += __uint_as_float(v[4 * j + 2 * (m & 1) + 1]);
            }
          }
          *reinterpret_cast<float2*>(srow + 8 * jp) = make_float2(e0, e1);
        }
        __syncwarp();
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          const float* row = gat + r * H2_EPITCH - 2 * r;
          c0s += row[0];
          c1s += row[32];
          c2s += row[64];
          c3s += row[96];
        }
        }
        ok = ok && tc::mbar_wait(&win_empty[buf], (uint32_t)(((sg >> 1) & 1) ^ 1));  // combined two stages ago
        float* win = outw + buf * (8 * 128) + warp * 128;
        win[lane] = c0s;
        win[32 + lane] = c1s;
        win[64 + lane] = c2s;
        win[96 + lane] = c3s;
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&win_full[buf]);
      }
    });
  }
  (void)ok;  // mbar_wait traps on a timeout (sinc_tc.cuh): no silent failure path
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 12) tc::tmem_dealloc(tmem, 512);
  if (blockIdx.x == 0 && threadIdx.x == 0 && a.phase) {
    a.phase[4] = clock64();
    a.phase[5] = (long long)tc::global_timer_ns();
  }
}

bool tc16_bwd_input_supported(int C, int K, int64_t T, const void* g, int g_dtype, const DeviceInfo& di) {
  // (the rows of the TMA tensor map must be 16-byte aligned: T % 4 for fp32, T % 8 for bf16 grad_y)
  const bool rows_ok = g_dtype == SINC_DTYPE_F32 ? T % 4 == 0 : (g_dtype == SINC_DTYPE_BF16 && T % 8 == 0);
  (void)C;  // any filter count: blocks of 64 filters, one launch each (launch_bwd_input_tc16)
  return K <= 256 && rows_ok &&
         reinterpret_cast<uintptr_t>(g) % 16 == 0 && di.max_smem_optin >= H2_SM_TOTAL + 1024;
}

namespace {
// cuTensorMapEncodeTiled through the runtime's driver entry point table: the library does not link libcuda.
using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) != cudaSuccess ||
        qr != cudaDriverEntryPointSuccess)
      p = nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}
// grad_y [B][C][T] as a 3-D tensor (T fastest); box = 32 positions x 64 filters of one utterance, swizzle = the
// row length of the box (128 bytes fp32, 64 bytes bf16)
// (C rows of the block starting at g; c_total rows per utterance)
bool make_grad_map(CUtensorMap* m, const void* g, bool bf16, int64_t B, int C, int c_total, int64_t T) {
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc) return false;
  const cuuint64_t esz = bf16 ? 2 : 4;
  const cuuint64_t dims[3] = {(cuuint64_t)T, (cuuint64_t)C, (cuuint64_t)B};
  const cuuint64_t strides[2] = {(cuuint64_t)T * esz, (cuuint64_t)T * esz * (cuuint64_t)c_total};
  const cuuint32_t box[3] = {32, 64, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  return enc(m, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(g), dims,
             strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             bf16 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
}  // namespace

cudaError_t launch_bwd_input_tc16(const Bank& bank, const void* gy, int gy_dtype, float* gx, int64_t B, int64_t T,
                                  const DeviceInfo& di, cudaStream_t st, const uint8_t* filter_table, int reserved_sms,
                                  long long* phase) {
  int sms = di.sm_count - reserved_sms;  // data-parallel: a few SMs stay free for the all-reduce in flight
  if (sms < 1) sms = 1;
  const size_t smem = H2_SM_TOTAL + 1024;
  auto kern = gy_dtype == SINC_DTYPE_BF16 ? bwd_input_tc16_kernel<true> : bwd_input_tc16_kernel<false>;
  cudaError_t e = ensure_dynamic_smem(reinterpret_cast<const void*>(kern), (int)smem);
  if (e != cudaSuccess) return e;
  const size_t esz = gy_dtype == SINC_DTYPE_BF16 ? 2 : 4;
  // the GEMM's K dimension holds 64 filters (x 2 tap parities): wider banks take one launch per block of 64 filters,
  // each adding its share to gx (every output sample is written by exactly one CTA per launch: no race)
  for (int cb0 = 0; cb0 < bank.C; cb0 += 64) {
    BwdInputTc16Args a;
    a.bank = bank;
    a.bank.low_hz = bank.low_hz + cb0;
    a.bank.band_hz = bank.band_hz + cb0;
    a.bank.C = bank.C - cb0 < 64 ? bank.C - cb0 : 64;
    a.g = reinterpret_cast<const uint8_t*>(gy) + (size_t)cb0 * (size_t)T * esz;
    a.c_total = bank.C;
    a.accumulate = cb0 > 0 ? 1 : 0;
    a.gx = gx;
    a.B = B;
    a.T = T;
    a.blocks_per_row = ceil_div64(T, H2_NS);
    a.nblocks = B * a.blocks_per_row;
    a.phase = phase ? phase + 6 : nullptr;  // sinc_options.phase_stamps (diagnostic, normally null)
    a.tab = bank.C <= 64 ? filter_table : nullptr;   // the table holds one block of 64 filters
#ifdef SINC_ABLATION
    a.ablate = getenv("SINC_ABLATE_BI") ? atoi(getenv("SINC_ABLATE_BI")) : 0;
#endif
    const int grid = (int)(a.nblocks < sms ? a.nblocks : sms);
    CUtensorMap gmap;
    memset(&gmap, 0, sizeof(gmap));
    if (!make_grad_map(&gmap, a.g, gy_dtype == SINC_DTYPE_BF16, B, a.bank.C, bank.C, T)) return cudaErrorNotSupported;
    e = launch_pdl(kern, grid, H2_THREADS, smem, st, a, gmap);
    if (e != cudaSuccess) return e;
    count_launch();
  }
  return cudaGetLastError();
}

}  // namespace sinc
