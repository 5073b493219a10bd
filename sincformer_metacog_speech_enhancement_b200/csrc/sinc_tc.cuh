// tcgen05 / TMEM / mbarrier / bulk-copy primitives for sm_100a, as inline PTX.
// Descriptor bit layouts follow the PTX ISA "tcgen05 matrix descriptor" / "instruction descriptor" tables
// (the same fields CUTLASS exposes as cute::UMMA::SmemDescriptor / InstrDescriptor).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace sinc {
namespace tc {

// ---- shared-memory matrix descriptor (64 bit) ------------------------------------------------------
//  [0,14)  start address >> 4        [16,30) leading-dimension byte offset >> 4
//  [32,46) stride-dimension byte offset >> 4      [46,48) version = 1 (Blackwell)
//  [49,52) base offset (0)           [52] LBO mode (0)          [61,64) swizzle (0 = none / interleave)
// No-swizzle K-major operand (rows = M or N index, 16-byte chunks along K):
//     addr(row, kchunk) = start + (row % 8) * 16 + (row / 8) * SBO + kchunk * LBO
// No-swizzle MN-major operand (16-byte chunks along M/N, rows along K):
//     addr(mn_chunk, k) = start + mn_chunk * SBO + (k % 8) * 16 + (k / 8) * LBO
__host__ __device__ constexpr uint64_t make_smem_desc(uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) |
         ((uint64_t)1 << 46);
}
__device__ __forceinline__ uint64_t desc_at(uint64_t desc_base, uint32_t smem_addr_bytes) {
  return desc_base | (uint64_t)((smem_addr_bytes >> 4) & 0x3FFF);
}

// ---- instruction descriptor (32 bit) for kind::tf32, fp32 accumulate -------------------------------
//  [4,6) D format (1 = f32)   [7,10) A format (2 = tf32)   [10,13) B format (2 = tf32)
//  [15] A major (0 = K, 1 = MN)   [16] B major   [17,23) N >> 3   [24,29) M >> 4
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- warp-uniform MMA issue -----------------------------------------------------------------------------
// tcgen05.mma takes its descriptors from uniform registers.  Inside `if (lane == 0)` nothing is warp-uniform for
// the compiler, so every operand of every MMA goes through ELECT + R2UR (~100 cycles per instruction, ncu r01v).
// Issuer warps therefore run their loops with ALL lanes (uniform control flow, uniform values) and predicate only
// the tcgen05 instructions on the elected lane.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ int uniform(int v) { return __shfl_sync(0xffffffffu, v, 0); }

// ---- TMEM allocation (whole-warp, .sync.aligned) -------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(dst_smem)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// Programmatic dependent launch (the three persistent 16-bit kernels are launched with the stream-serialisation attribute,
// sinc_internal.h: launch_pdl): griddep_wait() returns when every kernel this one depends on has completed and its memory
// is visible -- nothing that reads or writes global memory may come before it; griddep_launch() lets the NEXT kernel on
// the stream be set up and its CTAs take the SMs as ours leave them (they sit in their own griddep_wait until this grid is
// complete), which takes the ~4 us of launch processing off the critical path between two of our kernels.
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void griddep_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// generic-proxy writes to shared memory -> visible to the async proxy (tensor core / bulk copy reads)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- MMA issue (one thread) -------------------------------------------------------------------------
// D[tmem] (+)= A[smem] * B[smem]^T
__device__ __forceinline__ void mma_tf32_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]^T
__device__ __forceinline__ void mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive on an mbarrier when every tcgen05.mma issued so far by this thread has completed.
// (implies tcgen05.fence::before_thread_sync)
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_addr(bar))
               : "memory");
}

// ---- kind::f16 (fp16 / bf16 operands, fp32 accumulate): K = 16 elements (32 bytes) per instruction -------
//  instruction descriptor: A/B format 0 = f16, 1 = bf16 (independent for A and B)
constexpr uint32_t kFmtF16 = 0, kFmtBF16 = 1;
__host__ __device__ constexpr uint32_t make_idesc_f16(int M, int N, uint32_t a_fmt, uint32_t b_fmt) {
  return (1u << 4) | (a_fmt << 7) | (b_fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_f16_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// A from TMEM: 16-bit elements packed two per 32-bit column (even k in the low half), 8 columns per K-step
__device__ __forceinline__ void mma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// kind::f16 instruction descriptor bits 13 / 14: negate the A / B operand on the way into the multiplier
constexpr uint32_t kIdescNegA = 1u << 13, kIdescNegB = 1u << 14;

// fp32 pair -> packed bf16 pairs (v1 in the high half): hi = RN(v), nmid = RN(hi - v) = -RN(v - hi), i.e. the NEGATED
// second piece.  Four instructions per pair -- F2FP, two mixed-precision FHADD.BF16 (sub.rn.f32.bf16 takes the packed
// halves directly), F2FP -- instead of six (the plain form needs a shift and a mask to widen hi before the two
// subtractions).  The MMA that consumes the piece sets the negate bit of that operand in its instruction descriptor,
// so the sign costs nothing.
__device__ __forceinline__ void split_bf16x2_negmid(float v0, float v1, uint32_t& hi, uint32_t& nmid) {
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(v1), "f"(v0));
  float n0, n1;
  asm("{\n\t.reg .b16 lo, hi;\n\tmov.b32 {lo, hi}, %2;\n\tsub.rn.f32.bf16 %0, lo, %3;\n\tsub.rn.f32.bf16 %1, hi, %4;\n\t}\n"
      : "=f"(n0), "=f"(n1)
      : "r"(hi), "f"(v0), "f"(v1));
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(nmid) : "f"(n1), "f"(n0));
}

// ---- TMEM <-> registers: 32 lanes x 32 bit, 8 consecutive columns per call ---------------------------
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
}
// 16 TMEM lanes x 64 columns as the 16x256b fragment: thread t of the warp receives, for j = 0..7, u = 0..1, e = 0..1,
//     v[4*j + 2*u + e] = TMEM[lane0 + (t / 4) + 8*u][col0 + 8*j + 2*(t % 4) + e]
// (the mma.sync-style accumulator fragment; pinned on hardware by the diagnostics library's sinc_tc_ld16x256_probe).
// A thread therefore holds two ROWS eight lanes apart for the same columns -- what lets the grad-input epilogue add
// rows of the accumulator that belong to the same output sample without any cross-thread traffic.
__device__ __forceinline__ void tmem_ld16x256_x8(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.16x256b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(v[0]),
               "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
// 32 consecutive columns per call: one MIO instruction instead of four (the MMA issuers share that queue)
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
      "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]),
      "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]),
      "r"(v[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---- mbarrier ------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
      : "=r"(ok)
      : "r"(smem_addr(bar)), "r"(parity), "r"(2000u)   // suspend-time hint (ns): fewer polls while the phase is pending
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must never end in a hung GPU -- and never in a silently wrong answer either.  After
// kWaitLimitNs of wall time (%globaltimer) on ONE barrier the thread executes a trap: the launch fails with a sticky
// CUDA error that the next API call of the host reports (include/sinc_b200.h).  The bound is long enough for
// compute-sanitizer, ncu replay, cuda-gdb single-stepping of another warp and time-sliced contexts; the polls back
// off with the suspend-time hint of mbarrier.try_wait, so a long wait costs no issue slots.
// The diagnostics library (-DSINC_DIAG_SOFT_TIMEOUT) keeps the non-trapping variant: its probes report a timeout
// through their own status word.  Callers keep the `ok = ok && mbar_wait(..)` form; in the product `ok` stays true.
constexpr unsigned long long kWaitLimitNs =
#if defined(SINC_DIAG_SOFT_TIMEOUT)
    1000000000ull;
#elif defined(SINC_ABLATION)
    3000000000ull;    // experiment builds: fail fast
#else
    30000000000ull;
#endif
__device__ __forceinline__ uint64_t global_timer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// One phase check without a suspend-time hint.  Any mbarrier phase check has ~160 cycles of latency even when the
// phase is already complete (tools/mbar_wait_cost.py: try_wait 162, test_wait 156, try_wait with a hint 210), so the
// MMA issuers call this EARLY -- for the barrier of their next step, before issuing the MMAs of the current one -- and
// read the result afterwards: the latency hides behind the issue work instead of sitting between two batches of MMAs.
__device__ __forceinline__ bool mbar_try_wait_plain(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
      : "=r"(ok)
      : "r"(smem_addr(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait_plain(bar, parity)) return true;  // the common case in a well-fed pipeline
  // fast path: no timer access while the barrier flips within a few thousand polls (the normal case)
#pragma unroll 1
  for (int spin = 0; spin < 4096; ++spin)
    if (mbar_try_wait(bar, parity)) return true;
  const uint64_t t0 = global_timer_ns();
  for (;;) {
#pragma unroll 1
    for (int spin = 0; spin < 256; ++spin)
      if (mbar_try_wait(bar, parity)) return true;
    if (global_timer_ns() - t0 > kWaitLimitNs) {
#ifdef SINC_DIAG_SOFT_TIMEOUT
      return false;
#else
      __trap();
#endif
    }
  }
}

// ---- 1-D bulk copies (TMA engine, no tensor map) -------------------------------------------------------
// shared -> global, bulk-group completion
__device__ __forceinline__ void bulk_store(void* gdst, const void* ssrc, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_addr(ssrc)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void bulk_wait() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}
// global -> shared, completes on an mbarrier (complete_tx::bytes)
__device__ __forceinline__ void bulk_load(void* sdst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_addr(sdst)),
               "l"(gsrc), "r"(bytes), "r"(smem_addr(bar))
               : "memory");
}
// global -> shared through a 3-D tensor map (box written densely, optionally swizzled; out-of-range elements
// read as zero), completes on an mbarrier.  `tmap` = address of a __grid_constant__ CUtensorMap kernel parameter.
__device__ __forceinline__ void tensor_load_3d(uint32_t sdst, const void* tmap, int x, int y, int z, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
          sdst),
      "l"(tmap), "r"(smem_addr(bar)), "r"(x), "r"(y), "r"(z)
      : "memory");
}
__device__ __forceinline__ void tensormap_prefetch(const void* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}
// L2 prefetch of a contiguous global range (no destination, no completion tracking): bytes % 16 == 0, 16-byte aligned
__device__ __forceinline__ void bulk_prefetch_l2(const void* gsrc, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gsrc), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// round-to-nearest fp32 -> tf32 (result is an fp32 bit pattern with the low 13 mantissa bits zero)
__device__ __forceinline__ float to_tf32(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return __uint_as_float(r);
}

}  // namespace tc
}  // namespace sinc
