/*
 * sinc_b200_diag.h -- hardware probes and micro-benchmarks of the tcgen05 engine (libsinc_b200_diag.so).
 *
 * NOT part of the drop-in boundary (include/sinc_b200.h) and never loaded by the product path: only tests/ and
 * tools/ open this library.  The reference has no counterpart for anything here.  The probes exist so that the
 * hardware behaviour the kernels rely on -- in-place Toeplitz descriptors over a waveform tile, a TMEM-resident
 * A operand, the 16x256b TMEM load fragment, what an MMA and an mbarrier phase check cost -- is pinned by tests that
 * run on the GPU, instead of being folklore in comments.
 * Conventions as in sinc_b200.h: device pointers, cudaStream_t as void*, 0 = success.
 */
#ifndef SINC_B200_DIAG_H_
#define SINC_B200_DIAG_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Diagnostic: cycles per wait on an ALREADY COMPLETED mbarrier phase, one thread, dependent chain of `count` waits.
 * mode 0: mbarrier.try_wait with a suspend-time hint, 1: try_wait, 2: test_wait, 3: the kernels' bounded wait.
 * cycles_out2[0] = total SM cycles, [1] = number of waits that succeeded.  (Measured: 210 / 162 / 156 / 184 -- why
 * the pipelines are built to need few phase checks per unit of work.) */
int sinc_mbar_wait_cost(int mode, int count, long long* cycles_out2, void* stream);

/* Diagnostic: run `ksteps` tcgen05.mma (kind::tf32, M = 128) on caller-supplied shared-memory images and
 * raw matrix / instruction descriptors and return the 128 x N accumulator.  Not part of the replacement
 * surface (the reference has no counterpart); it exists so the operand layouts the tensor-core engine
 * relies on -- in-place Toeplitz descriptors over the waveform tile, TMEM-resident filter operand -- are
 * pinned by a hardware test (tests/test_gpu_tc_probe.py).  *status_out: 0 ok, 1 the MMA never completed. */
int sinc_tc_probe(const void* imgA, int bytesA, const void* imgB, int bytesB, unsigned long long descA,
                  unsigned long long descB, unsigned int idesc, int ksteps, int a_step, int b_step,
                  int a_from_tmem, int a_cols, int N, float* D_out, int* status_out, void* stream);


/* Diagnostic: cycles per tcgen05.mma (kind::f16, M = 128, N as given), `count` of them issued by one thread; the SM
 * cycle count of CTA 0 goes to cycles_out[0].  `a_from_tmem` is a bit field: bit 0 = A operand in TMEM (else shared
 * memory); bit 1 = non-zero operand data; bits 2..3 = instruction pattern (0: generic loop rotating through `nacc`
 * accumulators; 1: the forward kernel's sequence -- 17 K-steps x 3 products, B start += 32 bytes per K-step, `nacc` =
 * 1 one accumulator / 2 alternating / 3 runs of three; 2: the same with 128-byte aligned B starts; 3: pattern 1 with
 * a tcgen05.commit every `nacc` K-steps); bits 4..5 = what the CTA's other warps do meanwhile (0 nothing, 1
 * tcgen05.st, 2 LDS.128, 3 tcgen05.ld); bits 8..15 = number of CTAs (0 = one).  No counterpart in the reference;
 * tools/mma_rate*.py use it to establish what the tensor pipe costs and what it does not mind. */
int sinc_tc_mma_rate(int N, int nacc, int count, int a_from_tmem, long long* cycles_out, int* status_out, void* stream);

/* Diagnostic: the register fragment of tcgen05.ld.sync.aligned.16x256b.x8 (the load shape of the grad-input epilogue).
 * One warp fills 32 TMEM lanes x 64 columns with the value 256 * lane + column (tcgen05.st 32x32b), reads the block
 * back as two 16-lane halves with the 16x256b shape and stores thread t's 64 registers to out[t * 64 + i]
 * (i < 32: lanes 0..15, i >= 32: lanes 16..31).  Expected (PTX ISA, matrix fragment of .16x256b): register
 * 4*j + 2*u + e of a half holds lane (t / 4) + 8 * u, column 8 * j + 2 * (t % 4) + e. */
int sinc_tc_ld16x256_probe(uint32_t* out2048, void* stream);

/* Diagnostic: the rate at which 148 persistent CTAs can stream a (B,64,T) fp32 activation from HBM into shared memory
 * with no consumer at all -- the ceiling of the operand feed of the backward kernels.  mode 0/1/2: 1-D bulk copies of
 * 512 B / 1 KB / 2 KB row pieces (64 rows per stage), `depth` stages in flight per SM; mode 3: plain LDG.128.
 * T % 512 == 0, C == 64.  Time it with events around the call (tools/feed_probe.py). */
int sinc_feed_probe(const float* g, int64_t B, int32_t C, int64_t T, int mode, int depth, float* sink, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SINC_B200_DIAG_H_ */
