/*
 * sinc_b200.h -- C ABI of the B200-native SincConv hot path (libsinc_b200.so).
 *
 * This is the drop-in boundary for ONE path of MarvelMathesh/sincformer-metacog-speech-enhancement:
 * the Perception Agent's SincNet front-end, agents/perception.py::SincConv1d, forward and backward.
 * The reference has no FFI of its own (it is pure PyTorch); each entry point below cites the
 * reference code it replaces (paths relative to the reference repository root).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer on the current CUDA device unless its name ends in _host;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream); all launches are
 *     asynchronous on it, the library never synchronises the device;
 *   - the library allocates nothing and frees nothing: the caller owns x, y, grads and the
 *     workspace (size from sinc_workspace_bytes), and may reuse the workspace across calls that
 *     are ordered on one stream.  The workspace needs no initialisation and carries no state the
 *     caller must preserve: besides scratch it holds a self-validating copy of the synthesised
 *     filters (written by a parameter-gradient call, used by a later grad-input call only if its
 *     header matches that call's filter bank word for word and its checksums hold);
 *   - return value: 0 = success, < 0 = argument error (SINC_E_*), > 0 = a cudaError_t;
 *     sinc_strerror() turns either into text.  Nothing throws or exits across this boundary;
 *   - C = out_channels, K = kernel_size (odd; the reference bumps even K to K+1 in __init__,
 *     agents/perception.py:44-45 -- the host mirror does the same before calling in),
 *     H = (K-1)/2, x is (B,1,T) with unit time stride and x_stride_b elements between rows,
 *     y / grad_y are (B,C,T) contiguous, filter-major per utterance (strides C*T, T, 1).
 *   - element types: SINC_DTYPE_F32 (0) or SINC_DTYPE_BF16 (1) for y and grad_y; everything else fp32.
 */
#ifndef SINC_B200_H_
#define SINC_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SINC_ABI_VERSION 1

#define SINC_DTYPE_F32 0
#define SINC_DTYPE_BF16 1

/* kernel-family selector for sinc_set_engine(): 0 = pick the fastest supported for the shape
 * (default), 1 = FP32 FMA kernels only, 2 = tcgen05 tensor-core kernels only (16-bit split operands where a
 * kernel for the shape exists, else 3xTF32; error if neither supports the shape), 3 = tcgen05 with the
 * 3xTF32 split in every pass.  Every engine is hand-written sm_100a CUDA; there is no CPU path. */
#define SINC_ENGINE_AUTO 0
#define SINC_ENGINE_FMA 1
#define SINC_ENGINE_TC 2
#define SINC_ENGINE_TC_TF32 3 /* tcgen05 kernels, 3xTF32 split in every pass (fp32-grade gradients, half the rate) */

#define SINC_OK 0
#define SINC_E_NULL (-1)        /* a required pointer is NULL                               */
#define SINC_E_SHAPE (-2)       /* B, T, C or K out of range (K must be odd, all >= 1)      */
#define SINC_E_DTYPE (-3)       /* unknown element type code                                */
#define SINC_E_WORKSPACE (-4)   /* workspace NULL or smaller than sinc_workspace_bytes()    */
#define SINC_E_UNSUPPORTED (-5) /* shape too large for on-chip filter storage (see DESIGN)  */
#define SINC_E_ALIGN (-6)       /* a pointer is not aligned to its element type             */
#define SINC_E_DEVICE (-7)      /* current device is not compute capability 10.x            */
#define SINC_E_ENGINE (-8)      /* requested engine cannot run this shape                   */

/* The four small state tensors of the module plus its scalar attributes.
 * Replaces: the attributes/parameters/buffers set up in SincConv1d.__init__
 * (agents/perception.py:35-77): low_hz_ (C,1), band_hz_ (C,1), window (K,), n_ (1,H). */
typedef struct sinc_filterbank {
  const float* low_hz;   /* (C)   parameter low_hz_                                   */
  const float* band_hz;  /* (C)   parameter band_hz_                                  */
  const float* window;   /* (K)   buffer window                                       */
  const float* n;        /* (H)   buffer n_  (may be NULL when K == 1)                */
  int32_t C;
  int32_t K;
  double sample_rate;    /* module attribute sample_rate                              */
  double min_low_hz;     /* module attribute min_low_hz                               */
  double min_band_hz;    /* module attribute min_band_hz                              */
} sinc_filterbank;

int sinc_abi_version(void);
const char* sinc_strerror(int code);

/* Number of kernels this library has launched in this process (monotonic; for bench accounting). */
uint64_t sinc_kernel_launch_count(void);

/* Select the kernel family for subsequent calls from this process (SINC_ENGINE_*). */
int sinc_set_engine(int engine);
int sinc_get_engine(void);
/* Data-parallel use: the grad-input kernel is persistent (one CTA per SM, all of its shared memory and registers), so a
 * collective launched next to it cannot start until it ends.  With n > 0 that kernel leaves n SMs free (process-wide,
 * 0..64, default 0); the host mirror sets it around the grad-input launch when the parameter-gradient all-reduce
 * is in flight.  No counterpart in the reference (single process). */
int sinc_set_reserved_sms(int n);
/* Which engine the last successful sinc_forward / sinc_backward call of this process actually used:
 * bit 0 forward, bit 1 grad-input, bit 2 grad-filter; a set bit means tcgen05, clear means FMA.
 * Bits 3/4/5 (forward / grad-input / grad-filter): the tcgen05 kernel used 16-bit split operands. */
int sinc_last_engine_mask(void);

/* Bytes of device scratch sinc_forward / sinc_backward need for this shape (never 0). */
size_t sinc_workspace_bytes(int64_t B, int64_t T, int32_t C, int32_t K);

/* Filter synthesis alone: parameters -> (C,K) fp32 filter bank, row-major.
 * Replaces agents/perception.py:89-115 (the part of SincConv1d.forward before F.conv1d).
 * Debug / test entry point: the convolution kernels synthesise in their own prologue and never
 * read this output. */
int sinc_synthesize(const sinc_filterbank* fb, float* filters_out, void* stream);

/* Forward: y[b,c,t] = sum_k w[c,k] * x[b,t+k-H], zero padded, stride 1.
 * Replaces SincConv1d.forward (agents/perception.py:79-118) = synthesis + F.conv1d(stride=1,
 * padding=K//2).  y_dtype selects fp32 or bf16 storage of y (math is fp32 either way). */
int sinc_forward(const sinc_filterbank* fb, const float* x, int64_t x_stride_b, void* y,
                 int32_t y_dtype, int64_t B, int64_t T, void* workspace, size_t workspace_bytes,
                 void* stream);

/* Backward of sinc_forward.
 * Replaces autograd's ConvolutionBackward0 of agents/perception.py:117-118 (grad-input, grad-filter)
 * followed by the backward of the synthesis graph agents/perception.py:89-115 (grad low_hz_/band_hz_).
 *   grad_y      (B,C,T) in gy_dtype
 *   grad_x      (B,T) fp32, row stride T, or NULL to skip grad-input (waveform.requires_grad False)
 *   grad_low, grad_band  (C) fp32, or both NULL to skip the parameter gradients
 *   grad_filters (C,K) fp32 or NULL: the intermediate dL/dw, exposed for parity tests.
 * The (C,K) reduction order is fixed by the shape (bitwise reproducible run to run); the synthesis
 * chain rule is evaluated in fp64 ("why fp64"). */
int sinc_backward(const sinc_filterbank* fb, const float* x, int64_t x_stride_b, const void* grad_y,
                  int32_t gy_dtype, float* grad_x, float* grad_low, float* grad_band,
                  float* grad_filters, int64_t B, int64_t T, void* workspace,
                  size_t workspace_bytes, void* stream);

/* ---- next row N1: the GroupNorm(8, C) + GELU that follows the layer -----------------
 * Replaces agents/perception.py:233-234 (the GroupNorm module `sinc_norm` = nn.GroupNorm(8, C) of :157 applied to
 * the layer output, then F.gelu) with streaming kernels that touch the (B,C,T) activation as few times as possible.  fp32, T % 4 == 0,
 * 16-byte aligned tensors.
 *
 * sinc_forward_gn: sinc_forward (fp32 y) that also emits, from the convolution epilogue, the partial sums the
 *   GroupNorm statistics need: gn_partial is float2 [B][C/8][sinc_forward_gn_partials(T)] = (sum y, sum y^2) per
 *   utterance, group of 8 filters and output tile.  SINC_E_ENGINE if the 16-bit tensor-core forward cannot take the
 *   shape (C % 8 != 0, K > 265, T % 4 != 0) -- then use sinc_forward + sinc_gn_stats.
 * sinc_gn_stats: one read pass over y -> float2 partial [B*C][sinc_gn_chunks(T)].
 * sinc_gn_gelu_forward: z = gelu(gamma_c * (y - mean) * rstd + beta_c); `partial` holds, per group, rows_per_group
 *   rows of np pairs (rows_per_group = C/G, np = sinc_gn_chunks(T) after sinc_gn_stats; rows_per_group = 1,
 *   np = sinc_forward_gn_partials(T) after sinc_forward_gn); stats (float2 [B][G] = mean, rstd) is an output.
 * sinc_gn_gelu_backward: dy (and dgamma, dbeta (C) unless both NULL) from dz, y and the saved stats; `partial` is
 *   float2 scratch [B*C][sinc_gn_chunks(T)]. */
int64_t sinc_forward_gn_partials(int64_t T);
int sinc_forward_gn(const sinc_filterbank* fb, const float* x, int64_t x_stride_b, float* y, int64_t B, int64_t T,
                    float* gn_partial, void* stream);
int sinc_gn_chunks(int64_t T);
int sinc_gn_stats(const float* y, int64_t B, int64_t T, int32_t C, float* partial, void* stream);
int sinc_gn_gelu_forward(const float* y, const float* partial, int32_t np, int32_t rows_per_group, const float* gamma,
                         const float* beta, float* z, float* stats, int64_t B, int64_t T, int32_t C, int32_t G, float eps,
                         void* stream);
int sinc_gn_gelu_backward(const float* dz, const float* y, const float* stats, const float* gamma, const float* beta,
                          float* dy, float* dgamma, float* dbeta, float* partial, int64_t B, int64_t T, int32_t C,
                          int32_t G, void* stream);

/* Diagnostic: *status_host = 1 if any tcgen05 kernel of this process gave up waiting on one of its internal
 * barriers (a protocol bug; results are then invalid).  Synchronises the device. */
int sinc_tc_status(int* status_host);

/* Diagnostic: time stamps of the LAST launch of each 16-bit tcgen05 kernel, taken by CTA 0: out20[6*k + ...] =
 * {SM cycle counter, nanosecond timer} at kernel entry, after the prologue (filter synthesis, operand set-up) and
 * at exit, for k = 0 forward, 1 grad-input, 2 grad-filter.  cycles / ns = the SM clock the kernel really ran at.
 * out20[18] = 1 if that grad-input launch took its filters from the table the grad-filter kernel leaves in the
 * workspace, 0 if it synthesised them.  Synchronises the device. */
int sinc_tc_phase_stamps(long long* out20);

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

#ifdef __cplusplus
}
#endif
#endif /* SINC_B200_H_ */
