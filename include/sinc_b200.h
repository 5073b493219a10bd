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
 *   - the library has no mutable process-wide state (per-call options instead: sinc_options) apart from immutable
 *     per-device caches (SM count, the driver entry point of cuTensorMapEncodeTiled) and a launch counter;
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

#define SINC_ABI_VERSION 2

#define SINC_DTYPE_F32 0
#define SINC_DTYPE_BF16 1

/* kernel-family selector (sinc_options.engine): 0 = pick the fastest supported for the shape (default), 1 = FP32 FMA
 * kernels only, 2 = tcgen05 tensor-core kernels only (16-bit split operands where a kernel for the shape exists, else
 * 3xTF32; error if neither supports the shape), 3 = tcgen05 with the 3xTF32 split in every pass.  Every engine is
 * hand-written sm_100a CUDA; there is no CPU path. */
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

/* Number of kernels this library has launched in this process (monotonic counter; for bench accounting). */
uint64_t sinc_kernel_launch_count(void);

/* Per-call options of sinc_forward / sinc_backward (NULL = all defaults).  The library keeps NO mutable state between
 * calls: everything that used to be a process-wide setting travels with the call, so two modules on two devices or
 * threads cannot disturb each other.  No counterpart in the reference (single process, one engine).
 *   engine        SINC_ENGINE_*
 *   reserved_sms  data-parallel use: the persistent grad-input kernel (one CTA per SM, all of its shared memory and
 *                 registers) leaves this many SMs free (0..64), so that a collective launched next to it -- the
 *                 parameter-gradient all-reduce in flight -- can start before it ends
 *   engines_used  optional HOST int: receives which engine each pass of THIS call used -- bit 0 forward, bit 1
 *                 grad-input, bit 2 grad-filter: set = tcgen05, clear = FMA; bits 3/4/5 (same order): the tcgen05
 *                 kernel used 16-bit split operands
 *   phase_stamps  optional DEVICE buffer of 20 int64 (diagnostic, written by CTA 0 of each 16-bit tcgen05 kernel):
 *                 out[6*k + ...] = {SM cycle counter, nanosecond timer} at kernel entry, after the prologue and at
 *                 exit, for k = 0 forward, 1 grad-input, 2 grad-filter (cycles / ns = the SM clock the kernel really
 *                 ran at); out[18] = 1 if the grad-input launch took its filters from the workspace table */
typedef struct sinc_options {
  int32_t engine;
  int32_t reserved_sms;
  int32_t* engines_used;
  long long* phase_stamps;
} sinc_options;

/* Bytes of device scratch sinc_forward / sinc_backward need for this shape (never 0).  For lengths that are not a
 * multiple of 8 samples this includes room for a zero-padded copy of grad_y with 16-byte aligned rows: the 16-bit
 * tensor-core backward kernels stage grad_y with bulk copies / TMA boxes, so sinc_backward re-pitches a ragged grad_y
 * once (one extra streaming pass, exact) instead of dropping to a slower kernel family.  A smaller workspace (at least
 * the size for an aligned length) is still accepted: the backward then uses the kernels that take any length. */
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
                 void* stream, const sinc_options* opts);

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
                  size_t workspace_bytes, void* stream, const sinc_options* opts);

/* ---- next row N1: the GroupNorm(8, C) + GELU that follows the layer -----------------
 * Replaces agents/perception.py:233-234 (the GroupNorm module `sinc_norm` = nn.GroupNorm(8, C) of :157 applied to
 * the layer output, then F.gelu) with streaming kernels that touch the (B,C,T) activation as few times as possible.  fp32; any T
 * (16-byte accesses when T % 4 == 0 and the tensors are 16-byte aligned, 4-byte accesses otherwise).
 *
 * sinc_forward_gn: sinc_forward (fp32 y) that also emits, from the convolution epilogue, the partial sums the
 *   GroupNorm statistics need: gn_partial is float2 [B][C/8][sinc_forward_gn_partials(T)] = (sum y, sum y^2) per
 *   utterance, group of 8 filters and output tile.  SINC_E_ENGINE if the 16-bit tensor-core forward cannot take the
 *   shape (C % 8 != 0, K > 265) -- then use sinc_forward + sinc_gn_stats.
 * sinc_gn_stats: one read pass over y -> float2 partial [B*C][sinc_gn_chunks(T)].
 * sinc_gn_gelu_forward: z = gelu(gamma_c * (y - mean) * rstd + beta_c); `partial` holds, per group, rows_per_group
 *   rows of np pairs (rows_per_group = C/G, np = sinc_gn_chunks(T) after sinc_gn_stats; rows_per_group = 1,
 *   np = sinc_forward_gn_partials(T) after sinc_forward_gn); stats (float2 [B][G] = mean, rstd) is an output.
 * sinc_gn_gelu_backward: dy (and dgamma, dbeta (C) unless both NULL) from dz, y and the saved stats; `partial` is
 *   float2 scratch [B*C][sinc_gn_chunks(T)]. */
int64_t sinc_forward_gn_partials(int64_t T);
int sinc_forward_gn(const sinc_filterbank* fb, const float* x, int64_t x_stride_b, float* y, int64_t B, int64_t T,
                    float* gn_partial, void* stream, const sinc_options* opts);
int sinc_gn_chunks(int64_t T);
int sinc_gn_stats(const float* y, int64_t B, int64_t T, int32_t C, float* partial, void* stream);
int sinc_gn_gelu_forward(const float* y, const float* partial, int32_t np, int32_t rows_per_group, const float* gamma,
                         const float* beta, float* z, float* stats, int64_t B, int64_t T, int32_t C, int32_t G, float eps,
                         void* stream);
int sinc_gn_gelu_backward(const float* dz, const float* y, const float* stats, const float* gamma, const float* beta,
                          float* dy, float* dgamma, float* dbeta, float* partial, int64_t B, int64_t T, int32_t C,
                          int32_t G, void* stream);

/* A protocol failure inside a tcgen05 kernel (a pipeline barrier that never completes -- a bug, never a data-dependent
 * condition) is LOUD: after 30 s of wall time on one barrier the kernel executes a trap, the launch fails with a sticky
 * CUDA error and every later call returns it (> 0).  There is no way to obtain rc 0 with invalid results.
 * Hardware probes and micro-benchmarks (operand-layout probe, MMA rate, mbarrier cost) are NOT part of this library:
 * they live in libsinc_b200_diag.so (include/sinc_b200_diag.h), which only tests and tools load. */

#ifdef __cplusplus
}
#endif
#endif /* SINC_B200_H_ */
