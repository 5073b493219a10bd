// Internal declarations shared by the translation units of libsinc_b200.so (not part of the ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "sinc_b200.h"

namespace sinc {

// Device-side view of the module state (agents/perception.py:35-77) -- passed by value to kernels.
struct Bank {
  const float* low_hz;
  const float* band_hz;
  const float* window;
  const float* n;
  int C;
  int K;
  double sr;
  double min_low;
  double min_band;
};

inline Bank to_bank(const sinc_filterbank* fb) {
  Bank b;
  b.low_hz = fb->low_hz;
  b.band_hz = fb->band_hz;
  b.window = fb->window;
  b.n = fb->n;
  b.C = fb->C;
  b.K = fb->K;
  b.sr = fb->sample_rate;
  b.min_low = fb->min_low_hz;
  b.min_band = fb->min_band_hz;
  return b;
}

struct DeviceInfo {
  int sm_count;
  int cc_major;
  int cc_minor;
  int max_smem_optin;
};
// Cached per device (immutable after first query).
int get_device_info(DeviceInfo* out);

void count_launch(int n = 1);
uint64_t launch_count();
// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per (kernel, device) instead of once per launch
cudaError_t ensure_dynamic_smem(const void* kernel, int bytes);

// Launch with programmatic stream serialisation allowed (see sinc_tc.cuh: griddep_wait / griddep_launch).  The kernel MUST
// execute griddepcontrol.wait before its first global-memory access.
template <class... KArgs, class... Args>
cudaError_t launch_pdl(void (*kern)(KArgs...), int grid, int threads, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3((unsigned)threads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// ---- tiling constants shared between workspace sizing and the kernels -------------------------
constexpr int kChanBlock = 64;    // channels handled per CTA (filters kept in shared memory)
constexpr int kTapBlock = 256;    // taps handled per grad-filter CTA
constexpr size_t kFilterTableBytes = 4096 + 64 * 256 * 4;  // header + W[64][256] (layout: sinc_synth.cuh)
constexpr int kMaxPartialSlots = 2 * 160;  // grad-filter partial tiles (2 CTAs/SM, <=160 SMs)

inline int round_up(int v, int m) { return (v + m - 1) / m * m; }
inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// Workspace layout (all offsets 256-B aligned):
//   [0, partial_bytes)  grad-filter partial tiles: slots x Cpad x Kpad fp32
struct WorkspaceLayout {
  int Cpad;         // C rounded up to kChanBlock
  int Kpad;         // K rounded up to kTapBlock
  size_t partial_bytes;
  size_t table_offset;  // filter table (sinc_synth.cuh), from the 256-byte aligned workspace base
  size_t sync_offset;   // 256 bytes after the table: the arrival word of the fused finalize (sinc_tc16_bwd_filter.cu)
  size_t total_bytes;
};
WorkspaceLayout workspace_layout(int C, int K);

// Ragged lengths (T % 4 != 0 for fp32 / T % 8 != 0 for bf16 grad_y, or a grad_y pointer that is not 16-byte aligned): the
// tensor-core backward kernels stage grad_y with bulk copies and TMA boxes, which need every row to start on a 16-byte
// boundary.  sinc_backward then re-pitches grad_y (and the waveform) once into the tail of the workspace with rows of
// Tp = T rounded up to 8 samples, zero filled -- exact, because the convolution pads with zeros anyway -- and runs the
// same kernels on the padded copy: one extra streaming pass over grad_y instead of a 5x slower kernel family.
struct RepitchLayout {
  int64_t Tp;            // padded row length
  size_t g_offset;       // (B,C,Tp) in grad_y's element type (sized for fp32)
  size_t x_offset;       // (B,Tp) fp32
  size_t gx_offset;      // (B,Tp) fp32
  size_t total_bytes;    // workspace size including the padded copies
};
RepitchLayout repitch_layout(int64_t B, int64_t T, int C, int K);
cudaError_t launch_pad_rows(const void* src, int64_t src_pitch, void* dst, int64_t dst_pitch, int64_t rows, int64_t n,
                            int esz, cudaStream_t st);   // dst[r][0..dst_pitch) = src[r][0..n), zeros beyond n

// ---- FP32-FMA engine (sinc_fma.cu) ---------------------------------------------------------------
bool fma_supported(int C, int K, const DeviceInfo& di);
cudaError_t launch_synthesize(const Bank& bank, float* filters_out, cudaStream_t st);
cudaError_t launch_fwd_fma(const Bank& bank, const float* x, int64_t x_stride_b, void* y, int y_dtype,
                           int64_t B, int64_t T, const DeviceInfo& di, cudaStream_t st);
cudaError_t launch_bwd_input_fma(const Bank& bank, const void* gy, int gy_dtype, float* gx, int64_t B,
                                 int64_t T, const DeviceInfo& di, cudaStream_t st);
// writes `*slots_used` partial tiles into ws_partials
cudaError_t launch_bwd_filter_fma(const Bank& bank, const float* x, int64_t x_stride_b, const void* gy,
                                  int gy_dtype, float* ws_partials, int* slots_used, int64_t B,
                                  int64_t T, const DeviceInfo& di, cudaStream_t st);
// reduces the partial tiles (fp64, fixed order) and applies the synthesis chain rule in fp64
cudaError_t launch_finalize(const Bank& bank, const float* ws_partials, int slots, int Cpad, int Kpad, int layout,
                            float* grad_low, float* grad_band, float* grad_filters, cudaStream_t st);

// ---- tcgen05 engine (sinc_tc_fwd.cu, ...) -------------------------------------------------------------
bool tc_fwd_supported(int C, int K, int64_t T, const void* y, int y_dtype, const DeviceInfo& di);
cudaError_t launch_fwd_tc(const Bank& bank, const float* x, int64_t x_stride_b, void* y, int64_t B, int64_t T,
                          const DeviceInfo& di, cudaStream_t st);
bool tc_bwd_filter_supported(int C, int K, int64_t T, const void* g, int g_dtype, const DeviceInfo& di);
size_t tc_bwd_filter_partial_bytes(int C);
cudaError_t launch_bwd_filter_tc(const Bank& bank, const float* x, int64_t x_stride_b, const void* gy, float* ws_partials,
                                 int* slots_used, int64_t B, int64_t T, const DeviceInfo& di, cudaStream_t st);

bool tc_bwd_input_supported(int C, int K, const void* g, int g_dtype, const DeviceInfo& di);
cudaError_t launch_bwd_input_tc(const Bank& bank, const void* gy, float* gx, int64_t B, int64_t T, const DeviceInfo& di,
                                cudaStream_t st);


// 16-bit split variants (bf16 hi/mid, three kind::f16 MMAs): sinc_tc16_*.cu
bool tc16_bwd_input_supported(int C, int K, int64_t T, const void* g, int g_dtype, const DeviceInfo& di);
cudaError_t launch_bwd_input_tc16(const Bank& bank, const void* gy, int gy_dtype, float* gx, int64_t B, int64_t T,
                                  const DeviceInfo& di, cudaStream_t st, const uint8_t* filter_table, int reserved_sms,
                                  long long* phase);
bool tc16_fwd_supported(int C, int K, int64_t T, const void* y, int y_dtype, const DeviceInfo& di);
cudaError_t launch_fwd_tc16(const Bank& bank, const float* x, int64_t x_stride_b, void* y, int y_dtype, int64_t B,
                            int64_t T, const DeviceInfo& di, cudaStream_t st, float* gn_partial, long long* phase);
int64_t tc16_fwd_tiles_per_row(int64_t T);
bool tc16_bwd_filter_supported(int C, int K, int64_t T, const void* g, int g_dtype, const DeviceInfo& di);
size_t tc16_bwd_filter_partial_bytes(int C);
// `fin` non-null and C <= 64: the kernel finalises itself (fixed-order fp64 reduction of the partial tiles + chain rule in
// its tail) and the caller must NOT launch finalize_kernel; *fused_out says which happened.
struct FusedFinalize {
  unsigned long long* sync_word;  // 8-byte aligned word in the workspace (any contents)
  float* grad_low;
  float* grad_band;
  float* grad_filters;            // may be null
};
cudaError_t launch_bwd_filter_tc16(const Bank& bank, const float* x, int64_t x_stride_b, const void* gy, int gy_dtype,
                                   float* ws_partials, int* slots_used, int64_t B, int64_t T, const DeviceInfo& di,
                                   cudaStream_t st, uint8_t* filter_table, long long* phase, int reserved_sms,
                                   const FusedFinalize* fin, bool* fused_out);

}  // namespace sinc
