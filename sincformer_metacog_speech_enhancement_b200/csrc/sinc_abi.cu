// C ABI of libsinc_b200.so: argument validation, engine selection, launch accounting.
// Declarations and the reference code each entry point replaces: include/sinc_b200.h.
#include "sinc_internal.h"

namespace sinc {

// AUTO picks the tcgen05 kernel for a pass once that kernel is parity-green on hardware.
constexpr bool kAutoPrefersTcForward = true;  // parity-green on B200 (tests/test_gpu_tc.py, r01c)
constexpr bool kAutoPrefersTcBwdFilter = true;  // parity-green on B200 (r01h)
constexpr bool kAutoPrefersTcBwdInput = true;   // parity-green on B200 (r01j)

// Per-call options with the defaults filled in (include/sinc_b200.h: sinc_options).  Nothing here outlives the call.
struct Options {
  int engine = SINC_ENGINE_AUTO;
  int reserved_sms = 0;
  int32_t* engines_used = nullptr;
  long long* phase = nullptr;
};
static int read_options(const sinc_options* o, Options* out) {
  if (!o) return 0;
  if (o->engine < SINC_ENGINE_AUTO || o->engine > SINC_ENGINE_TC_TF32) return SINC_E_ENGINE;
  if (o->reserved_sms < 0 || o->reserved_sms > 128) return SINC_E_SHAPE;
  if (o->phase_stamps && reinterpret_cast<uintptr_t>(o->phase_stamps) % 8) return SINC_E_ALIGN;
  out->engine = o->engine;
  out->reserved_sms = o->reserved_sms;
  out->engines_used = o->engines_used;
  out->phase = o->phase_stamps;
  return 0;
}

WorkspaceLayout workspace_layout(int C, int K) {
  WorkspaceLayout wl;
  wl.Cpad = round_up(C, kChanBlock);
  wl.Kpad = round_up(K, kTapBlock);
  wl.partial_bytes = (size_t)kMaxPartialSlots * wl.Cpad * wl.Kpad * sizeof(float);
  const size_t tc_bytes = tc_bwd_filter_partial_bytes(C);
  if (tc_bytes > wl.partial_bytes) wl.partial_bytes = tc_bytes;
  const size_t tc16_bytes = tc16_bwd_filter_partial_bytes(C);
  if (tc16_bytes > wl.partial_bytes) wl.partial_bytes = tc16_bytes;
  wl.table_offset = (wl.partial_bytes + 255) / 256 * 256;
  wl.sync_offset = wl.table_offset + kFilterTableBytes;
  wl.total_bytes = wl.sync_offset + 256 + 256;
  return wl;
}

RepitchLayout repitch_layout(int64_t B, int64_t T, int C, int K) {
  const WorkspaceLayout wl = workspace_layout(C, K);
  RepitchLayout rl;
  rl.Tp = (T + 7) / 8 * 8;
  rl.g_offset = (wl.total_bytes + 255) / 256 * 256;
  rl.x_offset = rl.g_offset + ((size_t)B * C * rl.Tp * 4 + 255) / 256 * 256;
  rl.gx_offset = rl.x_offset + ((size_t)B * rl.Tp * 4 + 255) / 256 * 256;
  rl.total_bytes = rl.gx_offset + ((size_t)B * rl.Tp * 4 + 255) / 256 * 256 + 256;
  return rl;
}

static int check_bank(const sinc_filterbank* fb) {
  if (!fb || !fb->low_hz || !fb->band_hz || !fb->window) return SINC_E_NULL;
  if (fb->C < 1 || fb->K < 1 || (fb->K % 2) == 0) return SINC_E_SHAPE;
  if (fb->K > 1 && !fb->n) return SINC_E_NULL;
  return 0;
}

static int check_device(DeviceInfo* di) {
  int r = get_device_info(di);
  if (r != 0) return r;
  if (di->cc_major != 10) return SINC_E_DEVICE;
  return 0;
}

}  // namespace sinc

using namespace sinc;

extern "C" {

int sinc_abi_version(void) { return SINC_ABI_VERSION; }

const char* sinc_strerror(int code) {
  switch (code) {
    case SINC_OK: return "success";
    case SINC_E_NULL: return "sinc: a required pointer is NULL";
    case SINC_E_SHAPE: return "sinc: bad shape (need B,T,C >= 1 and odd K >= 1)";
    case SINC_E_DTYPE: return "sinc: unknown element type code";
    case SINC_E_WORKSPACE: return "sinc: workspace missing or smaller than sinc_workspace_bytes()";
    case SINC_E_UNSUPPORTED: return "sinc: kernel_size too large for on-chip filter storage";
    case SINC_E_ALIGN: return "sinc: pointer not aligned to its element type";
    case SINC_E_DEVICE: return "sinc: current device is not an sm_100-class (compute capability 10.x) GPU";
    case SINC_E_ENGINE: return "sinc: the requested engine cannot run this shape";
    default: break;
  }
  if (code > 0) return cudaGetErrorString((cudaError_t)code);
  return "sinc: unknown error";
}

uint64_t sinc_kernel_launch_count(void) { return launch_count(); }

size_t sinc_workspace_bytes(int64_t B, int64_t T, int32_t C, int32_t K) {
  if (C < 1) C = 1;
  if (K < 1) K = 1;
  if (B < 1) B = 1;
  if (T < 1) T = 1;
  // lengths that are not a multiple of 8 samples: room for the re-pitched copies the backward makes (sinc_internal.h)
  if (T % 8 != 0) return repitch_layout(B, T, C, K).total_bytes;
  return workspace_layout(C, K).total_bytes;
}

int sinc_synthesize(const sinc_filterbank* fb, float* filters_out, void* stream) {
  int r = check_bank(fb);
  if (r) return r;
  if (!filters_out) return SINC_E_NULL;
  DeviceInfo di;
  if ((r = check_device(&di))) return r;
  return (int)launch_synthesize(to_bank(fb), filters_out, (cudaStream_t)stream);
}

int sinc_forward(const sinc_filterbank* fb, const float* x, int64_t x_stride_b, void* y, int32_t y_dtype,
                 int64_t B, int64_t T, void* workspace, size_t workspace_bytes, void* stream,
                 const sinc_options* opts) {
  int r = check_bank(fb);
  if (r) return r;
  Options o;
  if ((r = read_options(opts, &o))) return r;
  if (!x || !y) return SINC_E_NULL;
  if (B < 1 || T < 1) return SINC_E_SHAPE;
  if (y_dtype != SINC_DTYPE_F32 && y_dtype != SINC_DTYPE_BF16) return SINC_E_DTYPE;
  if (reinterpret_cast<uintptr_t>(x) % 4 || reinterpret_cast<uintptr_t>(y) % (y_dtype == SINC_DTYPE_F32 ? 4 : 2))
    return SINC_E_ALIGN;
  (void)workspace;
  (void)workspace_bytes;
  DeviceInfo di;
  if ((r = check_device(&di))) return r;
  const Bank bank = to_bank(fb);
  const int engine = o.engine;
  const bool tc_ok = tc_fwd_supported(bank.C, bank.K, T, y, y_dtype, di);
  const bool tc16_ok = tc16_fwd_supported(bank.C, bank.K, T, y, y_dtype, di);
  if (engine == SINC_ENGINE_TC && !tc_ok && !tc16_ok) return SINC_E_ENGINE;
  if (engine == SINC_ENGINE_TC_TF32 && !tc_ok) return SINC_E_ENGINE;
  const bool want_tc = engine == SINC_ENGINE_TC || engine == SINC_ENGINE_TC_TF32 ||
                       (engine == SINC_ENGINE_AUTO && kAutoPrefersTcForward);
  const bool use_tc16 = want_tc && tc16_ok && engine != SINC_ENGINE_TC_TF32;
  const bool use_tc = want_tc && tc_ok && !use_tc16;
  cudaError_t e;
  if (use_tc16) {
    e = launch_fwd_tc16(bank, x, x_stride_b, y, y_dtype, B, T, di, (cudaStream_t)stream, nullptr, o.phase);
  } else if (use_tc) {
    e = launch_fwd_tc(bank, x, x_stride_b, y, B, T, di, (cudaStream_t)stream);
  } else {
    if (!fma_supported(bank.C, bank.K, di)) return SINC_E_UNSUPPORTED;
    e = launch_fwd_fma(bank, x, x_stride_b, y, y_dtype, B, T, di, (cudaStream_t)stream);
  }
  if (e != cudaSuccess) return (int)e;
  if (o.engines_used) *o.engines_used = ((use_tc || use_tc16) ? 1 : 0) | (use_tc16 ? 8 : 0);
  return 0;
}

int64_t sinc_forward_gn_partials(int64_t T) { return tc16_fwd_tiles_per_row(T); }

int sinc_forward_gn(const sinc_filterbank* fb, const float* x, int64_t x_stride_b, float* y, int64_t B, int64_t T,
                    float* gn_partial, void* stream, const sinc_options* opts) {
  int r = check_bank(fb);
  if (r) return r;
  Options o;
  if ((r = read_options(opts, &o))) return r;
  if (!x || !y || !gn_partial) return SINC_E_NULL;
  if (B < 1 || T < 1) return SINC_E_SHAPE;
  if (reinterpret_cast<uintptr_t>(x) % 4 || reinterpret_cast<uintptr_t>(y) % 4 || reinterpret_cast<uintptr_t>(gn_partial) % 8)
    return SINC_E_ALIGN;
  DeviceInfo di;
  if ((r = check_device(&di))) return r;
  const Bank bank = to_bank(fb);
  const int engine = o.engine;
  if (engine == SINC_ENGINE_FMA || engine == SINC_ENGINE_TC_TF32) return SINC_E_ENGINE;
  if (bank.C % 8 != 0 || !tc16_fwd_supported(bank.C, bank.K, T, y, SINC_DTYPE_F32, di)) return SINC_E_ENGINE;
  cudaError_t e = launch_fwd_tc16(bank, x, x_stride_b, y, SINC_DTYPE_F32, B, T, di, (cudaStream_t)stream, gn_partial, o.phase);
  if (e != cudaSuccess) return (int)e;
  if (o.engines_used) *o.engines_used = 9;
  return 0;
}

int sinc_backward(const sinc_filterbank* fb, const float* x, int64_t x_stride_b, const void* grad_y,
                  int32_t gy_dtype, float* grad_x, float* grad_low, float* grad_band, float* grad_filters,
                  int64_t B, int64_t T, void* workspace, size_t workspace_bytes, void* stream,
                  const sinc_options* opts) {
  int r = check_bank(fb);
  if (r) return r;
  Options o;
  if ((r = read_options(opts, &o))) return r;
  if (!grad_y) return SINC_E_NULL;
  if ((grad_low == nullptr) != (grad_band == nullptr)) return SINC_E_NULL;
  const bool want_params = grad_low != nullptr || grad_filters != nullptr;
  if (want_params && !x) return SINC_E_NULL;
  if (B < 1 || T < 1) return SINC_E_SHAPE;
  if (gy_dtype != SINC_DTYPE_F32 && gy_dtype != SINC_DTYPE_BF16) return SINC_E_DTYPE;
  if (reinterpret_cast<uintptr_t>(grad_y) % (gy_dtype == SINC_DTYPE_F32 ? 4 : 2)) return SINC_E_ALIGN;
  DeviceInfo di;
  if ((r = check_device(&di))) return r;
  const Bank bank = to_bank(fb);
  if (!fma_supported(bank.C, bank.K, di)) return SINC_E_UNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  // Ragged rows: re-pitch grad_y (and x) into the workspace tail and run the tensor-core kernels on the padded copy.
  // Only when the engine may use them, the 16-bit kernels take the shape otherwise, and the caller sized the workspace
  // with sinc_workspace_bytes(B, T, ..) for this T.
  float* grad_x_final = nullptr;
  int64_t T_final = T;
  {
    const int esz = gy_dtype == SINC_DTYPE_F32 ? 4 : 2;
    const bool rows_aligned = T % (16 / esz) == 0 && reinterpret_cast<uintptr_t>(grad_y) % 16 == 0;
    const bool tc_wanted = o.engine == SINC_ENGINE_AUTO || o.engine == SINC_ENGINE_TC;
    if (!rows_aligned && tc_wanted && bank.K <= 256 && T % 8 != 0) {
      const RepitchLayout rl = repitch_layout(B, T, bank.C, bank.K);
      if (workspace && workspace_bytes >= rl.total_bytes) {
        uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(workspace) + 255) / 256 * 256);
        void* g4 = base + rl.g_offset;
        float* x4 = reinterpret_cast<float*>(base + rl.x_offset);
        float* gx4 = reinterpret_cast<float*>(base + rl.gx_offset);
        cudaError_t e2 = launch_pad_rows(grad_y, T, g4, rl.Tp, B * bank.C, T, esz, st);
        if (e2 != cudaSuccess) return (int)e2;
        if (want_params) {
          e2 = launch_pad_rows(x, x_stride_b, x4, rl.Tp, B, T, 4, st);
          if (e2 != cudaSuccess) return (int)e2;
          x = x4;
          x_stride_b = rl.Tp;
        }
        grad_y = g4;
        if (grad_x) {
          grad_x_final = grad_x;
          grad_x = gx4;
        }
        T = rl.Tp;
      }
    }
  }
  cudaError_t e;
  int mask = 0;
  const int engine = o.engine;
  // The filter table (sinc_synth.cuh): written by the 16-bit grad-filter kernel, read -- after validation -- by the
  // 16-bit grad-input kernel.  Parameter gradients therefore go first when both are wanted; a grad-input-only call
  // finds the table of the preceding parameter-gradient call on the same workspace (data parallel: the all-reduce
  // of the parameter gradients is in flight in between) or synthesises the filters itself.
  uint8_t* filter_table = nullptr;
  {
    const WorkspaceLayout wl = workspace_layout(bank.C, bank.K);
    if (workspace && workspace_bytes >= wl.total_bytes)
      filter_table = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(workspace) + 255) / 256 * 256) + wl.table_offset;
  }
  if (want_params) {
    const WorkspaceLayout wl = workspace_layout(bank.C, bank.K);
    if (!workspace || workspace_bytes < wl.total_bytes) return SINC_E_WORKSPACE;
    float* partials = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(workspace) + 255) / 256 * 256);
    int slots = 0;
    const bool tc_ok = tc_bwd_filter_supported(bank.C, bank.K, T, grad_y, gy_dtype, di);
    const bool tc16_ok = tc16_bwd_filter_supported(bank.C, bank.K, T, grad_y, gy_dtype, di);
    const bool want_tc = engine == SINC_ENGINE_TC || engine == SINC_ENGINE_TC_TF32 ||
                         (engine == SINC_ENGINE_AUTO && kAutoPrefersTcBwdFilter);
    const bool use_tc16 = want_tc && tc16_ok && engine != SINC_ENGINE_TC_TF32;
    const bool use_tc = want_tc && tc_ok && !use_tc16;
    bool fused = false;
    if (use_tc16) {
      // C <= 64: the kernel reduces its partial tiles and applies the chain rule in its own tail (one launch fewer)
      FusedFinalize fin;
#ifdef SINC_ABLATION
      const bool split_finalize = getenv("SINC_SPLIT_FINALIZE") != nullptr;   // A/B against the separate finalize_kernel
#else
      const bool split_finalize = false;
#endif
      fin.sync_word = reinterpret_cast<unsigned long long*>(reinterpret_cast<uint8_t*>(partials) + wl.sync_offset);
      fin.grad_low = grad_low;
      fin.grad_band = grad_band;
      fin.grad_filters = grad_filters;
      e = launch_bwd_filter_tc16(bank, x, x_stride_b, grad_y, gy_dtype, partials, &slots, B, T, di, st, filter_table, o.phase,
                                 o.reserved_sms, (grad_low && grad_band && !split_finalize) ? &fin : nullptr, &fused);
    } else if (use_tc)
      e = launch_bwd_filter_tc(bank, x, x_stride_b, grad_y, partials, &slots, B, T, di, st);
    else
      e = launch_bwd_filter_fma(bank, x, x_stride_b, grad_y, gy_dtype, partials, &slots, B, T, di, st);
    if (e != cudaSuccess) return (int)e;
    if (!fused) {
      e = launch_finalize(bank, partials, slots, wl.Cpad, wl.Kpad, use_tc16 ? 2 : use_tc ? 1 : 0, grad_low, grad_band,
                          grad_filters, st);
      if (e != cudaSuccess) return (int)e;
    }
    mask |= (use_tc || use_tc16) ? 4 : 0;
    mask |= use_tc16 ? 32 : 0;
  }
  if (grad_x) {
    const bool tc_ok = tc_bwd_input_supported(bank.C, bank.K, grad_y, gy_dtype, di);
    const bool tc16_ok = tc16_bwd_input_supported(bank.C, bank.K, T, grad_y, gy_dtype, di);
    const bool want_tc = engine == SINC_ENGINE_TC || engine == SINC_ENGINE_TC_TF32 ||
                         (engine == SINC_ENGINE_AUTO && kAutoPrefersTcBwdInput);
    const bool use_tc16 = want_tc && tc16_ok && engine != SINC_ENGINE_TC_TF32;
    const bool use_tc = want_tc && tc_ok && !use_tc16;
    if (use_tc16)
      e = launch_bwd_input_tc16(bank, grad_y, gy_dtype, grad_x, B, T, di, st, filter_table, o.reserved_sms, o.phase);
    else if (use_tc)
      e = launch_bwd_input_tc(bank, grad_y, grad_x, B, T, di, st);
    else
      e = launch_bwd_input_fma(bank, grad_y, gy_dtype, grad_x, B, T, di, st);
    if (e != cudaSuccess) return (int)e;
    mask |= (use_tc || use_tc16) ? 2 : 0;
    mask |= use_tc16 ? 16 : 0;
  }
  if (grad_x_final) {  // the padded grad-input rows back to the caller's (B,T) tensor
    e = launch_pad_rows(grad_x, T, grad_x_final, T_final, B, T_final, 4, st);
    if (e != cudaSuccess) return (int)e;
  }
  if (o.engines_used) *o.engines_used = mask;
  return 0;
}

}  // extern "C"
