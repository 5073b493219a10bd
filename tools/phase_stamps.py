"""Where the time of each tcgen05 kernel goes at cfg 2: prologue vs main loop, and the SM clock the kernel really ran
at (SM cycle counter / nanosecond timer, both read by CTA 0) -- steady-state back-to-back launches, like the bench."""
import ctypes, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import sincformer_metacog_speech_enhancement_b200 as sb
from sincformer_metacog_speech_enhancement_b200 import _lib
from sincformer_metacog_speech_enhancement_b200.sinc import _bank

dev = torch.device("cuda:0"); lib = _lib.load()
C, K, T, B = 64, 251, 48000, 32
m = sb.SincConv1d(C, K, sample_rate=16000).to(dev)
fb = _bank(m.low_hz_.detach(), m.band_hz_.detach(), m.window, m.n_, m.sample_rate, m.min_low_hz, m.min_band_hz)
P = lambda t: ctypes.c_void_p(t.data_ptr())
x = torch.randn(B, 1, T, device=dev); y = torch.empty(B, C, T, device=dev); g = torch.randn(B, C, T, device=dev)
gx = torch.empty(B, 1, T, device=dev); gp = torch.empty(2, C, device=dev)
nbytes = lib.sinc_workspace_bytes(B, T, C, K); ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
stamps2 = torch.zeros(2, 20, dtype=torch.int64, device=dev)
opts = []
for i in range(2):
    o = _lib.Options(); o.phase_stamps = stamps2[i].data_ptr()     # per-call diagnostic option of the C ABI
    opts.append(o)
def step(i):
    opt = opts[i & 1]
    assert lib.sinc_forward(ctypes.byref(fb), P(x), T, P(y), 0, B, T, P(ws), nbytes, st, ctypes.byref(opt)) == 0
    assert lib.sinc_backward(ctypes.byref(fb), P(x), T, P(g), 0, P(gx), P(gp[0]), P(gp[1]), None, B, T, P(ws), nbytes, st,
                             ctypes.byref(opt)) == 0
NSTEP = int(sys.argv[1]) if len(sys.argv) > 1 else 200
for i in range(NSTEP):     # ~90 ms of sustained load before the stamps that count
    step(i)
torch.cuda.synchronize()
prev = stamps2[NSTEP & 1].cpu().tolist()        # step NSTEP - 2
out = stamps2[(NSTEP - 1) & 1].cpu().tolist()   # step NSTEP - 1 (the last)
# CTA 0's view of the gaps between the kernels of back-to-back steps (its own end to its own start in the next kernel)
f0, f2, b0, b2, g0, g2 = out[1], out[5], out[7], out[11], out[13], out[17]
print(json.dumps({"cta0_gap_us": {"fwd_end->grad_filter_start": (g0 - f2) / 1e3, "grad_filter_main_end->grad_input_start": (b0 - g2) / 1e3, "grad_filter_tail_us": (out[19] - g2) / 1e3,
                                  "grad_input_end(previous step)->fwd_start": (f0 - prev[11]) / 1e3},
                  "step_us_cta0_fwd_start_to_fwd_start": (f0 - prev[1]) / 1e3}))
print(json.dumps({"grad_input_filters_from_table": int(out[18])}))
for k, name in enumerate(("fwd", "bwd_input", "bwd_filter")):
    c0, t0, c1, t1, c2, t2 = out[6 * k: 6 * k + 6]
    print(json.dumps({"kernel": name, "prologue_us": (t1 - t0) / 1e3, "main_us": (t2 - t1) / 1e3,
                      "prologue_cycles": c1 - c0, "main_cycles": c2 - c1,
                      "sm_mhz": round((c2 - c0) / ((t2 - t0) / 1e3), 1)}), flush=True)
