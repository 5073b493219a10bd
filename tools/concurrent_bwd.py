"""Experiment: grad-filter and grad-input CONCURRENTLY, each on half of the SMs, walking grad_y in the same order, so
that the second reader finds the activation in L2 (DRAM reads of the backward: 798 MB -> ~400 MB).  Through the C ABI,
two streams.  usage: concurrent_bwd.py [sms_for_grad_input ...]"""
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
x = torch.randn(B, 1, T, device=dev); g = torch.randn(B, C, T, device=dev); gx = torch.empty(B, 1, T, device=dev)
gp = torch.empty(2, C, device=dev); gx_ref = torch.empty_like(gx); gp_ref = torch.empty_like(gp)
nbytes = lib.sinc_workspace_bytes(B, T, C, K)
ws = torch.empty(nbytes, dtype=torch.uint8, device=dev); ws2 = torch.empty(nbytes, dtype=torch.uint8, device=dev)
s_main = torch.cuda.current_stream(); s_aux = torch.cuda.Stream()
def opt(reserved):
    o = _lib.Options(); o.reserved_sms = reserved; return o
def sequential(gxo=gx, gpo=gp):
    st = ctypes.c_void_p(s_main.cuda_stream)
    assert lib.sinc_backward(ctypes.byref(fb), P(x), T, P(g), 0, P(gxo), P(gpo[0]), P(gpo[1]), None, B, T, P(ws), nbytes, st, None) == 0
def concurrent(n_bi):
    n_gf = 148 - n_bi
    ev = torch.cuda.Event(); ev.record(s_main); s_aux.wait_event(ev)
    o1, o2 = opt(148 - n_gf), opt(148 - n_bi)
    assert lib.sinc_backward(ctypes.byref(fb), P(x), T, P(g), 0, None, P(gp[0]), P(gp[1]), None, B, T, P(ws), nbytes,
                             ctypes.c_void_p(s_main.cuda_stream), ctypes.byref(o1)) == 0
    assert lib.sinc_backward(ctypes.byref(fb), None, T, P(g), 0, P(gx), None, None, None, B, T, P(ws2), nbytes,
                             ctypes.c_void_p(s_aux.cuda_stream), ctypes.byref(o2)) == 0
    ev2 = torch.cuda.Event(); ev2.record(s_aux); s_main.wait_event(ev2)
def timeit(fn, n=30):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
sequential(gx_ref, gp_ref); torch.cuda.synchronize()
print(json.dumps({"sequential_ms": timeit(sequential)}), flush=True)
for n_bi in [int(a) for a in sys.argv[1:]] or [74, 70, 78, 66, 82]:
    ms = timeit(lambda: concurrent(n_bi))
    torch.cuda.synchronize()
    ok = bool(torch.equal(gx, gx_ref)) and float((gp - gp_ref).abs().max() / gp_ref.abs().max()) < 1e-4
    print(json.dumps({"grad_input_sms": n_bi, "grad_filter_sms": 148 - n_bi, "concurrent_ms": ms, "results_match": ok}), flush=True)
