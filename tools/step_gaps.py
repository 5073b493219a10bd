"""Where the step's time goes between the kernels: events around every launch inside back-to-back cfg-2 steps, against
each kernel looped alone.  usage: step_gaps.py [steps]"""
import ctypes, os, sys, statistics
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import sincformer_metacog_speech_enhancement_b200 as sb
from sincformer_metacog_speech_enhancement_b200 import _lib
from sincformer_metacog_speech_enhancement_b200.sinc import _bank

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 40
B, T, C, K = 32, 48000, 64, 251
dev = torch.device("cuda:0"); lib = _lib.load()
m = sb.SincConv1d(C, K, sample_rate=16000).to(dev)
fb = _bank(m.low_hz_.detach(), m.band_hz_.detach(), m.window, m.n_, m.sample_rate, m.min_low_hz, m.min_band_hz)
P = lambda t: ctypes.c_void_p(t.data_ptr())
x = torch.randn(B, 1, T, device=dev); y = torch.empty(B, C, T, device=dev); g = torch.randn(B, C, T, device=dev)
gx = torch.empty(B, 1, T, device=dev); gp = torch.empty(2, C, device=dev)
nbytes = lib.sinc_workspace_bytes(B, T, C, K); ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
fwd = lambda: lib.sinc_forward(ctypes.byref(fb), P(x), T, P(y), 0, B, T, P(ws), nbytes, st, None)
gf = lambda: lib.sinc_backward(ctypes.byref(fb), P(x), T, P(g), 0, None, P(gp[0]), P(gp[1]), None, B, T, P(ws), nbytes, st, None)
bi = lambda: lib.sinc_backward(ctypes.byref(fb), P(x), T, P(g), 0, P(gx), None, None, None, B, T, P(ws), nbytes, st, None)
both = lambda: lib.sinc_backward(ctypes.byref(fb), P(x), T, P(g), 0, P(gx), P(gp[0]), P(gp[1]), None, B, T, P(ws), nbytes, st, None)
ev = lambda: torch.cuda.Event(enable_timing=True)


def loop(fns, n):
    for _ in range(5):
        for f in fns: assert f() == 0
    torch.cuda.synchronize()
    e0, e1 = ev(), ev()
    e0.record()
    for _ in range(n):
        for f in fns: f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


print(f"step (fwd, backward in one call), no events: {loop([fwd, both], steps):.1f} us")
print(f"step (fwd, gf, bi as three calls), no events: {loop([fwd, gf, bi], steps):.1f} us")
alone = {n: loop([f], steps) for n, f in (("fwd", fwd), ("gf", gf), ("bi", bi))}
print("looped alone: " + "  ".join(f"{k} {v:.1f}" for k, v in alone.items()) + f"   sum {sum(alone.values()):.1f} us")
for pair in (("fwd", "gf"), ("gf", "bi"), ("bi", "fwd")):
    f = {"fwd": fwd, "gf": gf, "bi": bi}
    t = loop([f[pair[0]], f[pair[1]]], steps)
    print(f"pair {pair[0]} -> {pair[1]}: {t:.1f} us  (alone sum {alone[pair[0]] + alone[pair[1]]:.1f})")
evs = [[ev() for _ in range(4)] for _ in range(steps)]
for _ in range(5):
    fwd(); gf(); bi()
torch.cuda.synchronize()
for s in range(steps):
    evs[s][0].record(); fwd(); evs[s][1].record(); gf(); evs[s][2].record(); bi(); evs[s][3].record()
torch.cuda.synchronize()
seg = [[evs[s][i].elapsed_time(evs[s][i + 1]) * 1e3 for s in range(steps)] for i in range(3)]
whole = [evs[s][0].elapsed_time(evs[s + 1][0]) * 1e3 for s in range(steps - 1)]
print("inside the step (events between the launches): " + "  ".join(
    f"{n} {statistics.median(v):.1f}" for n, v in zip(("fwd", "gf", "bi"), seg)) + f"   step {statistics.median(whole):.1f} us")
print(f"step (fwd, backward in one call), no events, again: {loop([fwd, both], steps):.1f} us")
