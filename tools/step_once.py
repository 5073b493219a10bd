"""A few cfg-2 steps (forward + backward) through the C ABI -- the command ncu wraps.  usage: step_once.py [steps] [B] [T]"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import sincformer_metacog_speech_enhancement_b200 as sb
from sincformer_metacog_speech_enhancement_b200 import _lib
from sincformer_metacog_speech_enhancement_b200.sinc import _bank

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 4
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
T = int(sys.argv[3]) if len(sys.argv) > 3 else 48000
dev = torch.device("cuda:0"); lib = _lib.load()
C, K = 64, 251
m = sb.SincConv1d(C, K, sample_rate=16000).to(dev)
fb = _bank(m.low_hz_.detach(), m.band_hz_.detach(), m.window, m.n_, m.sample_rate, m.min_low_hz, m.min_band_hz)
P = lambda t: ctypes.c_void_p(t.data_ptr())
x = torch.randn(B, 1, T, device=dev); y = torch.empty(B, C, T, device=dev); g = torch.randn(B, C, T, device=dev)
gx = torch.empty(B, 1, T, device=dev); gp = torch.empty(2, C, device=dev)
nbytes = lib.sinc_workspace_bytes(B, T, C, K); ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
for _ in range(steps):
    assert lib.sinc_forward(ctypes.byref(fb), P(x), T, P(y), 0, B, T, P(ws), nbytes, st, None) == 0
    assert lib.sinc_backward(ctypes.byref(fb), P(x), T, P(g), 0, P(gx), P(gp[0]), P(gp[1]), None, B, T, P(ws), nbytes, st, None) == 0
torch.cuda.synchronize()
print("ok", float(y.abs().max()), float(gx.abs().max()), float(gp.abs().max()))
