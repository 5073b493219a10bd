"""Grad-input prologue split (ablation build, SINC_ABLATE_BI=64 prints clock deltas from CTAs 0 and 100)."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("SINC_ABLATE_BI", "64")
import torch
import sincformer_metacog_speech_enhancement_b200 as sb
from sincformer_metacog_speech_enhancement_b200 import _lib
from sincformer_metacog_speech_enhancement_b200.sinc import _bank
dev = torch.device("cuda:0"); lib = _lib.load()
C, K, T, B = 64, 251, 48000, 32
m = sb.SincConv1d(C, K, sample_rate=16000).to(dev)
fb = _bank(m.low_hz_.detach(), m.band_hz_.detach(), m.window, m.n_, m.sample_rate, m.min_low_hz, m.min_band_hz)
P = lambda t: ctypes.c_void_p(t.data_ptr())
x = torch.randn(B, 1, T, device=dev); g = torch.randn(B, C, T, device=dev)
gx = torch.empty(B, 1, T, device=dev); gp = torch.empty(2, C, device=dev)
nbytes = lib.sinc_workspace_bytes(B, T, C, K); ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
for _ in range(3):
    assert lib.sinc_backward(ctypes.byref(fb), P(x), T, P(g), 0, P(gx), P(gp[0]), P(gp[1]), None, B, T, P(ws), nbytes, st, None) == 0
torch.cuda.synchronize()
