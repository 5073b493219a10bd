"""Forward prologue split (ablation build, SINC_ABLATE_FW=32 prints clock deltas from CTAs 0 and 77)."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("SINC_ABLATE_FW", "32")
import torch
import sincformer_metacog_speech_enhancement_b200 as sb
from sincformer_metacog_speech_enhancement_b200 import _lib
from sincformer_metacog_speech_enhancement_b200.sinc import _bank
dev = torch.device("cuda:0"); lib = _lib.load()
C, K, T, B = 64, 251, 48000, 32
m = sb.SincConv1d(C, K, sample_rate=16000).to(dev)
fb = _bank(m.low_hz_.detach(), m.band_hz_.detach(), m.window, m.n_, m.sample_rate, m.min_low_hz, m.min_band_hz)
P = lambda t: ctypes.c_void_p(t.data_ptr())
x = torch.randn(B, 1, T, device=dev); y = torch.empty(B, C, T, device=dev)
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
for _ in range(4):
    assert lib.sinc_forward(ctypes.byref(fb), P(x), T, P(y), 0, B, T, None, 0, st, None) == 0
torch.cuda.synchronize()
