"""Temporary: time the forward kernel with parts of it switched off (SINC_FWD_DEBUG bits)."""
import ctypes, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if len(sys.argv) > 1:
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
    nbytes = lib.sinc_workspace_bytes(B, T, C, K); ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    fn = lambda: lib.sinc_forward(ctypes.byref(fb), P(x), T, P(y), 0, B, T, P(ws), nbytes, st)
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): fn()
    e1.record(); torch.cuda.synchronize()
    print("debug", os.environ.get("SINC_FWD_DEBUG"), "fwd ms", round(e0.elapsed_time(e1) / 20, 4), flush=True)
else:
    for d in (0, 1, 2, 4, 8, 3, 6, 7, 15):
        subprocess.run([sys.executable, __file__, "x"], env=dict(os.environ, SINC_FWD_DEBUG=str(d)), timeout=120)
