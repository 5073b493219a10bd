"""Fixed (prologue + tail) vs per-utterance cost of the three tcgen05 kernels: time each through the C ABI at several
batch sizes (T = 48000) and fit t = a + b*B.  `a` is what filter synthesis, operand setup and the drain cost."""
import ctypes
import json
import sys

import torch

import sincformer_metacog_speech_enhancement_b200 as sb
from sincformer_metacog_speech_enhancement_b200 import _lib
from sincformer_metacog_speech_enhancement_b200.sinc import _bank


def main():
    dev = torch.device("cuda:0")
    lib = _lib.load()
    C, K, T = 64, 251, 48000
    m = sb.SincConv1d(C, K, sample_rate=16000).to(dev)
    fb = _bank(m.low_hz_.detach(), m.band_hz_.detach(), m.window, m.n_, m.sample_rate, m.min_low_hz, m.min_band_hz)
    P = lambda t: ctypes.c_void_p(t.data_ptr())
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    out = {}
    for B in (37, 74, 148):   # whole multiples of rows per SM pair: no load-imbalance term in the fit
        x = torch.randn(B, 1, T, device=dev)
        g = torch.randn(B, C, T, device=dev)
        y = torch.empty(B, C, T, device=dev)
        gx = torch.empty(B, 1, T, device=dev)
        gp = torch.empty(2, C, device=dev)
        nbytes = lib.sinc_workspace_bytes(B, T, C, K)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        fns = {
            "fwd": lambda: lib.sinc_forward(ctypes.byref(fb), P(x), T, P(y), 0, B, T, P(ws), nbytes, stream, None),
            "bwd_input": lambda: lib.sinc_backward(ctypes.byref(fb), P(x), T, P(g), 0, P(gx), None, None, None, B, T,
                                                   P(ws), nbytes, stream, None),
            "bwd_filter": lambda: lib.sinc_backward(ctypes.byref(fb), P(x), T, P(g), 0, None, P(gp[0]), P(gp[1]), None,
                                                    B, T, P(ws), nbytes, stream, None),
            "bwd_both": lambda: lib.sinc_backward(ctypes.byref(fb), P(x), T, P(g), 0, P(gx), P(gp[0]), P(gp[1]), None,
                                                  B, T, P(ws), nbytes, stream, None),
        }
        for name, fn in fns.items():
            for _ in range(3):
                assert fn() == 0
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 10
            e0.record()
            for _ in range(reps):
                fn()
            e1.record()
            torch.cuda.synchronize()
            out.setdefault(name, {})[B] = e0.elapsed_time(e1) / reps
        del x, g, y
    res = {}
    for name, d in out.items():
        b = (d[148] - d[37]) / 111.0
        a = d[37] - 37 * b
        res[name] = {"ms": d, "fixed_ms": a, "per_utterance_ms": b, "mid_residual_ms": d[74] - (a + 74 * b)}
    print(json.dumps(res))


if __name__ == "__main__":
    sys.exit(main())
