"""grad-filter multi-window debug: error of G per filter for several (B, T), with/without grad-input before it."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
import sinc_oracle as O
import sincformer_metacog_speech_enhancement_b200 as sb
dev = torch.device("cuda:0")
orc = O.OracleSincConv1d(64, 251, sample_rate=16000)
import json
CFG = json.loads(os.environ.get("GF_CFG", "[[16,48000,0],[16,48000,1],[32,24576,1],[64,12288,1]]"))
REPS = int(os.environ.get("GF_REPS", "2"))
for (B, T, need_x) in CFG:
    torch.manual_seed(1)
    x = torch.randn(B, 1, T); g = torch.randn(B, 64, T)
    R = O.full_backward(orc, x, g, torch.float32)["gfilt"].double().reshape(64, 251)
    sb.set_engine("tc")
    dut = sb.SincConv1d(64, 251, sample_rate=16000).to(dev)
    res = []
    for rep in range(REPS):
        xd = x.to(dev).requires_grad_(bool(need_x))
        dut.low_hz_.grad = dut.band_hz_.grad = None
        with dut.capture(grad_filters=True) as rec:
            y = dut(xd); y.backward(g.to(dev)); torch.cuda.synchronize()
        G = rec["grad_filters"].cpu().double().reshape(64, 251)
        E = (G - R).abs().max(1).values / R.abs().max()
        bad = [(c, round(float(E[c]) * 1e5, 1)) for c in range(64) if E[c] > 5e-5]
        res.append((round(float(E.max()) * 1e5, 2), bad))
    print(B, T, need_x, res, flush=True)
sb.set_engine("auto")
