"""Accuracy of the parameter gradients per engine against the fp64 oracle (run on the GPU box)."""
import sys, os, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
import sinc_oracle as O
import sincformer_metacog_speech_enhancement_b200 as sb

B, T = int(sys.argv[1]) if len(sys.argv) > 1 else 8, int(sys.argv[2]) if len(sys.argv) > 2 else 48000
torch.manual_seed(0)
orc = O.OracleSincConv1d(64, 251, sample_rate=16000)
x = torch.randn(B, 1, T); g = torch.randn(B, 64, T)
o64 = O.full_backward(orc, x, g, torch.float64)
o32 = O.full_backward(orc, x, g, torch.float32)
mr = lambda a, b: float((a.double() - b.double()).abs().max() / b.double().abs().max())
print(json.dumps({"ref32_vs_64": {k: mr(o32[k], o64[k]) for k in ("y", "gx", "gfilt", "g_low", "g_band")}}))
dev = torch.device("cuda:0")
for eng in ("fma", "tc_tf32", "tc"):
    sb.set_engine(eng)
    dut = sb.SincConv1d(64, 251, sample_rate=16000).to(dev)
    xd = x.to(dev).requires_grad_(True)
    with dut.capture(grad_filters=True) as rec:
        y = dut(xd); y.backward(g.to(dev)); torch.cuda.synchronize()
    out = {"y": mr(y.detach().cpu(), o64["y"]), "gx": mr(xd.grad.cpu(), o64["gx"]),
           "gfilt": mr(rec["grad_filters"].cpu(), o64["gfilt"]),
           "g_low": mr(dut.low_hz_.grad.cpu(), o64["g_low"]), "g_band": mr(dut.band_hz_.grad.cpu(), o64["g_band"])}
    print(json.dumps({eng: out}))
sb.set_engine("auto")
if len(sys.argv) > 3:
    sb.set_engine(sys.argv[3])
    dut = sb.SincConv1d(64, 251, sample_rate=16000).to(dev)
    with dut.capture(grad_filters=True) as rec:
        y = dut(x.to(dev)); y.backward(g.to(dev)); torch.cuda.synchronize()
    G = rec["grad_filters"].cpu().double().reshape(64, 251); R = o64["gfilt"].double().reshape(64, 251)
    E = (G - R).abs() / R.abs().max()
    print("err by tap mod 8:", [float(E[:, r::8].max()) for r in range(8)])
    print("err by filter (first 8):", [float(E[c].max()) for c in range(8)], "max over filters", float(E.max(1).values.max()), "min", float(E.max(1).values.min()))
    kk = E.max(0).values
    print("worst taps:", torch.topk(kk, 8).indices.tolist(), "best:", torch.topk(-kk, 4).indices.tolist())
    sb.set_engine("auto")
if len(sys.argv) > 3:
    print("per-filter max err x1e5:", [round(float(E[c].max()) * 1e5, 1) for c in range(64)])
    cw = int(E.max(1).values.argmax())
    print("worst filter", cw, "taps err x1e5:", [round(float(v) * 1e5, 1) for v in E[cw, 100:140]])
    # does the error look like a missing/extra contribution of specific samples?  project the error on per-utterance tails
    D = (G - R)[cw]
    print("err sign pattern (taps 4..7 mod 8, first 24):", [round(float(D[k]) / float(R.abs().max()) * 1e5, 1) for k in range(4, 200, 8)][:24])
