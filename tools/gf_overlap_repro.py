"""Does the grad-filter flush/MMA overlap still corrupt the result?  (-DSINC_ABLATION build; SINC_ABLATE_GF=8 removes the
issuers' wait for the previous window's flush.)  Prints per configuration the max error of G against the oracle over REPS
repetitions and which filters are off."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
import sinc_oracle as O
import sincformer_metacog_speech_enhancement_b200 as sb
dev = torch.device("cuda:0")
orc = O.OracleSincConv1d(64, 251, sample_rate=16000)
REPS = int(os.environ.get("GF_REPS", "6"))
NEED_X = os.environ.get("GF_NEED_X", "1") == "1"
SHAPES = json.loads(os.environ.get("GF_SHAPES", "[[64, 12288]]"))
for (B, T) in SHAPES:
    torch.manual_seed(1)
    x = torch.randn(B, 1, T); g = torch.randn(B, 64, T)
    R = O.full_backward(orc, x, g, torch.float32)["gfilt"].double().reshape(64, 251)
    dut = sb.SincConv1d(64, 251, sample_rate=16000, engine="tc").to(dev)
    xd = x.to(dev); gd = g.to(dev)
    sigs = set(); worst = 0.0; bad_filters = set()
    for rep in range(REPS):
        xr = xd.clone().requires_grad_(NEED_X)
        dut.low_hz_.grad = dut.band_hz_.grad = None
        with dut.capture(grad_filters=True) as rec:
            dut(xr).backward(gd); torch.cuda.synchronize()
        G = rec["grad_filters"].cpu().double().reshape(64, 251)
        E = (G - R).abs().max(1).values / R.abs().max()
        worst = max(worst, float(E.max()))
        bad_filters |= {c for c in range(64) if E[c] > 1e-4}
        sigs.add(hash(G.numpy().tobytes()))
    print(json.dumps({"B": B, "T": T, "mask": os.environ.get("SINC_ABLATE_GF", "0"), "need_x": NEED_X, "reps": REPS, "distinct_results": len(sigs),
                      "worst_rel_err": worst, "filters_over_1e-4": sorted(bad_filters)}), flush=True)
