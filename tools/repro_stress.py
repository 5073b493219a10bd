"""Run-to-run bitwise reproducibility of the whole step at full size (cfg 2) -- catches protocol races that small
parity shapes do not reach (the grad-filter flush/MMA overlap of r02k was found this way)."""
import sys, os, hashlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import sincformer_metacog_speech_enhancement_b200 as sb

dev = torch.device("cuda:0")
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 6
for (B, T, dt) in ((32, 48000, torch.float32), (8, 240000, torch.float32), (32, 48000, torch.bfloat16)):
    torch.manual_seed(0)
    m = sb.SincConv1d(64, 251, sample_rate=16000, out_dtype=dt).to(dev)
    x = torch.randn(B, 1, T, device=dev)
    g = torch.randn(B, 64, T, device=dev).to(dt)
    sigs = []
    for r in range(reps):
        xr = x.clone().requires_grad_(True)
        m.low_hz_.grad = m.band_hz_.grad = None
        y = m(xr)
        y.backward(g)
        torch.cuda.synchronize()
        h = hashlib.sha1()
        for t in (y.detach(), xr.grad, m.low_hz_.grad, m.band_hz_.grad):
            h.update(t.float().cpu().numpy().tobytes())
        sigs.append(h.hexdigest()[:12])
    print(B, T, str(dt), "identical" if len(set(sigs)) == 1 else f"DIFFERENT {sigs}", flush=True)
