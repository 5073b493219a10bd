"""Small forward + backward cases for compute-sanitizer (one tool per gpurun call, B200_PROFILING.md):
   compute-sanitizer --tool memcheck|racecheck|initcheck|synccheck python tools/sanitize_case.py
Every 16-bit tcgen05 kernel, the 3xTF32 kernels, the FMA kernels, finalize and the GroupNorm/GELU kernels run once on
shapes that exercise partial tiles, utterance ends inside a CTA's range and two accumulation windows."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
import sinc_oracle as O
import sincformer_metacog_speech_enhancement_b200 as sb

dev = torch.device("cuda:0")
cases = [("tc", 3, 4100 // 4 * 4), ("tc", 160, 5120), ("tc_tf32", 2, 2048), ("fma", 2, 1001), ("tc", 2, 1001)]
if len(sys.argv) > 1:
    cases = cases[: int(sys.argv[1])]
for engine, B, T in cases:
    orc = O.OracleSincConv1d(64, 251, sample_rate=16000)
    dut = sb.SincConv1d(64, 251, sample_rate=16000, engine=engine).to(dev)
    gen = torch.Generator().manual_seed(B + T)
    x = torch.randn(B, 1, T, generator=gen); g = torch.randn(B, 64, T, generator=gen)
    xr = x.to(dev).requires_grad_(True)
    with dut.capture() as rec:
        y = dut(xr); y.backward(g.to(dev)); torch.cuda.synchronize()
    ref = O.full_backward(orc, x, g, torch.float32)
    ey = float((y.detach().cpu() - ref["y"]).abs().max() / ref["y"].abs().max())
    eg = float((xr.grad.cpu() - ref["gx"]).abs().max() / ref["gx"].abs().max())
    print(engine, B, T, {k: rec[k] for k in ("fwd", "bwd_input", "bwd_filter")}, f"y {ey:.1e} gx {eg:.1e}", flush=True)
    assert ey <= 1e-5 and eg <= 1e-4
fe = sb.SincFrontEnd(64, 251, sample_rate=16000).to(dev)
z = fe(torch.randn(2, 1, 4096, device=dev, requires_grad=True)); z.sum().backward(); torch.cuda.synchronize()
print("frontend ok", flush=True)
