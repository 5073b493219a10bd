"""N1 measurement: conv -> GroupNorm(8,64) -> GELU, fused front-end vs the same SincConv followed by torch's own
GroupNorm + GELU, forward + backward at cfg 2 (B=32, T=48000).  One JSON line."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.nn as nn, torch.nn.functional as F
import sincformer_metacog_speech_enhancement_b200 as sb

dev = torch.device("cuda:0")
B, T, C = 32, 48000, 64
torch.manual_seed(0)
x = torch.randn(B, 1, T, device=dev)
dz = torch.randn(B, C, T, device=dev)
fe = sb.SincFrontEnd(C, 251, 16000).to(dev)
conv = fe.sinc_conv
gn = nn.GroupNorm(8, C).to(dev)


def timed(fn, steps=20, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def step_fused():
    for p in fe.parameters():
        p.grad = None
    xr = x.detach().requires_grad_(True)
    fe(xr).backward(dz)


def step_unfused():
    for p in list(conv.parameters()) + list(gn.parameters()):
        p.grad = None
    xr = x.detach().requires_grad_(True)
    F.gelu(gn(conv(xr))).backward(dz)


def step_conv_only():
    for p in conv.parameters():
        p.grad = None
    xr = x.detach().requires_grad_(True)
    conv(xr).backward(dz)


def fwd_fused():
    with torch.no_grad():
        fe(x)


def fwd_unfused():
    with torch.no_grad():
        F.gelu(gn(conv(x)))


def fwd_conv():
    with torch.no_grad():
        conv(x)


out = {"shape": [B, C, T]}
out["conv_only_fwd_bwd_ms"] = timed(step_conv_only)
out["fused_frontend_fwd_bwd_ms"] = timed(step_fused)
out["conv_plus_torch_gn_gelu_fwd_bwd_ms"] = timed(step_unfused)
out["conv_only_fwd_ms"] = timed(fwd_conv)
out["fused_frontend_fwd_ms"] = timed(fwd_fused)
out["conv_plus_torch_gn_gelu_fwd_ms"] = timed(fwd_unfused)
act = B * C * T * 4
out["gn_gelu_fwd_added_ms"] = out["fused_frontend_fwd_ms"] - out["conv_only_fwd_ms"]
out["gn_gelu_fwd_hbm_gbs"] = 2 * act / out["gn_gelu_fwd_added_ms"] / 1e6          # read y + write z
out["gn_gelu_bwd_added_ms"] = (out["fused_frontend_fwd_bwd_ms"] - out["fused_frontend_fwd_ms"]) - \
    (out["conv_only_fwd_bwd_ms"] - out["conv_only_fwd_ms"])
out["gn_gelu_bwd_hbm_gbs"] = 5 * act / out["gn_gelu_bwd_added_ms"] / 1e6          # (dz, y) twice + write dy
out["torch_gn_gelu_fwd_added_ms"] = out["conv_plus_torch_gn_gelu_fwd_ms"] - out["conv_only_fwd_ms"]
out["torch_gn_gelu_bwd_added_ms"] = (out["conv_plus_torch_gn_gelu_fwd_bwd_ms"] - out["conv_plus_torch_gn_gelu_fwd_ms"]) - \
    (out["conv_only_fwd_bwd_ms"] - out["conv_only_fwd_ms"])
print(json.dumps(out))
