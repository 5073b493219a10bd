"""Measured context numbers on the GPU box (not part of the product path):
  * cuBLAS FP32 (SIMT) and TF32 8192^3 matmul rates -> practical FP32-FMA / TF32 pipe ceilings,
  * the bar to beat: torch's own conv1d (cuDNN) on the same SincConv shapes, TF32 on and off, fwd+bwd."""
import json
import sys
import time

import torch
import torch.nn.functional as F


def ev_time(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def main():
    out = {"gpu": torch.cuda.get_device_name(0)}
    n = 8192
    a = torch.randn(n, n, device="cuda")
    b = torch.randn(n, n, device="cuda")
    for tf32 in (False, True):
        torch.backends.cuda.matmul.allow_tf32 = tf32
        ms = ev_time(lambda: a @ b)
        out["matmul_tf32_tflops" if tf32 else "matmul_fp32_tflops"] = 2 * n ** 3 / ms / 1e9
    ah, bh = a.bfloat16(), b.bfloat16()
    out["matmul_bf16_tflops"] = 2 * n ** 3 / ev_time(lambda: ah @ bh) / 1e9
    del a, b, ah, bh

    B, T, C, K = 32, 48000, 64, 251
    x = torch.randn(B, 1, T, device="cuda", requires_grad=True)
    w = torch.randn(C, 1, K, device="cuda", requires_grad=True)
    g = torch.randn(B, C, T, device="cuda")
    for tf32 in (True, False):
        torch.backends.cudnn.allow_tf32 = tf32

        def step():
            x.grad = None
            w.grad = None
            y = F.conv1d(x, w, stride=1, padding=K // 2)
            y.backward(g)
        ms = ev_time(step, reps=5, warm=2)
        out[f"cudnn_conv1d_fwd_bwd_ms_tf32_{tf32}"] = ms
        out[f"cudnn_conv1d_audio_s_per_s_tf32_{tf32}"] = B * T / 16000 / (ms * 1e-3)
        with torch.no_grad():
            msf = ev_time(lambda: F.conv1d(x, w, stride=1, padding=K // 2), reps=5, warm=2)
        out[f"cudnn_conv1d_fwd_ms_tf32_{tf32}"] = msf
    # HBM write-only and read-only streams (the forward is a 393 MB write stream, the backward kernels read streams)
    buf = torch.empty(B * C * T, device="cuda")
    ms_w = ev_time(lambda: buf.fill_(1.0))
    out["hbm_write_only_gbs"] = buf.numel() * 4 / ms_w / 1e6
    ms_r = ev_time(lambda: buf.sum())
    out["hbm_read_only_gbs_torch_sum"] = buf.numel() * 4 / ms_r / 1e6
    dst = torch.empty_like(buf)
    ms_c = ev_time(lambda: dst.copy_(buf))
    out["hbm_copy_gbs"] = 2 * buf.numel() * 4 / ms_c / 1e6
    print(json.dumps(out))


if __name__ == "__main__":
    main()
