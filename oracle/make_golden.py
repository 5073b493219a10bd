"""Generate tests/golden/*.npz from the UNMODIFIED reference module (agents/perception.py::SincConv1d).

Run in the build container only (needs /root/reference):   python oracle/make_golden.py
The reference ships no golden vectors for this path, so these files -- outputs of the
reference itself on seeded inputs -- are the pin used on machines where /root/reference does not exist.

Per case the file holds: the module state, x, g (grad_out), and
  y, gx, gfilt, g_low32, g_band32  -- reference evaluated on CPU in fp32 (the gate for y / gx / gfilt)
  g_low64, g_band64 -- copy.deepcopy(ref).double() fed x.double(), g.double() (the gate for the parameter grads)
"""
import copy
import os
import sys

import numpy as np
import torch

REF = os.environ.get("SINC_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
from agents.perception import SincConv1d as RefSincConv1d  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

# name, C, K, sr, B, T, param mode
CASES = [
    ("c64_k251_sr8000_b1_t1_init", 64, 251, 8000, 1, 1, "init"),
    ("c64_k251_sr16000_b2_t100_init", 64, 251, 16000, 2, 100, "init"),
    ("c64_k251_sr8000_b1_t251_perturbed", 64, 251, 8000, 1, 251, "perturbed"),
    ("c64_k251_sr16000_b2_t1000_edge", 64, 251, 16000, 2, 1000, "edge"),
    ("c8_k250_sr8000_b3_t4097_perturbed", 8, 250, 8000, 3, 4097, "perturbed"),
    ("c5_k3_sr8000_b2_t37_init", 5, 3, 8000, 2, 37, "init"),
    ("c4_k1_sr8000_b2_t19_init", 4, 1, 8000, 2, 19, "init"),
    ("c1_k251_sr8000_b1_t600_init", 1, 251, 8000, 1, 600, "init"),
    ("c70_k31_sr44100_b2_t333_perturbed", 70, 31, 44100, 2, 333, "perturbed"),
]

INIT_KATS = [(64, 251, 8000), (64, 251, 16000), (1, 251, 8000), (8, 250, 8000), (4, 1, 8000), (64, 251, None)]


def set_params(m, mode, seed):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        if mode == "perturbed":
            m.low_hz_.mul_(1 + 0.2 * (2 * torch.rand(m.low_hz_.shape, generator=g) - 1))
            m.band_hz_.mul_(1 + 0.2 * (2 * torch.rand(m.band_hz_.shape, generator=g) - 1))
        elif mode == "edge":
            m.low_hz_.mul_(1 + 0.1 * (2 * torch.rand(m.low_hz_.shape, generator=g) - 1))
            m.low_hz_[3] = -m.low_hz_[3]          # negative parameter: |.| and sign flip
            m.low_hz_[7] = 0.0                    # sign(0) = 0
            m.band_hz_[11] = 0.0
            m.band_hz_[20] = -m.band_hz_[20]
            m.band_hz_[40] = 1.0e5               # clamped at Nyquist -> zero band gradient
            # exactly at the clamp boundary: low + min_band + |band| == sr/2 passes the gradient
            lowv = m.min_low_hz + m.low_hz_[50].abs()
            m.band_hz_[50] = m.sample_rate / 2.0 - lowv - m.min_band_hz


def run_case(name, C, K, sr, B, T, mode, idx):
    ref = RefSincConv1d(C, K, sample_rate=sr)
    set_params(ref, mode, 1000 + idx)
    gx_ = torch.Generator().manual_seed(2000 + idx)
    x = torch.randn(B, 1, T, generator=gx_)
    g = torch.randn(B, C, T, generator=gx_)

    def fwd_bwd(mod, dt):
        m = copy.deepcopy(mod).to(dt)
        xr = x.to(dt).clone().requires_grad_(True)
        # same graph as the reference forward, but keep a handle on the filters for grad-filter:
        import torch.nn.functional as F
        captured = {}
        orig = F.conv1d

        def spy(inp, w, *a, **k):
            w.retain_grad()
            captured["w"] = w
            return orig(inp, w, *a, **k)
        F.conv1d = spy
        try:
            y = m(xr)
        finally:
            F.conv1d = orig
        y.backward(g.to(dt))
        return (y.detach(), xr.grad.detach(), captured["w"].grad.detach().squeeze(1),
                m.low_hz_.grad.detach(), m.band_hz_.grad.detach())

    y, gx, gf, gl, gb = fwd_bwd(ref, torch.float32)
    y64, gx64, gf64, gl64, gb64 = fwd_bwd(ref, torch.float64)
    np.savez_compressed(
        os.path.join(OUT, name + ".npz"),
        C=C, K=ref.kernel_size, K_arg=K, sr=sr, B=B, T=T,
        low_hz_=ref.low_hz_.detach().numpy(), band_hz_=ref.band_hz_.detach().numpy(),
        window=ref.window.numpy(), n_=ref.n_.numpy(),
        x=x.numpy(), g=g.numpy(),
        y=y.numpy(), gx=gx.numpy(), gfilt=gf.numpy(), g_low32=gl.numpy(), g_band32=gb.numpy(),
        g_low64=gl64.numpy(), g_band64=gb64.numpy(),
        y64_absmax=float(y64.abs().max()), y32_vs_y64=float((y.double() - y64).abs().max() / y64.abs().max()))
    print(f"{name}: y {tuple(y.shape)}  |y|max {y.abs().max():.3e}  g_low64 max {gl64.abs().max():.3e}")


def main():
    os.makedirs(OUT, exist_ok=True)
    for i, c in enumerate(CASES):
        run_case(*c, i)
    kat = {}
    for (C, K, sr) in INIT_KATS:
        m = RefSincConv1d(C, K, sample_rate=sr)
        key = f"C{C}_K{K}_sr{sr}"
        for n, v in m.state_dict().items():
            kat[f"{key}/{n}"] = v.numpy()
        kat[f"{key}/kernel_size"] = np.int64(m.kernel_size)
        kat[f"{key}/sample_rate"] = np.float64(m.sample_rate)
        with torch.no_grad():
            import torch.nn.functional as F
            captured = {}
            orig = F.conv1d

            def spy(inp, w, *a, **k):
                captured["w"] = w
                return orig(inp, w, *a, **k)
            F.conv1d = spy
            try:
                m(torch.zeros(1, 1, 4))
                kat[f"{key}/filters32"] = captured["w"].squeeze(1).numpy()
                copy.deepcopy(m).double()(torch.zeros(1, 1, 4, dtype=torch.float64))
                kat[f"{key}/filters64"] = captured["w"].squeeze(1).numpy()
            finally:
                F.conv1d = orig
    np.savez_compressed(os.path.join(OUT, "init_kat.npz"), **kat)
    print("init_kat:", len(kat), "arrays")


if __name__ == "__main__":
    main()
