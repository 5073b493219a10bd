"""Pin oracle/sinc_oracle.py against the UNMODIFIED reference (build container only; needs /root/reference).

   python oracle/pin_against_reference.py

Asserts, for a matrix of shapes / sample rates / parameter settings:
  * constructor state bit-identical (low_hz_, band_hz_, window, n_, kernel_size, sample_rate);
  * forward output and autograd gradients of the oracle module == the reference module, bit for bit,
    in fp32 and in fp64 (same torch ops in the same order);
  * the closed-form synthesis VJP (oracle.synthesis_vjp) == fp64 autograd of the reference to 1e-8.
"""
import copy
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.environ.get("SINC_REFERENCE", "/root/reference"))
from agents.perception import SincConv1d as Ref  # noqa: E402
import sinc_oracle as O  # noqa: E402


def main():
    n = 0
    for (C, K, sr) in [(64, 251, 8000), (64, 251, 16000), (1, 251, 8000), (8, 250, 8000), (4, 1, 8000),
                       (5, 3, 8000), (70, 31, 44100), (64, 251, None), (64, 251, 0)]:
        ref = Ref(C, K, sample_rate=sr)
        orc = O.OracleSincConv1d(C, K, sample_rate=sr)
        assert ref.kernel_size == orc.kernel_size and ref.sample_rate == orc.sample_rate
        for k, v in ref.state_dict().items():
            assert torch.equal(v, orc.state_dict()[k]), (C, K, sr, k)
        for mode in ("init", "perturbed"):
            if mode == "perturbed":
                g = torch.Generator().manual_seed(7)
                with torch.no_grad():
                    ref.low_hz_.mul_(1 + 0.2 * (2 * torch.rand(ref.low_hz_.shape, generator=g) - 1))
                    ref.band_hz_.mul_(-(1 + 0.2 * (2 * torch.rand(ref.band_hz_.shape, generator=g) - 1)))
                    if C > 2:
                        ref.band_hz_[1] = 1e5
                        ref.low_hz_[2] = 0.0
                orc.load_state_dict(ref.state_dict())
            for (B, T) in [(1, 1), (2, 100), (1, 251), (3, 1000)]:
                gen = torch.Generator().manual_seed(B * 1000 + T)
                x = torch.randn(B, 1, T, generator=gen)
                gy = torch.randn(B, C, T, generator=gen)
                for dt in (torch.float32, torch.float64):
                    r = copy.deepcopy(ref).to(dt)
                    xr = x.to(dt).clone().requires_grad_(True)
                    yr = r(xr)
                    yr.backward(gy.to(dt))
                    o = O.full_backward(orc, x, gy, dt)
                    assert torch.equal(yr.detach(), o["y"]), ("y", C, K, sr, B, T, dt)
                    assert torch.equal(xr.grad, o["gx"]), ("gx", C, K, sr, B, T, dt)
                    assert torch.equal(r.low_hz_.grad, o["g_low"]), ("g_low", C, K, sr, B, T, dt)
                    assert torch.equal(r.band_hz_.grad, o["g_band"]), ("g_band", C, K, sr, B, T, dt)
                    if dt == torch.float64:
                        gl, gb = O.synthesis_vjp(o["gfilt"].numpy(), orc.low_hz_, orc.band_hz_, orc.window, orc.n_,
                                                 orc.sample_rate, orc.min_low_hz, orc.min_band_hz)
                        # absolute floor: for tiny T the gradient is the round-off of an exact cancellation
                        floor = 1e-13 * float(o["gfilt"].abs().max())
                        for mine, theirs, nm in ((gl, r.low_hz_.grad, "vjp low"), (gb, r.band_hz_.grad, "vjp band")):
                            err = float((torch.as_tensor(mine) - theirs).abs().max())
                            assert err <= 1e-8 * float(theirs.abs().max()) + floor, (nm, C, K, sr, B, T, err)
                        gxd, gfd = O.sinc_backward_conv_direct(x, copy.deepcopy(orc).double().filters(), gy)
                        assert O.max_rel(gxd, xr.grad) < 1e-10 and O.max_rel(gfd, o["gfilt"]) < 1e-10
                    n += 1
    print(f"oracle pinned against the reference: {n} (shape, params, dtype) cases bit-identical")


if __name__ == "__main__":
    main()
