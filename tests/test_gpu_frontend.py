"""Next row N1: fused GroupNorm(8, C) + GELU behind the SincConv layer, against the reference's own
ops (oracle SincConv + torch.nn.GroupNorm + F.gelu on CPU; agents/perception.py:232-234).
Tolerances: activations 1e-5 max-norm relative vs the fp64 evaluation, gradients 1e-4."""
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

import sinc_oracle as O
import sincformer_metacog_speech_enhancement_b200 as sb
from conftest import max_rel

pytestmark = pytest.mark.gpu


def dev():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


@pytest.mark.parametrize("B,C,G,T", [(2, 64, 8, 4096), (3, 16, 8, 100), (1, 64, 8, 48000), (2, 24, 3, 8192 + 4), (4, 8, 1, 8),
                                     (2, 64, 8, 4099), (3, 16, 2, 16385), (2, 8, 8, 5)])   # ragged lengths: 4-byte path
def test_fused_groupnorm_gelu_vs_torch(B, C, G, T):
    gen = torch.Generator().manual_seed(B * 100 + T)
    ref = nn.GroupNorm(G, C)
    with torch.no_grad():
        ref.weight.copy_(1 + 0.3 * torch.randn(C, generator=gen))
        ref.bias.copy_(0.2 * torch.randn(C, generator=gen))
    dut = sb.FusedGroupNormGELU.from_reference(ref).to(dev())
    y = (0.7 * torch.randn(B, C, T, generator=gen) + 0.1)
    dz = torch.randn(B, C, T, generator=gen)
    r64 = nn.GroupNorm(G, C).double()
    r64.load_state_dict({k: v.double() for k, v in ref.state_dict().items()})
    y64 = y.double().requires_grad_(True)
    z64 = F.gelu(r64(y64))
    z64.backward(dz.double())
    yd = y.to(dev()).requires_grad_(True)
    z = dut(yd)
    z.backward(dz.to(dev()))
    torch.cuda.synchronize()
    assert max_rel(z.detach().cpu(), z64.detach()) <= 1e-5
    assert max_rel(yd.grad.cpu(), y64.grad) <= 1e-4
    assert max_rel(dut.weight.grad.cpu(), r64.weight.grad) <= 1e-4
    assert max_rel(dut.bias.grad.cpu(), r64.bias.grad) <= 1e-4


@pytest.mark.parametrize("B,T,sr", [(2, 4096, 16000), (3, 6000, 8000), (1, 16000, 8000)])
def test_sinc_frontend_matches_reference_ops(B, T, sr):
    """conv -> GroupNorm(8,64) -> GELU with the statistics taken from the convolution epilogue."""
    gen = torch.Generator().manual_seed(T)
    orc = O.OracleSincConv1d(64, 251, sample_rate=sr)
    gn = nn.GroupNorm(8, 64)
    with torch.no_grad():
        gn.weight.copy_(1 + 0.3 * torch.randn(64, generator=gen))
        gn.bias.copy_(0.2 * torch.randn(64, generator=gen))
    fe = sb.SincFrontEnd(64, 251, sample_rate=sr)
    fe.sinc_conv.load_state_dict(orc.state_dict())
    fe.sinc_norm.load_state_dict(gn.state_dict())
    fe = fe.to(dev())
    x = torch.randn(B, 1, T, generator=gen)
    dz = torch.randn(B, 64, T, generator=gen)
    # fp64 evaluation of the reference ops on CPU
    import copy
    o64, g64 = copy.deepcopy(orc).double(), copy.deepcopy(gn).double()
    x64 = x.double().requires_grad_(True)
    z64 = F.gelu(g64(o64(x64)))
    z64.backward(dz.double())
    n0 = sb.kernel_launch_count()
    xd = x.to(dev()).requires_grad_(True)
    z = fe(xd)
    torch.cuda.synchronize()
    assert sb.kernel_launch_count() - n0 == 2, "fused forward = convolution (with statistics) + one apply pass"
    z.backward(dz.to(dev()))
    torch.cuda.synchronize()
    assert max_rel(z.detach().cpu(), z64.detach()) <= 2e-5     # conv (1e-5 gate) seen through the normalisation
    assert max_rel(xd.grad.cpu(), x64.grad) <= 1e-4
    assert max_rel(fe.sinc_norm.weight.grad.cpu(), g64.weight.grad) <= 1e-4
    assert max_rel(fe.sinc_norm.bias.grad.cpu(), g64.bias.grad) <= 1e-4
    elo = (fe.sinc_conv.low_hz_.grad.cpu().double() - o64.low_hz_.grad).abs().max()
    eba = (fe.sinc_conv.band_hz_.grad.cpu().double() - o64.band_hz_.grad).abs().max()
    assert elo <= 1e-4 * o64.low_hz_.grad.abs().max() + 1e-6 * o64.low_hz_.grad.abs().max()
    assert eba <= 1e-4 * o64.band_hz_.grad.abs().max() + 1e-6 * o64.band_hz_.grad.abs().max()


def test_sinc_frontend_unfused_fallback_and_state_dict():
    fe = sb.SincFrontEnd(64, 251, sample_rate=8000).to(dev())
    keys = set(fe.state_dict().keys())
    assert keys == {"sinc_conv.low_hz_", "sinc_conv.band_hz_", "sinc_conv.window", "sinc_conv.n_",
                    "sinc_norm.weight", "sinc_norm.bias"}       # the slice of PerceptionAgent's state_dict
    # T % 4 != 0 (the reference's gelu(GroupNorm(conv)) takes any length): unfused conv + the 4-byte GroupNorm path
    gen = torch.Generator().manual_seed(3)
    xr = torch.randn(2, 1, 4098, generator=gen)
    orc = O.OracleSincConv1d(64, 251, sample_rate=8000).double()
    gn = nn.GroupNorm(8, 64).double()
    z64 = F.gelu(gn(orc(xr.double()))).detach()
    assert max_rel(fe(xr.to(dev())).cpu(), z64) <= 2e-5
    with pytest.raises(RuntimeError):
        fe(torch.randn(2, 1, 0, device=dev()))                   # empty input: the same rule as SincConv1d
    x = torch.randn(2, 4096, device=dev())
    sb.set_engine("fma")
    try:
        z_fma = fe(x)                                            # unfused: conv, statistics pass, apply pass
    finally:
        sb.set_engine("auto")
    z = fe(x)
    assert max_rel(z, z_fma) <= 2e-5
