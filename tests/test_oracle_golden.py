"""The oracle (oracle/sinc_oracle.py) against golden vectors produced by the unmodified reference
(oracle/make_golden.py).  CPU only.  On the machine that generated the vectors the fp32 results are
bit-identical; across CPUs oneDNN may pick another kernel, hence the (tight) tolerances."""
import numpy as np
import pytest
import torch

import sinc_oracle as O
from conftest import golden_case_names, load_golden, max_rel


def oracle_from(g):
    m = O.OracleSincConv1d(int(g["C"]), int(g["K_arg"]), sample_rate=int(g["sr"]))
    with torch.no_grad():
        m.low_hz_.copy_(torch.from_numpy(g["low_hz_"]))
        m.band_hz_.copy_(torch.from_numpy(g["band_hz_"]))
    assert torch.equal(m.window, torch.from_numpy(g["window"]))
    assert torch.equal(m.n_, torch.from_numpy(g["n_"]))
    return m


@pytest.mark.parametrize("name", golden_case_names())
def test_oracle_matches_reference_golden(name):
    g = load_golden(name)
    m = oracle_from(g)
    x, gy = torch.from_numpy(g["x"]), torch.from_numpy(g["g"])
    o32 = O.full_backward(m, x, gy, torch.float32)
    assert max_rel(o32["y"], g["y"]) <= 2e-6
    assert max_rel(o32["gx"], g["gx"]) <= 2e-6
    assert max_rel(o32["gfilt"], g["gfilt"]) <= 2e-6
    o64 = O.full_backward(m, x, gy, torch.float64)
    assert max_rel(o64["g_low"], g["g_low64"]) <= 1e-9
    assert max_rel(o64["g_band"], g["g_band64"]) <= 1e-9
    # closed-form chain rule (what the CUDA finalize kernel implements) == fp64 autograd of the reference
    gl, gb = O.synthesis_vjp(o64["gfilt"].numpy(), m.low_hz_, m.band_hz_, m.window, m.n_, m.sample_rate,
                             m.min_low_hz, m.min_band_hz)
    floor = 1e-13 * float(o64["gfilt"].abs().max())
    assert np.abs(gl - g["g_low64"]).max() <= 1e-8 * np.abs(g["g_low64"]).max() + floor
    assert np.abs(gb - g["g_band64"]).max() <= 1e-8 * np.abs(g["g_band64"]).max() + floor


def test_oracle_constructor_state(init_kat):
    for key in sorted({k.split("/")[0] for k in init_kat}):
        C, K, sr = (p[1:] if i < 1 else p for i, p in enumerate(key.split("_")))
        C, K = int(C), int(K[1:])
        sr = None if sr[2:] == "None" else int(sr[2:])
        m = O.OracleSincConv1d(C, K, sample_rate=sr)
        assert m.kernel_size == int(init_kat[f"{key}/kernel_size"])
        assert float(m.sample_rate) == float(init_kat[f"{key}/sample_rate"])
        for n in ("low_hz_", "band_hz_", "window", "n_"):
            assert np.array_equal(m.state_dict()[n].numpy(), init_kat[f"{key}/{n}"]), (key, n)
        assert max_rel(m.filters().detach(), init_kat[f"{key}/filters32"]) <= 2e-6


def test_direct_sums_agree_with_library_backward():
    m = O.OracleSincConv1d(6, 9, sample_rate=8000)
    gen = torch.Generator().manual_seed(0)
    x, gy = torch.randn(2, 1, 40, generator=gen), torch.randn(2, 6, 40, generator=gen)
    w = m.filters().detach().double()
    gx, gw = O.sinc_backward_conv(x.double(), w, gy.double())
    gxd, gwd = O.sinc_backward_conv_direct(x, w, gy)
    assert max_rel(gxd, gx) < 1e-12 and max_rel(gwd, gw) < 1e-12
