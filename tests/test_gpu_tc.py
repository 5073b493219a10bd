"""The tcgen05 (tensor-core) engine: operand-layout probes on hardware and parity of the TC kernels with the
oracle, at the same tolerances as the FMA engine (tests/test_gpu_parity.py)."""
import ctypes
import os
import subprocess
import sys
import json

import numpy as np
import pytest
import torch

import sinc_oracle as O
import sincformer_metacog_speech_enhancement_b200 as sb
from sincformer_metacog_speech_enhancement_b200 import _lib
from conftest import ROOT, max_rel
from test_gpu_parity import make_pair, run_dut, check_against, dev, TOL_Y

pytestmark = pytest.mark.gpu


@pytest.fixture(params=["tc", "tc_tf32"])
def tc_engine(request):
    """Both tensor-core engines: 'tc' (16-bit split operands: scaled fp16 forward, bf16 hi/mid backward) and
    'tc_tf32' (3xTF32 split in every pass)."""
    sb.set_engine(request.param)
    yield request.param
    sb.set_engine("auto")


def no_trap():
    """A tcgen05 kernel whose pipeline barrier times out executes a trap (sticky CUDA error): a clean synchronize
    after the launches is the 'no protocol failure' check -- there is no status word that could be forgotten."""
    torch.cuda.synchronize()


@pytest.mark.parametrize("case", ["kmajor_ss", "toeplitz_b", "toeplitz_b_n128", "toeplitz_b_n256", "tmem_a_toeplitz_b"])
def test_operand_layout_probe(case):
    """In-place Toeplitz B descriptors (overlapping core matrices) and the TMEM A operand, bit-exact on
    integer-valued data (tools/tc_probe.py)."""
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "tc_probe.py"), "--one", case],
                       capture_output=True, text=True, timeout=300)
    lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
    assert lines, p.stderr[-2000:]
    res = json.loads(lines[-1])
    assert res["rc"] == 0 and res["status"] == 0 and res["mismatches"] == 0, res


TC_FWD_SHAPES = [
    (64, 251, 16000, 2, 1000),
    (64, 251, 8000, 1, 16000),      # BASELINE cfg 1
    (64, 251, 16000, 3, 4100),      # last tile of each row is partial
    (64, 251, 8000, 5, 100),        # T < K
    (64, 251, 8000, 2, 4),
    (8, 250, 8000, 2, 2048),
    (70, 31, 44100, 2, 776),        # 3 filter blocks, the last one ragged
    (130, 9, 8000, 1, 600),
    (32, 253, 16000, 2, 1536),      # largest K the 256-deep GEMM holds
    (4, 1, 8000, 2, 64),
    (5, 3, 8000, 3, 52),
]


@pytest.mark.parametrize("C,K,sr,B,T", TC_FWD_SHAPES)
def test_tc_forward_vs_oracle(tc_engine, C, K, sr, B, T):
    orc, dut = make_pair(C, K, sr)
    gen = torch.Generator().manual_seed(C * 11 + T)
    with torch.no_grad():
        orc.low_hz_.mul_(1 + 0.1 * (2 * torch.rand(orc.low_hz_.shape, generator=gen) - 1))
        orc.band_hz_.mul_(1 + 0.1 * (2 * torch.rand(orc.band_hz_.shape, generator=gen) - 1))
    dut.load_state_dict(orc.state_dict())
    x = torch.randn(B, 1, T, generator=gen)
    with torch.no_grad(), dut.capture() as rec:
        y = dut(x.to(dev()))
    no_trap()
    assert rec["fwd"] == ("tc16" if tc_engine == "tc" and K <= 265 else "tc_tf32"), rec
    y32 = orc(x).detach()
    y64 = O.full_backward(orc, x, torch.zeros(B, C, T), torch.float64)["y"]
    gap = max_rel(y32, y64)
    assert max_rel(y.cpu(), y64) <= TOL_Y
    assert max_rel(y.cpu(), y32) <= TOL_Y + (gap if gap > 0.5 * TOL_Y else 0.0)


def test_tc_forward_full_size_matches_fma_engine(tc_engine):
    torch.manual_seed(0)
    dut = sb.SincConv1d(64, 251, sample_rate=16000).to(dev())
    x = torch.randn(32, 1, 48000, device=dev())
    with torch.no_grad():
        y_tc = dut(x)
        sb.set_engine("fma")
        y_fma = dut(x)
        sb.set_engine(tc_engine)
    no_trap()
    assert max_rel(y_tc, y_fma) <= 5e-6        # two independent fp32-grade evaluations (~2e-6 each vs exact)
    orc = O.OracleSincConv1d(64, 251, sample_rate=16000)
    assert max_rel(y_tc[31].cpu(), orc(x[31:32].cpu()).detach()[0]) <= TOL_Y


def test_tc_engine_refuses_unsupported_shapes(tc_engine):
    big = sb.SincConv1d(4, 301, sample_rate=8000).to(dev())
    with pytest.raises(RuntimeError, match="engine"):
        big(torch.randn(1, 1, 1024, device=dev()))          # K + 3 > 256: no tensor-core forward holds these taps
    dut = sb.SincConv1d(8, 31, sample_rate=8000).to(dev())
    x = torch.randn(1, 1, 4099, device=dev())
    if tc_engine == "tc_tf32":
        with pytest.raises(RuntimeError, match="engine"):
            dut(x)                                          # 3xTF32 forward: rows must be 16-byte aligned (T % 4 == 0)
    else:
        with dut.capture() as rec:                          # 16-bit forward: any length (plain-store epilogue for ragged rows)
            y = dut(x)
        assert rec["fwd"] == "tc16"
        orc = O.OracleSincConv1d(8, 31, sample_rate=8000)
        assert max_rel(y.cpu(), orc(x.cpu()).detach()) <= TOL_Y


@pytest.mark.parametrize("T", [48001, 47999, 4098, 897, 5])
@pytest.mark.parametrize("out_dtype", [torch.float32, torch.bfloat16])
def test_tc16_forward_ragged_lengths(T, out_dtype):
    """Utterance lengths that are not a multiple of 4 (fp32) / 8 (bf16): the rows of y are not 16-byte aligned, the
    16-bit tensor-core forward writes them with plain stores instead of bulk stores -- no drop to the FMA engine."""
    orc, dut = make_pair(64, 251, 16000, out_dtype=out_dtype)
    x = torch.randn(3, 1, T, generator=torch.Generator().manual_seed(T))
    with torch.no_grad(), dut.capture() as rec:
        y = dut(x.to(dev()))
    no_trap()
    assert rec["fwd"] == "tc16", rec
    ref = orc(x).detach()
    if out_dtype == torch.float32:
        assert max_rel(y.cpu(), ref) <= TOL_Y
    else:
        assert bool(((y.float().cpu() - ref).abs() <= 2.0 ** -8 * ref.abs() + 1e-3 * ref.abs().max()).all())


TC_BWD_SHAPES = [
    (64, 251, 16000, 2, 1000),
    (64, 251, 16000, 1, 16000),
    (64, 251, 16000, 7, 4099),      # ragged T: the backward re-pitches grad_y once and runs the 16-bit kernels on the copy
    (64, 251, 16000, 3, 4102),
    (16, 65, 8000, 5, 1001),
    (33, 127, 8000, 3, 3000),
    (64, 251, 16000, 3, 4100),
    (64, 251, 8000, 5, 100),
    (64, 251, 8000, 2, 4),
    (64, 251, 16000, 40, 2048),     # several tiles per CTA slot
    (64, 251, 16000, 24, 32768),    # > 8 tiles per CTA: two accumulation windows (double-buffered flush) in grad-filter
    (96, 251, 16000, 3, 4096),      # more than 64 filters: grad-input in two launches (blocks of 64), the second adds to gx
    (130, 129, 8000, 2, 3001),      # three blocks (64 + 64 + 2), ragged length on top
    (8, 250, 8000, 2, 2048),
    (70, 31, 44100, 2, 776),
    (130, 9, 8000, 1, 600),
    (3, 311, 16000, 2, 1536),       # forward falls back to FMA (K > 253); grad-filter TC holds K <= 312
    (4, 1, 8000, 2, 64),
]


@pytest.mark.parametrize("engine", ["tc", "tc_tf32"])
@pytest.mark.parametrize("C,K,sr,B,T", TC_BWD_SHAPES)
def test_tc_backward_vs_oracle(engine, C, K, sr, B, T):
    orc, dut = make_pair(C, K, sr)
    gen = torch.Generator().manual_seed(C * 13 + T)
    with torch.no_grad():
        orc.low_hz_.mul_(1 + 0.1 * (2 * torch.rand(orc.low_hz_.shape, generator=gen) - 1))
        orc.band_hz_.mul_(1 + 0.1 * (2 * torch.rand(orc.band_hz_.shape, generator=gen) - 1))
    dut.load_state_dict(orc.state_dict())
    x = torch.randn(B, 1, T, generator=gen)
    g = torch.randn(B, C, T, generator=gen)
    # the forward runs under AUTO (engine "tc" refuses shapes its forward cannot take); the backward under `engine`
    out = run_dut(dut, x, g, bwd_engine=engine)
    eng = out["engines"]
    # ragged lengths: under "tc" the backward re-pitches grad_y once and stays on the 16-bit kernels (needs both passes
    # to fit: K <= 256); under "tc_tf32" only grad-input takes any length
    ragged16 = engine == "tc" and T % 4 != 0 and K <= 256
    if (T % 4 == 0 or ragged16) and K <= (256 if engine == "tc" else 312):
        assert eng["bwd_filter"] == ("tc16" if engine == "tc" else "tc_tf32"), eng
    if K <= 256 and (C <= 64 or engine == "tc"):   # the 16-bit grad-input takes any filter count (blocks of 64), 3xTF32 up to 64
        assert eng["bwd_input"] == ("tc16" if engine == "tc" and (T % 4 == 0 or ragged16) else "tc_tf32"), eng
    check_against(orc, out, x, g)


def test_tc_backward_long_accumulation(tc_engine):
    """Many tiles per CTA: several accumulator flushes, fp32 accumulation chains of 4096 samples."""
    torch.manual_seed(1)
    orc, dut = make_pair(64, 251, 16000)
    B, T = 16, 48000
    x = torch.randn(B, 1, T)
    g = torch.randn(B, 64, T)
    out = run_dut(dut, x, g, need_x=False)
    o64 = O.full_backward(orc, x, g, torch.float64)
    # tensor-core fp32 accumulation truncates (round-toward-zero): ~1e-5 on G, ten times the FMA engine's error,
    # still an order of magnitude inside the 1e-4 gate
    assert max_rel(out["gfilt"], o64["gfilt"]) <= 5e-5
    assert max_rel(out["g_low"], o64["g_low"]) <= 1e-4 and max_rel(out["g_band"], o64["g_band"]) <= 1e-4


@pytest.mark.parametrize("B,T", [(3, 4096), (2, 10240), (5, 104)])
def test_tc16_bf16_activation_mode(B, T):
    """bf16-output mode on the 16-bit tensor-core kernels: bf16 y (rounded once from fp32 accumulators) and bf16
    grad_y (exact in one bf16 piece: two MMAs per product instead of three)."""
    orc, dut = make_pair(64, 251, 16000, out_dtype=torch.bfloat16)
    gen = torch.Generator().manual_seed(B * 1000 + T)
    x = torch.randn(B, 1, T, generator=gen)
    g = torch.randn(B, 64, T, generator=gen).bfloat16().float()
    sb.set_engine("tc")
    try:
        xd = x.to(dev()).requires_grad_(True)
        with dut.capture() as rec:
            y = dut(xd)
            y.backward(g.to(dev()).bfloat16())
            torch.cuda.synchronize()
    finally:
        sb.set_engine("auto")
    assert y.dtype == torch.bfloat16
    assert rec["fwd"] == rec["bwd_input"] == rec["bwd_filter"] == "tc16", rec
    ref = orc(x).detach()
    err = (y.float().cpu() - ref).abs()
    assert bool((err <= 2.0 ** -8 * ref.abs() + 1e-3 * ref.abs().max()).all())
    o32 = O.full_backward(orc, x, g, torch.float32)
    o64 = O.full_backward(orc, x, g, torch.float64)
    assert max_rel(xd.grad.cpu(), o32["gx"]) <= 1e-4
    from test_gpu_parity import param_floor
    fl, fb = param_floor(orc, o64["gfilt"].numpy())
    assert (dut.low_hz_.grad.cpu().double() - o64["g_low"]).abs().max() <= 1e-4 * o64["g_low"].abs().max() + fl
    assert (dut.band_hz_.grad.cpu().double() - o64["g_band"]).abs().max() <= 1e-4 * o64["g_band"].abs().max() + fb


def _abi_backward(lib, m, ws, x, g, want_x, want_p):
    from sincformer_metacog_speech_enhancement_b200.sinc import _bank
    B, C, T = g.shape
    fb = _bank(m.low_hz_.detach(), m.band_hz_.detach(), m.window, m.n_, m.sample_rate, m.min_low_hz, m.min_band_hz)
    P = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None
    gx = torch.empty(B, 1, T, device=dev()) if want_x else None
    gp = torch.empty(2, C, device=dev()) if want_p else None
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    stamps = torch.zeros(20, dtype=torch.int64, device=dev())
    opt = _lib.Options()
    opt.engine = _lib.ENGINE_TC
    opt.phase_stamps = stamps.data_ptr()
    rc = lib.sinc_backward(ctypes.byref(fb), P(x), T, P(g), 0, P(gx), P(gp[0]) if want_p else None,
                           P(gp[1]) if want_p else None, None, B, T, P(ws), ws.numel(), st, ctypes.byref(opt))
    assert rc == 0
    torch.cuda.synchronize()
    return gx, gp, int(stamps[18].item())


def test_filter_table_in_workspace_is_validated():
    """The grad-filter kernel leaves the synthesised filters in the workspace for the grad-input kernel
    (sinc_synth.cuh).  The consumer must take them only when they provably belong to its filter bank: same call,
    or an earlier parameter-gradient call with identical parameters -- never after the parameters changed or the
    workspace bytes were overwritten."""
    lib = _lib.load()
    sb.set_engine("tc")
    try:
        torch.manual_seed(5)
        B, C, K, T = 3, 64, 251, 4096
        m = sb.SincConv1d(C, K, sample_rate=16000).to(dev())
        x = torch.randn(B, 1, T, device=dev())
        g = torch.randn(B, C, T, device=dev())
        ws = torch.zeros(lib.sinc_workspace_bytes(B, T, C, K), dtype=torch.uint8, device=dev())
        gx_cold, _, used = _abi_backward(lib, m, ws, x, g, True, False)
        assert used == 0                                   # zeroed workspace: no table, filters synthesised in-kernel
        gx_both, gp, used = _abi_backward(lib, m, ws, x, g, True, True)
        assert used == 1                                   # one call: parameter gradients first, table consumed
        assert torch.equal(gx_both, gx_cold)               # same fp32 taps either way: bit-identical grad-input
        gx_split, _, used = _abi_backward(lib, m, ws, x, g, True, False)
        assert used == 1 and torch.equal(gx_split, gx_cold)   # data-parallel order: table of the preceding call
        # parameters change -> the header no longer matches -> synthesised in-kernel, and the result is the new one
        with torch.no_grad():
            m.band_hz_.mul_(1.03)
        gx_new, _, used = _abi_backward(lib, m, ws, x, g, True, False)
        assert used == 0 and not torch.equal(gx_new, gx_cold)
        gx_new2, _, used = _abi_backward(lib, m, ws, x, g, True, True)
        assert used == 1 and torch.equal(gx_new2, gx_new)
        # workspace bytes overwritten behind a valid header (recycled allocation): the row checksum rejects the table
        nfl = ws.numel() // 4
        wsf = ws[: nfl * 4].view(torch.float32)
        tail = wsf[-(64 * 256 + 64):]                      # inside the table rows (the table is the last region)
        tail[1000:1004] += 1.0
        gx_bad, _, used = _abi_backward(lib, m, ws, x, g, True, False)
        assert used == 0 and torch.equal(gx_bad, gx_new)
    finally:
        sb.set_engine("auto")


def test_tmem_load_16x256b_fragment():
    """The grad-input epilogue adds accumulator rows inside a thread, which relies on the register fragment of
    tcgen05.ld 16x256b: register 4j + 2u + e of thread t = lane (t / 4) + 8u, column 8j + 2(t % 4) + e."""
    lib = _lib.load_diag()
    out = torch.zeros(32 * 64, dtype=torch.int32, device=dev())
    assert lib.sinc_tc_ld16x256_probe(ctypes.c_void_p(out.data_ptr()), None) == 0
    torch.cuda.synchronize()
    got = out.cpu().numpy().reshape(32, 2, 8, 2, 2)        # [thread][16-lane half][j][u][e]
    t = np.arange(32)[:, None, None, None, None]
    half = np.arange(2)[None, :, None, None, None]
    j = np.arange(8)[None, None, :, None, None]
    u = np.arange(2)[None, None, None, :, None]
    e = np.arange(2)[None, None, None, None, :]
    want = 256 * (16 * half + t // 4 + 8 * u) + (8 * j + 2 * (t % 4) + e)
    assert np.array_equal(got, want)
