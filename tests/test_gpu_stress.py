"""Protocol stress: the warp-specialised kernels hand tiles between roles through mbarrier pipelines, and a slip in such
a protocol shows up as run-to-run varying results in whole TMEM lane quarters (16 filters) long before a tolerance
notices -- and only for some (rows, length) combinations: many utterance ends inside a CTA's range, several accumulation
windows per CTA, tail K-steps.  (A grad-filter restructuring in round 2 passed every other test in this tree and still
produced 1e-2 errors at 64 x 12 288; this file exists because of it.)  Every shape runs the whole step several times:
all runs must be bit-identical and within the parity gates of the oracle."""
import hashlib

import pytest
import torch

import sinc_oracle as O
import sincformer_metacog_speech_enhancement_b200 as sb
from conftest import max_rel
from test_gpu_parity import make_pair, dev

pytestmark = pytest.mark.gpu

SHAPES = [
    (64, 12288),    # 24 tiles per utterance, ~10 tiles per CTA: most CTAs cross an utterance end, two accumulation windows
    (64, 6144),
    (128, 6144),    # utterance ends in every CTA, several per CTA
    (64, 12800),
    (24, 32768),
    (96, 4096),     # 8 tiles per utterance: every tile boundary pattern
    (37, 20000),    # ragged tail tiles (T % 512 = 32), prime row count
    (300, 1536),    # many short utterances: three tiles each
    (11, 131072),   # long utterances: many windows per CTA, few utterance ends
]


@pytest.mark.parametrize("B,T", SHAPES)
@pytest.mark.parametrize("engine", ["tc", "tc_tf32"])
def test_step_is_reproducible_and_correct(engine, B, T):
    orc, dut = make_pair(64, 251, 16000)
    dut.engine = engine
    gen = torch.Generator().manual_seed(B * 7 + T)
    x = torch.randn(B, 1, T, generator=gen)
    g = torch.randn(B, 64, T, generator=gen)
    xd, gd = x.to(dev()), g.to(dev())
    sigs, last = set(), None
    for _ in range(5):
        xr = xd.clone().requires_grad_(True)
        dut.low_hz_.grad = dut.band_hz_.grad = None
        with dut.capture(grad_filters=True) as rec:
            y = dut(xr)
            y.backward(gd)
            torch.cuda.synchronize()
        last = dict(y=y.detach(), gx=xr.grad, G=rec["grad_filters"], gl=dut.low_hz_.grad.clone(), gb=dut.band_hz_.grad.clone())
        h = hashlib.sha1()
        for t in last.values():
            h.update(t.cpu().numpy().tobytes())
        sigs.add(h.hexdigest())
    assert len(sigs) == 1, f"{len(sigs)} different results in 5 runs: a protocol race"
    o32 = O.full_backward(orc, x, g, torch.float32)
    o64 = O.full_backward(orc, x, g, torch.float64)
    assert max_rel(last["y"].cpu(), o32["y"]) <= 1e-5
    assert max_rel(last["gx"].cpu(), o32["gx"]) <= 1e-4
    assert max_rel(last["G"].cpu(), o32["gfilt"]) <= 1e-4
    assert max_rel(last["gl"].cpu(), o64["g_low"]) <= 1e-4 and max_rel(last["gb"].cpu(), o64["g_band"]) <= 1e-4


@pytest.mark.parametrize("B,T", [(64, 12288), (16, 960000 // 4)])
def test_bf16_step_is_reproducible(B, T):
    dut = sb.SincConv1d(64, 251, sample_rate=16000, out_dtype=torch.bfloat16, engine="tc").to(dev())
    gen = torch.Generator(device=dev()).manual_seed(B + T)
    x = torch.randn(B, 1, T, device=dev(), generator=gen)
    g = torch.randn(B, 64, T, device=dev(), generator=gen).bfloat16()
    sigs = set()
    for _ in range(4):
        xr = x.clone().requires_grad_(True)
        dut.low_hz_.grad = dut.band_hz_.grad = None
        y = dut(xr)
        y.backward(g)
        torch.cuda.synchronize()
        h = hashlib.sha1()
        for t in (y.detach().float(), xr.grad, dut.low_hz_.grad, dut.band_hz_.grad):
            h.update(t.cpu().numpy().tobytes())
        sigs.add(h.hexdigest())
    assert len(sigs) == 1


def _step(dut, xd, gd):
    xr = xd.clone().requires_grad_(True)
    dut.low_hz_.grad = dut.band_hz_.grad = None
    y = dut(xr)
    y.backward(gd)
    return y.detach(), xr.grad, dut.low_hz_.grad.clone(), dut.band_hz_.grad.clone()


def test_cuda_graph_replay_of_the_step():
    """The whole step (forward, grad-filter with its fused finalize, grad-input) captured in a CUDA graph and replayed:
    every replay launches the grad-filter kernel with the SAME arguments -- including the tag of its arrival word, which
    therefore has to leave the word in a state the next replay can start from -- and must reproduce the eager result bit
    for bit on new data."""
    orc, dut = make_pair(64, 251, 16000)
    B, T = 6, 16000
    gen = torch.Generator().manual_seed(5)
    xs = [torch.randn(B, 1, T, generator=gen).to(dev()) for _ in range(4)]
    gs = [torch.randn(B, 64, T, generator=gen).to(dev()) for _ in range(4)]
    eager = [_step(dut, x, g) for x, g in zip(xs, gs)]
    torch.cuda.synchronize()

    x_static = xs[0].clone().requires_grad_(True)
    g_static = gs[0].clone()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            dut.low_hz_.grad = dut.band_hz_.grad = x_static.grad = None
            dut(x_static).backward(g_static)
    torch.cuda.current_stream().wait_stream(side)
    dut.low_hz_.grad = dut.band_hz_.grad = x_static.grad = None
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        y_static = dut(x_static)
        y_static.backward(g_static)
    for i in (1, 2, 3, 1, 0):
        with torch.no_grad():
            x_static.copy_(xs[i])
            g_static.copy_(gs[i])
        graph.replay()
        torch.cuda.synchronize()
        y, gx, gl, gb = eager[i]
        assert torch.equal(y_static.detach(), y) and torch.equal(x_static.grad, gx)
        assert torch.equal(dut.low_hz_.grad, gl) and torch.equal(dut.band_hz_.grad, gb), f"replay of batch {i}"


@pytest.mark.parametrize("fill", [0x00, 0xFF, 0x5A])
def test_fused_finalize_starts_from_any_workspace_contents(fill):
    """The arrival word of the fused finalize lives in the caller's workspace and is never initialised by the host: whatever
    a fresh (or recycled) allocation holds, the first CTA to arrive replaces it."""
    from sincformer_metacog_speech_enhancement_b200 import _lib
    from test_gpu_tc import _abi_backward
    lib = _lib.load()
    orc, dut = make_pair(64, 251, 16000)
    gen = torch.Generator().manual_seed(9)
    x = torch.randn(8, 1, 8192, generator=gen)
    g = torch.randn(8, 64, 8192, generator=gen)
    xd, gd = x.to(dev()), g.to(dev())
    ws = torch.zeros(lib.sinc_workspace_bytes(8, 8192, 64, 251), dtype=torch.uint8, device=dev())
    _, ref, _ = _abi_backward(lib, dut, ws, xd, gd, False, True)
    for _ in range(3):          # the same workspace again and again: every launch leaves the word reusable
        _, again, _ = _abi_backward(lib, dut, ws, xd, gd, False, True)
        assert torch.equal(again, ref)
    ws.fill_(fill)
    _, out, _ = _abi_backward(lib, dut, ws, xd, gd, False, True)
    assert torch.equal(out, ref)
    o64 = O.full_backward(orc, x, g, torch.float64)
    assert max_rel(out[0].cpu(), o64["g_low"].flatten()) <= 1e-4 and max_rel(out[1].cpu(), o64["g_band"].flatten()) <= 1e-4


@pytest.mark.parametrize("B,T,C", [(1, 512, 64), (1, 2048, 64), (3, 1024, 40), (2, 5120, 64), (1, 40000, 17)])
def test_fused_finalize_with_fewer_ctas_than_filters(B, T, C):
    """Small inputs launch fewer grad-filter CTAs than there are filters (one CTA per 512-sample tile): the finalising CTAs
    then take several filters each."""
    orc, dut = make_pair(C, 251, 16000)
    gen = torch.Generator().manual_seed(B + T + C)
    x = torch.randn(B, 1, T, generator=gen)
    g = torch.randn(B, C, T, generator=gen)
    outs = []
    for _ in range(3):
        with dut.capture() as rec:
            outs.append(_step(dut, x.to(dev()), g.to(dev())))
            torch.cuda.synchronize()
        assert rec["bwd_filter"] == "tc16"
    for o in outs[1:]:
        for a, b in zip(outs[0], o):
            assert torch.equal(a, b)
    o64 = O.full_backward(orc, x, g, torch.float64)
    floor = 1e-4
    assert max_rel(outs[0][2].cpu(), o64["g_low"]) <= floor and max_rel(outs[0][3].cpu(), o64["g_band"]) <= floor
