"""CPU oracle for the SincConv hot path.  TEST INFRASTRUCTURE ONLY.

This file is the checker, never the product: only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it.  The shipped
package (``sincformer_metacog_speech_enhancement_b200``) must not, and does not.

What it restates (reference = MarvelMathesh/sincformer-metacog-speech-enhancement,
all citations relative to ``/root/reference``):

* ``erb_filterbank_state``   <- ``agents/perception.py:35-77``  (SincConv1d.__init__)
* ``synthesize_filters``     <- ``agents/perception.py:89-115`` (parameter -> (C,K) filter bank)
* ``sinc_forward``           <- ``agents/perception.py:117-118`` (conv1d, stride 1, pad K//2)
* ``sinc_backward_conv``     <- autograd of ``:117`` (grad-input, grad-filter)
* ``synthesis_vjp``          <- autograd of ``:89-115`` written out in closed form
* ``OracleSincConv1d``       <- the module, same ctor / parameter / buffer names

The arithmetic of the reference lives in PyTorch (``torch.sin``, ``F.conv1d`` -> oneDNN on CPU);
pinned only as ``torch>=2.0.0`` (``requirements.txt:1``), here torch 2.11.0.  The oracle calls the
same library entry points in the same order, so on CPU/fp32 it reproduces the reference bit for
bit -- that is what ``oracle/pin_against_reference.py`` asserts against the *imported, unmodified*
reference module, and what ``tests/golden/*.npz`` (generated from the reference itself by
``oracle/make_golden.py``) pins on boxes where ``/root/reference`` does not exist.

Parity status: PINNED.  The reference ships no golden vectors for this path (zero
coverage of ``agents/``); the pins are outputs of the reference itself run in the build container.

Two precisions are used by the tests:
  * ``dtype=torch.float32`` -- the gate for y, grad-input and grad-filter (1e-5 / 1e-4 / 1e-4,
    max-norm relative).
  * ``dtype=torch.float64`` -- the gate for the two parameter gradients (1e-4), because the
    reference's own fp32 gradients for ``low_hz_``/``band_hz_`` are round-off noise.
    The fp64 oracle promotes the *fp32 values* of parameters and buffers, exactly as
    ``copy.deepcopy(ref).double()`` does.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np
import torch
import torch.nn.functional as F

DEFAULT_SAMPLE_RATE = 8000  # reference config.py:17 (used when sample_rate is falsy, perception.py:41)


# --------------------------------------------------------------------------------------
# a1: constructor state (perception.py:35-77)
# --------------------------------------------------------------------------------------
def _hz_to_erb_rate(hz: float) -> float:
    return 21.4 * math.log10(1 + hz / 228.7)


@dataclass
class FilterbankState:
    out_channels: int
    kernel_size: int
    sample_rate: float
    min_low_hz: float
    min_band_hz: float
    low_hz_: torch.Tensor   # (C,1) fp32
    band_hz_: torch.Tensor  # (C,1) fp32
    window: torch.Tensor    # (K,)  fp32
    n_: torch.Tensor        # (1,(K-1)/2) fp32


def erb_filterbank_state(out_channels, kernel_size, sample_rate=None,
                         min_low_hz=50, min_band_hz=50) -> FilterbankState:
    """Deterministic ERB-rate initial state.  Restates perception.py:39-77.

    * falsy ``sample_rate`` -> 8000 (``x or default`` idiom, :41)
    * even ``kernel_size`` -> +1 (:44-45)
    * C+1 points linear on the ERB-rate scale between min_low_hz and sr/2-min_band_hz in
      float64 NumPy, mapped back to Hz (:52-59); low = points[:-1], band = diff(points),
      stored fp32 (C,1) (:62-65)
    * window[k] = 0.54-0.46*cos(2*pi*k/K), fp32 torch ops -- note /K, not /(K-1) (:68-71)
    * n_ = 2*pi*arange(-(K-1)/2, 0)/sr as (1,H) fp32 (:74-77)
    """
    sr = sample_rate or DEFAULT_SAMPLE_RATE
    K = kernel_size + 1 if kernel_size % 2 == 0 else kernel_size

    lo_edge = min_low_hz
    hi_edge = sr / 2 - min_band_hz
    erb_grid = np.linspace(_hz_to_erb_rate(lo_edge), _hz_to_erb_rate(hi_edge), out_channels + 1)
    hz_grid = 228.7 * (10 ** (erb_grid / 21.4) - 1)

    low = torch.Tensor(hz_grid[:-1]).view(-1, 1)
    band = torch.Tensor(np.diff(hz_grid)).view(-1, 1)

    taps = torch.linspace(0, K - 1, K)
    window = 0.54 - 0.46 * torch.cos(2 * math.pi * taps / K)

    half = (K - 1) / 2.0
    n_ = 2 * math.pi * torch.arange(-half, 0).view(1, -1) / sr
    return FilterbankState(out_channels, K, sr, min_low_hz, min_band_hz, low, band, window, n_)


# --------------------------------------------------------------------------------------
# a2: filter synthesis (perception.py:89-115)
# --------------------------------------------------------------------------------------
def synthesize_filters(low_hz_, band_hz_, window, n_, sample_rate, min_low_hz=50, min_band_hz=50):
    """(C,1),(C,1),(K,),(1,H) -> (C,K) in the dtype of the inputs.  Differentiable (torch ops).

    Same op order as the reference so that fp32 results are bit-identical:
    low=min_low+|low_hz_|; high=clamp(low+min_band+|band_hz_|, max=sr/2); f=hz/sr;
    left=(sin(f_high*n)-sin(f_low*n))/(n/2+1e-8); center=2*(f_high-f_low);
    u=[left,center,flip(left)]*window; w=u/(sum|u|+1e-8).
    """
    low = min_low_hz + torch.abs(low_hz_)
    high = torch.clamp(low + min_band_hz + torch.abs(band_hz_), max=sample_rate / 2.0)
    f_lo = low / sample_rate
    f_hi = high / sample_rate
    left = (torch.sin(f_hi * n_) - torch.sin(f_lo * n_)) / (n_ / 2.0 + 1e-8)
    centre = 2 * (f_hi - f_lo)
    right = torch.flip(left, dims=[1])
    u = torch.cat([left, centre, right], dim=1) * window
    return u / (u.abs().sum(dim=1, keepdim=True) + 1e-8)


# --------------------------------------------------------------------------------------
# a3 / a4: the convolution and its adjoints (perception.py:117-118 and its autograd)
# --------------------------------------------------------------------------------------
def sinc_forward(x, filters):
    """x (B,1,T), filters (C,K) -> y (B,C,T);  y[b,c,t] = sum_k w[c,k] x[b,0,t+k-K//2], zero padded."""
    K = filters.shape[1]
    return F.conv1d(x, filters.unsqueeze(1), stride=1, padding=K // 2)


def sinc_backward_conv(x, filters, g):
    """Adjoint of sinc_forward: returns (grad_x (B,1,T), grad_filters (C,K)) via torch's own
    convolution backward -- the same library path the reference's autograd takes."""
    K = filters.shape[1]
    xr = x.detach().clone().requires_grad_(True)
    wr = filters.detach().clone().requires_grad_(True)
    y = F.conv1d(xr, wr.unsqueeze(1), stride=1, padding=K // 2)
    gx, gw = torch.autograd.grad(y, (xr, wr), g)
    return gx, gw


def sinc_backward_conv_direct(x, filters, g):
    """Same adjoints from the defining sums in float64 NumPy (small cases only): an independent
    check on the library path.  gx[b,t]=sum_c sum_k w[c,k] g[b,c,t-k+H]; G[c,k]=sum_b sum_t g[b,c,t] x[b,t+k-H]."""
    x64 = x.detach().double().numpy()[:, 0, :]
    w64 = filters.detach().double().numpy()
    g64 = g.detach().double().numpy()
    B, T = x64.shape
    C, K = w64.shape
    H = K // 2
    xp = np.zeros((B, T + 2 * H))
    xp[:, H:H + T] = x64
    gp = np.zeros((B, C, T + 2 * H))
    gp[:, :, H:H + T] = g64
    G = np.zeros((C, K))
    gx = np.zeros((B, T))
    for k in range(K):
        G[:, k] = np.einsum('bct,bt->c', g64, xp[:, k:k + T])
        # t-k+H  -> padded index t - k + 2H
        gx += np.einsum('c,bct->bt', w64[:, k], gp[:, :, 2 * H - k:2 * H - k + T])
    return gx[:, None, :], G


# --------------------------------------------------------------------------------------
# a5: closed-form VJP of the synthesis (autograd of perception.py:89-115), float64 NumPy
# --------------------------------------------------------------------------------------
def synthesis_vjp(G, low_hz_, band_hz_, window, n_, sample_rate, min_low_hz=50, min_band_hz=50):
    """grad-filter (C,K) -> (grad low_hz_ (C,1), grad band_hz_ (C,1)), all math in float64 on the
    float32 *values* of parameters/buffers.  Formulas: SURVEY §8 a5; they are the chain rule of
    ``synthesize_filters`` written out, including torch's conventions sign(0)=0, clamp passes
    the gradient where input <= max, d|v|/dv = sign(v)."""
    G = np.asarray(G, dtype=np.float64)
    lo_p = low_hz_.detach().double().numpy().reshape(-1, 1)
    bd_p = band_hz_.detach().double().numpy().reshape(-1, 1)
    win = window.detach().double().numpy().reshape(1, -1)
    n = n_.detach().double().numpy().reshape(1, -1)
    sr = float(sample_rate)
    C, K = G.shape
    H = (K - 1) // 2

    low = min_low_hz + np.abs(lo_p)
    pre = low + min_band_hz + np.abs(bd_p)
    unclamped = (pre <= sr / 2.0).astype(np.float64)
    high = np.minimum(pre, sr / 2.0)
    f_lo, f_hi = low / sr, high / sr
    d = n / 2.0 + 1e-8
    if H > 0:
        left = (np.sin(f_hi * n) - np.sin(f_lo * n)) / d
    else:
        left = np.zeros((C, 0))
    centre = 2 * (f_hi - f_lo)
    u = np.concatenate([left, centre, left[:, ::-1]], axis=1) * win
    S = np.abs(u).sum(axis=1, keepdims=True) + 1e-8

    dot = (G * u).sum(axis=1, keepdims=True)
    dL_du = G / S - np.sign(u) * dot / (S * S)
    P = dL_du * win
    Pl = P[:, :H] + P[:, ::-1][:, :H]      # left tap j and its mirror K-1-j
    Pc = P[:, H:H + 1]
    if H > 0:
        d_fhi = (Pl * n * np.cos(f_hi * n) / d).sum(axis=1, keepdims=True) + 2 * Pc
        d_flo = -(Pl * n * np.cos(f_lo * n) / d).sum(axis=1, keepdims=True) - 2 * Pc
    else:
        d_fhi, d_flo = 2 * Pc, -2 * Pc
    d_high = d_fhi / sr
    d_low_direct = d_flo / sr
    g_band = unclamped * d_high * np.sign(bd_p)
    g_low = (d_low_direct + unclamped * d_high) * np.sign(lo_p)
    return g_low, g_band


# --------------------------------------------------------------------------------------
# the module, with the reference's constructor / names (drop-in for tests and CPU baseline)
# --------------------------------------------------------------------------------------
class OracleSincConv1d(torch.nn.Module):
    """CPU restatement of ``agents.perception.SincConv1d`` (perception.py:23-118)."""

    def __init__(self, out_channels, kernel_size, sample_rate=None, min_low_hz=50, min_band_hz=50):
        super().__init__()
        st = erb_filterbank_state(out_channels, kernel_size, sample_rate, min_low_hz, min_band_hz)
        self.out_channels = out_channels
        self.kernel_size = st.kernel_size
        self.sample_rate = st.sample_rate
        self.min_low_hz = min_low_hz
        self.min_band_hz = min_band_hz
        self.low_hz_ = torch.nn.Parameter(st.low_hz_)
        self.band_hz_ = torch.nn.Parameter(st.band_hz_)
        self.register_buffer('window', st.window)
        self.register_buffer('n_', st.n_)

    def filters(self):
        return synthesize_filters(self.low_hz_, self.band_hz_, self.window, self.n_,
                                  self.sample_rate, self.min_low_hz, self.min_band_hz)

    def forward(self, waveform):
        return sinc_forward(waveform, self.filters())


def full_backward(mod: OracleSincConv1d, x, g, dtype):
    """fwd+bwd of the oracle module evaluated in ``dtype`` (fp32 or fp64) with autograd.
    Returns dict(y, gx, gfilt, g_low, g_band) as tensors of that dtype."""
    import copy
    m = copy.deepcopy(mod).to(dtype)
    xr = x.detach().to(dtype).clone().requires_grad_(True)
    w = m.filters()
    w.retain_grad()
    y = sinc_forward(xr, w)
    y.backward(g.detach().to(dtype))
    return dict(y=y.detach(), gx=xr.grad.detach(), gfilt=w.grad.detach(),
                g_low=m.low_hz_.grad.detach(), g_band=m.band_hz_.grad.detach())


def max_rel(a, b):
    """max-norm relative error ||a-b||_inf / ||b||_inf (SURVEY §8c metric)."""
    a = torch.as_tensor(a).double()
    b = torch.as_tensor(b).double()
    den = b.abs().max().item()
    num = (a - b).abs().max().item()
    if den == 0.0:
        return 0.0 if num == 0.0 else float('inf')
    return num / den
