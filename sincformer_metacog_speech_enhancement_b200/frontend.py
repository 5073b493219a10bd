"""Next row N1: the first three lines of ``PerceptionAgent.forward`` as one fused front-end.

    reference, agents/perception.py:232-234        x = self.sinc_conv(waveform)
                                                   x = self.sinc_norm(x)          # nn.GroupNorm(8, 64), :157
                                                   x = F.gelu(x)

``SincFrontEnd`` computes the same thing on B200 with the GroupNorm statistics taken from the convolution epilogue
and ONE pass for normalise + affine + GELU (read y, write z); backward is two streaming passes in front of the
SincConv backward.  ``FusedGroupNormGELU`` is the norm + activation alone (same ``weight`` / ``bias`` names as
``nn.GroupNorm``, so ``sinc_norm``'s state_dict loads).  fp32 activations, CUDA only, no fallback.
"""
from __future__ import annotations

import ctypes

import torch
import torch.nn as nn

from . import _lib
from . import sinc as _sinc
from .sinc import SincConv1d, SincConvFunction, _bank, _check_input, _options, _ptr, _small


def _stream(dev):
    return ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


class _GroupNormGELU(torch.autograd.Function):
    """z = gelu(group_norm(y, G, weight, bias, eps)).  ``partial`` may carry the per-tile sums emitted by the
    convolution epilogue (rows_per_group = 1); otherwise one read pass computes them."""

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, y, weight, bias, G, eps, partial):
        lib = _lib.load()
        if not y.is_cuda or y.dtype != torch.float32 or y.dim() != 3:
            raise RuntimeError("FusedGroupNormGELU (B200): expects a float32 CUDA tensor (batch, channels, samples)")
        y = y.contiguous()
        B, C, T = y.shape
        if C % G:
            raise RuntimeError(f"FusedGroupNormGELU: need channels % groups == 0, got C={C} G={G}")
        dev = y.device
        w = _small(weight, "weight", dev)
        b = _small(bias, "bias", dev)
        z = torch.empty_like(y)
        stats = torch.empty((B, G, 2), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            if partial is None:
                npart, rows = lib.sinc_gn_chunks(T), C // G
                partial = torch.empty((B * C, npart, 2), dtype=torch.float32, device=dev)
                _lib.check(lib.sinc_gn_stats(_ptr(y), B, T, C, _ptr(partial), _stream(dev)), "sinc_gn_stats")
            else:
                npart, rows = partial.shape[-2], 1
            _lib.check(lib.sinc_gn_gelu_forward(_ptr(y), _ptr(partial), npart, rows, _ptr(w), _ptr(b), _ptr(z), _ptr(stats),
                                                B, T, C, G, float(eps), _stream(dev)), "sinc_gn_gelu_forward")
        ctx.save_for_backward(y, stats, w, b)
        ctx.G = G
        return z

    @staticmethod
    @torch.autograd.function.once_differentiable
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, dz):
        lib = _lib.load()
        y, stats, w, b = ctx.saved_tensors
        B, C, T = y.shape
        dev = y.device
        dz = dz.float().contiguous()
        dy = torch.empty_like(y)
        need_p = ctx.needs_input_grad[1] or ctx.needs_input_grad[2]
        dwb = torch.empty((2, C), dtype=torch.float32, device=dev) if need_p else None
        scratch = torch.empty((B * C, lib.sinc_gn_chunks(T), 2), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(lib.sinc_gn_gelu_backward(_ptr(dz), _ptr(y), _ptr(stats), _ptr(w), _ptr(b), _ptr(dy),
                                                 _ptr(dwb[0] if need_p else None), _ptr(dwb[1] if need_p else None),
                                                 _ptr(scratch), B, T, C, ctx.G, _stream(dev)), "sinc_gn_gelu_backward")
        return (dy, dwb[0] if ctx.needs_input_grad[1] else None, dwb[1] if ctx.needs_input_grad[2] else None,
                None, None, None)


class FusedGroupNormGELU(nn.Module):
    """``F.gelu(nn.GroupNorm(num_groups, num_channels, eps)(x))`` in one pass over x (plus one statistics pass when the
    producer did not supply them).  Parameter names follow ``nn.GroupNorm``."""

    def __init__(self, num_groups, num_channels, eps=1e-5):
        super().__init__()
        if num_channels % num_groups:
            raise ValueError("num_channels must be divisible by num_groups")
        self.num_groups, self.num_channels, self.eps = num_groups, num_channels, eps
        self.weight = nn.Parameter(torch.ones(num_channels))
        self.bias = nn.Parameter(torch.zeros(num_channels))

    @classmethod
    def from_reference(cls, gn: nn.GroupNorm):
        m = cls(gn.num_groups, gn.num_channels, gn.eps)
        m.load_state_dict(gn.state_dict())
        return m.to(gn.weight.device)

    def forward(self, x, partial=None):
        return _GroupNormGELU.apply(x, self.weight, self.bias, self.num_groups, self.eps, partial)

    def extra_repr(self):
        return f"{self.num_groups}, {self.num_channels}, eps={self.eps}"


class _SincConvWithStats(torch.autograd.Function):
    """SincConv forward that also returns the GroupNorm partial sums from its epilogue; backward = SincConv backward."""

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, x, low_hz_, band_hz_, window, n_, sample_rate, min_low_hz, min_band_hz, engine=None, debug=None):
        lib = _lib.load()
        _check_input(x)           # the same device / shape / dtype / empty-input rules as SincConv1d.forward
        dev = x.device
        low, band = _small(low_hz_, "low_hz_", dev), _small(band_hz_, "band_hz_", dev)
        win, nn_ = _small(window, "window", dev), _small(n_, "n_", dev)
        xd = x.detach()
        if xd.stride(2) != 1 and xd.shape[2] > 1:
            xd = xd.contiguous()
        B, _, T = xd.shape
        C, K = low.numel(), win.numel()
        y = torch.empty((B, C, T), dtype=torch.float32, device=dev)
        partial = torch.empty((B, C // 8, lib.sinc_forward_gn_partials(T), 2), dtype=torch.float32, device=dev)
        fb = _bank(low, band, win, nn_, sample_rate, min_low_hz, min_band_hz)
        opts, _ = _options(engine)
        with torch.cuda.device(dev):
            rc = lib.sinc_forward_gn(ctypes.byref(fb), _ptr(xd), xd.stride(0) if B > 1 else T, _ptr(y), B, T,
                                     _ptr(partial), _stream(dev), ctypes.byref(opts))
        _lib.check(rc, "sinc_forward_gn")
        ctx.save_for_backward(xd, low, band, win, nn_)
        ctx.attrs = (float(sample_rate), float(min_low_hz), float(min_band_hz))
        ctx.sync = None
        ctx.engine = engine
        ctx.debug = debug
        ctx.in_dtypes = (x.dtype, low_hz_.dtype, band_hz_.dtype)
        ctx.mark_non_differentiable(partial)
        return y, partial

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_y, _grad_partial):
        out = SincConvFunction.backward(ctx, grad_y)
        return out[:10]


class SincFrontEnd(nn.Module):
    """``gelu(sinc_norm(sinc_conv(waveform)))`` -- lines :232-234 of ``PerceptionAgent.forward`` -- with the submodule
    names of the reference (``sinc_conv``, ``sinc_norm``), so the corresponding slice of its state_dict loads."""

    def __init__(self, out_channels=64, kernel_size=251, sample_rate=None, num_groups=8, eps=1e-5):
        super().__init__()
        self.sinc_conv = SincConv1d(out_channels, kernel_size, sample_rate)
        self.sinc_norm = FusedGroupNormGELU(num_groups, out_channels, eps)

    @classmethod
    def from_reference(cls, perception_agent):
        sc, gn = perception_agent.sinc_conv, perception_agent.sinc_norm
        m = cls(sc.out_channels, sc.kernel_size, sc.sample_rate, gn.num_groups, gn.eps)
        m.sinc_conv.load_state_dict({k: v.detach().float() for k, v in sc.state_dict().items()})
        m.sinc_norm.load_state_dict(gn.state_dict())
        return m.to(gn.weight.device)

    def forward(self, waveform):
        if waveform.dim() == 2:
            waveform = waveform.unsqueeze(1)           # (batch, samples) as in PerceptionAgent.forward :228-229
        sc, gn = self.sinc_conv, self.sinc_norm
        C, T = sc.out_channels, waveform.shape[-1]
        engine = sc.engine or _sinc._default_engine
        fused = (gn.num_channels // gn.num_groups == 8 and C % 8 == 0 and sc.kernel_size <= 265
                 and waveform.dtype == torch.float32 and sc.low_hz_.dtype == torch.float32 and engine in ("auto", "tc"))
        if fused:
            y, partial = _SincConvWithStats.apply(waveform, sc.low_hz_, sc.band_hz_, sc.window, sc.n_, sc.sample_rate,
                                                  sc.min_low_hz, sc.min_band_hz, sc.engine, sc._debug)
            return gn(y, partial)
        return gn(sc(waveform))
