"""Test infrastructure only: float64 torch stand-ins for the handful of C entry points the
*differentiable* host compositions call (pyramid octave loop, iSTFT adjoint).  Patched over
``nnaudio_b200._C`` they let the CPU suite check the host-side autograd wiring — which stages
are chained, which padding each one uses, how octave gradients are summed, how the one-sided
mirror is folded — against the reference's autograd goldens without a GPU.  The kernels
themselves are checked on the GPU by tests/test_backward.py; nothing here ships.
"""
from __future__ import annotations

import numpy as np
import torch

from nnaudio_b200 import _C


def _pad(x, pad, pad_mode):
    if pad == 0:
        return x
    mode = "reflect" if pad_mode == _C.PAD_REFLECT else "constant"
    return torch.nn.functional.pad(x[:, None, :], (pad, pad), mode=mode)[:, 0, :]


def _framed(x, w_re, w_im, hop, center, pad_mode):
    """(B, L), (F, K) x2 -> (B, F, T, 2) = (conv(x, w_re), -conv(x, w_im)) in float64."""
    K = w_re.shape[1]
    xp = _pad(x.double(), K // 2 if center else 0, pad_mode)[:, None, :]
    re = torch.nn.functional.conv1d(xp, w_re.double()[:, None, :], stride=hop)
    im = -torch.nn.functional.conv1d(xp, w_im.double()[:, None, :], stride=hop)
    return torch.stack((re, im), -1)


def cqt1992v2_forward(x, k_real, k_imag, packed, k_begin, k_end, hop, center, pad_mode, scale,
                      scale_all, out_format, sqrt_eps, path=None):
    assert out_format == _C.FMT_COMPLEX and scale is None and scale_all == 1.0
    return _framed(x, k_real, k_imag, hop, center, pad_mode).float()


def stft_forward(x, wcos, wsin, packed, n_fft, hop, center, pad_mode, out_format, sqrt_eps, path=None):
    assert out_format == _C.FMT_COMPLEX
    return _framed(x, wcos.reshape(wcos.shape[0], -1), wsin.reshape(wsin.shape[0], -1), hop, center,
                   pad_mode).float()


def framed_backward_input(g, packed_adj, K, hop, center, pad_mode, L_in):
    w_re, w_im = packed_adj
    x = torch.zeros((g.shape[0], L_in), dtype=torch.float64, requires_grad=True)
    with torch.enable_grad():
        y = _framed(x, w_re, w_im, hop, center, pad_mode)
        (dx,) = torch.autograd.grad(y, x, g.double())
    return dx.float()


def framed_backward_weight(g, x, K, hop, center, pad_mode):
    F = g.shape[1]
    w_re = torch.zeros((F, K), dtype=torch.float64, requires_grad=True)
    w_im = torch.zeros((F, K), dtype=torch.float64, requires_grad=True)
    with torch.enable_grad():
        y = _framed(x, w_re, w_im, hop, center, pad_mode)
        d_re, d_im = torch.autograd.grad(y, (w_re, w_im), g.double())
    return d_re.float(), d_im.float()


def istft_forward(X, packed, window, n_fft, hop, center, length):
    from oracle import nnaudio_oracle as oracle

    kc, ks, onesided = packed
    y = oracle.istft(X.numpy(), kc.numpy(), ks.numpy(), window.numpy().reshape(1, -1, 1), hop,
                     center=center, onesided=onesided, length=length)
    return torch.from_numpy(np.ascontiguousarray(y)).float()


def fir_decimate(x, fir, factor):
    taps = fir.numel()
    return torch.nn.functional.conv1d(x[:, None, :].double(), fir.reshape(1, 1, -1).double(), stride=factor,
                                      padding=(taps - 1) // 2)[:, 0, :].float()


def fir_decimate_adjoint(g, fir, factor, L_in):
    x = torch.zeros((g.shape[0], L_in), dtype=torch.float64, requires_grad=True)
    with torch.enable_grad():
        taps = fir.numel()
        y = torch.nn.functional.conv1d(x[:, None, :], fir.reshape(1, 1, -1).double(), stride=factor,
                                       padding=(taps - 1) // 2)[:, 0, :]
        (dx,) = torch.autograd.grad(y, x, g.double())
    return dx.float()


def install(monkeypatch):
    """Route the calls of the differentiable host paths to the stand-ins above."""
    monkeypatch.setattr(_C, "_dev_f32", lambda t, name: t)
    monkeypatch.setattr(_C, "pack_basis", lambda w_re, w_im: torch.zeros(1))
    monkeypatch.setattr(_C, "pack_adjoint_basis", lambda w_re, w_im: (w_re.clone(), w_im.clone()))
    monkeypatch.setattr(_C, "pack_istft_basis",
                        lambda kc, ks, f_in, onesided: (kc.clone(), ks.clone(), bool(onesided)))
    monkeypatch.setattr(_C, "cqt1992v2_forward", cqt1992v2_forward)
    monkeypatch.setattr(_C, "stft_forward", stft_forward)
    monkeypatch.setattr(_C, "framed_backward_input", framed_backward_input)
    monkeypatch.setattr(_C, "framed_backward_weight", framed_backward_weight)
    monkeypatch.setattr(_C, "istft_forward", istft_forward)
    monkeypatch.setattr(_C, "fir_decimate", fir_decimate, raising=False)
    monkeypatch.setattr(_C, "fir_decimate_adjoint", fir_decimate_adjoint, raising=False)
