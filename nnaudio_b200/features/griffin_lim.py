"""``Griffin_Lim`` — drop-in for ``nnAudio.features.griffin_lim.Griffin_Lim``
(griffin_lim.py:9-148): fast Griffin-Lim phase recovery.

The reference alternates ``torch.istft`` / ``torch.stft`` (cuFFT).  Here every iteration is the
fused inverse kernel (GEMM + overlap-add + window-sum-square) followed by the fused forward
kernel (framing + DFT GEMM, Complex output); the phase update between them is elementwise.

PARITY UNPINNED against the reference *binary*: under torch >= 2.0 the reference module cannot
run at all (``torch.istft`` rejects its real-view input, ``torch.stft`` needs ``return_complex``),
so there are no reference outputs to record.  The oracle (``oracle.griffin_lim``) restates the
reference source line by line with numpy FFTs and the tests pin this module to it with the same
initial phase.
"""
from __future__ import annotations

import types

import numpy as np
import torch
import torch.nn as nn
from scipy.signal import get_window

from .stft import STFT, _inverse_stft


class Griffin_Lim(nn.Module):
    """``forward(S)``: magnitude spectrograms ``(B, n_fft//2+1, T)`` -> waveforms
    ``(B, hop_length * (T - 1))`` (``center=True``).  ``rand_phase`` (optional, same shape as
    ``S``) replaces the internally drawn ``randn`` initial phase for reproducible runs."""

    def __init__(
        self,
        n_fft,
        n_iter=32,
        hop_length=None,
        win_length=None,
        window="hann",
        center=True,
        pad_mode="reflect",
        momentum=0.99,
        device="cpu",
    ):
        super().__init__()
        self.n_fft = n_fft
        self.n_iter = n_iter
        self.center = center
        self.pad_mode = pad_mode
        self.momentum = momentum
        self.device = device
        self.win_length = n_fft if win_length is None else win_length
        self.hop_length = n_fft // 4 if hop_length is None else hop_length
        # kept for attribute parity (griffin_lim.py:85-87); moves with .to() / .cuda()
        self.register_buffer(
            "w", torch.tensor(get_window(window, int(self.win_length), fftbins=True)).float(),
            persistent=False)
        # torch.stft in the reference loop runs with its default center=True (griffin_lim.py:120-127);
        # only the inverse honours ``center``
        self._stft = STFT(n_fft=n_fft, win_length=self.win_length, hop_length=self.hop_length,
                          window=window, center=True, pad_mode=pad_mode, iSTFT=True,
                          output_format="Complex", verbose=False)
        # the reference module owns no state_dict entries: keep the kernels out of checkpoints
        self._stft._non_persistent_buffers_set.update(self._stft._buffers.keys())
        self._inv = types.SimpleNamespace(n_fft=n_fft, stride=self.hop_length, center=center)
        # the reference creates its window on ``device`` (griffin_lim.py:85-87), so
        # Griffin_Lim(n_fft, device='cuda')(S_cuda) works without a .to()
        if device is not None and torch.device(device).type != "cpu":
            self.to(device)

    def _inverse(self, spec):
        st = self._stft
        return _inverse_stft(self._inv, spec, st.kernel_cos_inv, st.kernel_sin_inv, st.window_mask,
                             True, None)

    @torch.no_grad()
    def forward(self, S, rand_phase=None):
        assert (
            S.dim() == 3
        ), "Please make sure your input is in the shape of (batch, freq_bins, timesteps)"
        if rand_phase is None:
            rand_phase = torch.randn(*S.shape, device=S.device)
        angles = torch.stack((torch.cos(2 * np.pi * rand_phase), torch.sin(2 * np.pi * rand_phase)), -1)
        rebuilt = torch.zeros_like(angles)
        S4 = S.unsqueeze(-1)
        decay = self.momentum / (1 + self.momentum)
        for _ in range(self.n_iter):
            tprev = rebuilt
            inverse = self._inverse(S4 * angles)                    # spec -> wav
            rebuilt = self._stft(inverse, output_format="Complex")  # wav -> spec
            angles = rebuilt - decay * tprev
            angles = angles / (torch.sqrt(angles.pow(2).sum(-1)).unsqueeze(-1) + 1e-16)
        return self._inverse(S4 * angles)
