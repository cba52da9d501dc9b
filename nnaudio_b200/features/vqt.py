"""``VQT`` — drop-in for ``nnAudio.features.vqt.VQT`` (vqt.py:9-215): the
CQT2010v2 pyramid with one wavelet bank per octave
(``cqt_kernels_real_{i}`` / ``cqt_kernels_imag_{i}``) and the ``gamma``
bandwidth offset.  ``gamma=0`` is bit-identical to :class:`CQT2010v2`
(tests/test_vqt.py:30-41 in the reference)."""
from __future__ import annotations

from time import time

import numpy as np
import torch
import torch.nn as nn

from .. import design
from ._common import PackedBasis, as_matrix, broadcast_dim
from .cqt import _ScaleCache, _check_format_and_norm, _pyramid_forward, _v2_normalization


class VQT(nn.Module):
    def __init__(
        self,
        sr=22050,
        hop_length=512,
        fmin=32.70,
        fmax=None,
        n_bins=84,
        filter_scale=1,
        bins_per_octave=12,
        norm=True,
        basis_norm=1,
        gamma=0,
        window="hann",
        pad_mode="reflect",
        earlydownsample=True,
        trainable=False,
        output_format="Magnitude",
        verbose=True,
    ):
        super().__init__()
        self.norm = norm
        self.hop_length = hop_length
        self.pad_mode = pad_mode
        self.n_bins = n_bins
        self.earlydownsample = earlydownsample
        self.trainable = trainable
        self.output_format = output_format
        self.filter_scale = filter_scale
        self.bins_per_octave = bins_per_octave
        self.sr = sr
        self.gamma = gamma
        self.basis_norm = basis_norm

        Q = float(filter_scale) / (2 ** (1 / bins_per_octave) - 1)

        if verbose:
            print("Creating low pass filter ...", end="\r")
        start = time()
        lowpass = torch.tensor(design.lowpass_fir(0.50, 256, 0.001))
        self.register_buffer("lowpass_filter", lowpass[None, None, :])
        if verbose:
            print("Low pass filter created, time used = {:.4f} seconds".format(time() - start))

        n_filters = min(bins_per_octave, n_bins)
        self.n_filters = n_filters
        self.n_octaves = int(np.ceil(float(n_bins) / bins_per_octave))
        if verbose:
            print("num_octave = ", self.n_octaves)

        self.fmin_t = fmin * 2 ** (self.n_octaves - 1)
        remainder = n_bins % bins_per_octave
        if remainder == 0:
            fmax_t = self.fmin_t * 2 ** ((bins_per_octave - 1) / bins_per_octave)
        else:
            fmax_t = self.fmin_t * 2 ** ((remainder - 1) / bins_per_octave)
        self.fmin_t = fmax_t / 2 ** (1 - 1 / bins_per_octave)
        if fmax_t > sr / 2:
            raise ValueError(
                "The top bin {}Hz has exceeded the Nyquist frequency, \
                            please reduce the n_bins".format(
                    fmax_t
                )
            )

        if self.earlydownsample:
            if verbose:
                print("Creating early downsampling filter ...", end="\r")
            start = time()
            sr, self.hop_length, self.downsample_factor, early_fir = design.early_downsample_plan(
                sr, hop_length, fmax_t, Q, self.n_octaves
            )
            self.earlydownsample = early_fir is not None
            self.register_buffer(
                "early_downsample_filter",
                torch.tensor(early_fir)[None, None, :] if early_fir is not None else None,
            )
            if verbose:
                print("Early downsampling filter created, \
                        time used = {:.4f} seconds".format(time() - start))
        else:
            self.downsample_factor = 1.0

        # final normalisation lengths use the post-early-downsample sr (vqt.py:107-117)
        alpha = 2.0 ** (1.0 / bins_per_octave) - 1.0
        freqs = fmin * 2.0 ** (np.r_[0:n_bins] / np.double(bins_per_octave))
        self.frequencies = freqs
        lenghts = np.ceil(Q * sr / (freqs + gamma / alpha))
        self.n_fft = int(2 ** (np.ceil(np.log2(int(max(lenghts))))))
        self.register_buffer("lenghts", torch.tensor(lenghts).float())

        # one bank per octave, built from the ORIGINAL sr (reference quirk, vqt.py:120-134)
        my_sr = self.sr
        for i in range(self.n_octaves):
            if i > 0:
                my_sr /= 2
            Q = float(self.filter_scale) / (2 ** (1 / self.bins_per_octave) - 1)
            basis, self.n_fft, _, _ = design.cqt_bank(
                Q,
                my_sr,
                self.fmin_t * 2 ** -i,
                self.n_filters,
                self.bins_per_octave,
                norm=self.basis_norm,
                topbin_check=False,
                gamma=self.gamma,
            )
            self.register_buffer(
                "cqt_kernels_real_{}".format(i),
                torch.tensor(basis.real.astype(np.float32)).unsqueeze(1),
            )
            self.register_buffer(
                "cqt_kernels_imag_{}".format(i),
                torch.tensor(basis.imag.astype(np.float32)).unsqueeze(1),
            )
        self._scale = _ScaleCache()
        self._packed = [PackedBasis() for _ in range(self.n_octaves)]

    def _banks(self):
        re = [as_matrix(getattr(self, f"cqt_kernels_real_{i}")) for i in range(self.n_octaves)]
        im = [as_matrix(getattr(self, f"cqt_kernels_imag_{i}")) for i in range(self.n_octaves)]
        packed = [self._packed[i].get(re[i], im[i]) for i in range(self.n_octaves)]
        return re, im, packed

    def _bank_tensors(self):
        return [(getattr(self, f"cqt_kernels_real_{i}"), getattr(self, f"cqt_kernels_imag_{i}"))
                for i in range(self.n_octaves)]

    def forward(self, x, output_format=None, normalization_type="librosa"):
        output_format = output_format or self.output_format
        _check_format_and_norm(output_format, normalization_type)
        x = broadcast_dim(x)
        return _pyramid_forward(self, x, output_format,
                                _v2_normalization(self, normalization_type, output_format))
