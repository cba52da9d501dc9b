"""Host-side gates of the block-partial STFT kernel (no GPU): which (n_fft, hop) pairs the library
accepts, and `is_hann_dft` -- the check on the module's OWN buffers that decides whether the
analytically generated block basis may stand in for them (anything else keeps the dense kernel)."""
import numpy as np
import pytest
import torch

import nnaudio_b200.features as nb
from nnaudio_b200 import _C
from nnaudio_b200.features._common import is_hann_dft


def _mat(buf):
    return buf.detach().reshape(buf.shape[0], buf.shape[-1])


@pytest.mark.parametrize("n_fft,hop,ok", [
    (2048, 512, True), (2048, 1024, True), (512, 256, True), (512, 128, True), (256, 64, True),
    (2048, 256, False),   # R = 8: no epilogue for it
    (2048, 500, False), (400, 160, False), (512, 96, False), (128, 32, False),  # hop % 64 != 0
])
def test_block_layout_shapes(n_fft, hop, ok):
    assert _C.block_layout_ok(n_fft, hop) is ok
    assert (_C.lib().nnab_packed_block_bytes(n_fft, hop) > 0) is ok


def test_reference_default_basis_is_hann_dft():
    st = nb.STFT(n_fft=512, hop_length=128, verbose=False)
    assert is_hann_dft(_mat(st.wcos), _mat(st.wsin))
    mel = nb.MelSpectrogram(sr=16000, n_fft=1024, hop_length=256, verbose=False)
    assert is_hann_dft(_mat(mel.stft.wcos), _mat(mel.stft.wsin))


@pytest.mark.parametrize("ctor", [
    dict(window="hamming"), dict(window="blackman"), dict(win_length=400), dict(freq_scale="linear", fmin=50, fmax=4000),
    dict(freq_scale="log", fmin=50, fmax=4000), dict(freq_bins=200),
], ids=lambda d: next(iter(d)) + "=" + str(next(iter(d.values()))))
def test_other_bases_keep_the_dense_kernel(ctor):
    st = nb.STFT(n_fft=512, hop_length=128, sr=16000, verbose=False, **ctor)
    wc, ws = _mat(st.wcos), _mat(st.wsin)
    if st.freq_bins is not None and st.freq_bins < wc.shape[0]:
        wc, ws = wc[: st.freq_bins], ws[: st.freq_bins]
    assert not is_hann_dft(wc, ws)


def test_perturbed_or_loaded_buffers_are_detected():
    st = nb.STFT(n_fft=256, hop_length=64, verbose=False)
    wc, ws = _mat(st.wcos).clone(), _mat(st.wsin).clone()
    assert is_hann_dft(wc, ws)
    wc2 = wc.clone(); wc2[37, 101] += 5e-5          # one element of a "trained" basis
    assert not is_hann_dft(wc2, ws)
    assert not is_hann_dft(wc, -ws)                  # sign convention matters (re - i im)
    assert not is_hann_dft(wc[:, :128], ws[:, :128])  # wrong shape


def test_block_basis_request_needs_forward_only_module(monkeypatch):
    """Trainable modules never ask for the block layout (their bases change under the optimiser)."""
    asked = []
    monkeypatch.setattr(_C, "pack_basis_block", lambda w, hop: asked.append(("block", hop)) or torch.zeros(1))
    monkeypatch.setattr(_C, "pack_basis", lambda a, b, layout=0: asked.append(("dense", layout)) or torch.zeros(1))
    monkeypatch.setattr(_C, "_dev_f32", lambda t, name: t)
    st = nb.STFT(n_fft=512, hop_length=128, verbose=False)
    st._bases(block_ok=True)
    tr = nb.STFT(n_fft=512, hop_length=128, trainable=True, verbose=False)
    tr._bases(block_ok=True)
    assert asked == [("block", 128), ("dense", 0)]
    np.testing.assert_equal(len(asked), 2)
