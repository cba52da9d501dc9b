"""EXPERIMENTAL (branch radix2-wip): host-side selection of the decimation-in-time layout."""
import torch

import nnaudio_b200 as nb
from nnaudio_b200.features._common import is_dft_structured


def _pair(mod):
    return mod.wcos.detach()[:, 0, :].contiguous(), mod.wsin.detach()[:, 0, :].contiguous()


def test_structure_check_accepts_plain_windowed_dft_bases_only():
    for kw in (dict(n_fft=512), dict(n_fft=2048, window="hamming"), dict(n_fft=1024, win_length=600),
               dict(n_fft=512, window=("gaussian", 60))):
        assert is_dft_structured(*_pair(nb.STFT(verbose=False, **kw))), kw
    # not a one-sided DFT grid, too small, truncated, or trained away from the structure
    assert not is_dft_structured(*_pair(nb.STFT(n_fft=1024, freq_scale="linear", fmin=50, fmax=6000, sr=22050,
                                               verbose=False)))
    assert not is_dft_structured(*_pair(nb.STFT(n_fft=1024, freq_scale="log", fmin=50, fmax=6000, sr=22050,
                                               verbose=False)))
    assert not is_dft_structured(*_pair(nb.STFT(n_fft=256, verbose=False)))
    wc, ws = _pair(nb.STFT(n_fft=512, verbose=False))
    assert not is_dft_structured(wc[:200], ws[:200])
    wc2 = wc.clone()
    wc2[37, 100] += 1e-3
    assert not is_dft_structured(wc2, ws)
    for kw in (dict(n_fft=512), dict(n_fft=2048, window="hamming")):
        assert is_dft_structured(*_pair(nb.STFT(verbose=False, **kw)), radix=4), kw
    wc4, ws4 = _pair(nb.STFT(n_fft=512, verbose=False))
    ws4 = ws4.clone()
    ws4[200, 33] += 1e-3          # breaks the quarter-period relation (and the mirror)
    assert not is_dft_structured(wc4, ws4, radix=4)
    q = nb.CQT1992v2(sr=22050, fmin=220, n_bins=24, verbose=False)
    assert not is_dft_structured(q.cqt_kernels_real[:, 0, :], q.cqt_kernels_imag[:, 0, :])


def test_layout_request_follows_format_trainable_and_hop(monkeypatch):
    from nnaudio_b200 import _C

    seen = []
    monkeypatch.setenv("NNAUDIO_B200_EXPERIMENTAL", "1")
    monkeypatch.setattr(_C, "_dev_f32", lambda t, name: t)
    monkeypatch.setattr(_C, "pack_basis", lambda a, b, layout=0: seen.append(layout) or torch.zeros(1))
    st = nb.STFT(n_fft=512, hop_length=128, verbose=False)
    st._bases(radix_ok=True)
    st._bases(radix_ok=False)               # Phase output: dense packing, cached separately
    nb.STFT(n_fft=512, hop_length=100, verbose=False)._bases(radix_ok=True)     # hop not a multiple of 128
    nb.STFT(n_fft=512, hop_length=128, trainable=True, verbose=False)._bases(radix_ok=True)
    assert seen == [_C.LAYOUT_RADIX2, _C.LAYOUT_DENSE, _C.LAYOUT_DENSE, _C.LAYOUT_DENSE]
    # radix 4 only on request, when the hop is a multiple of 256 and the quarter-period structure holds
    seen.clear()
    monkeypatch.setenv("NNAUDIO_B200_RADIX", "4")
    nb.STFT(n_fft=1024, hop_length=256, verbose=False)._bases(radix_ok=True)
    nb.STFT(n_fft=1024, hop_length=128, verbose=False)._bases(radix_ok=True)
    assert seen == [_C.LAYOUT_RADIX4, _C.LAYOUT_RADIX2]
    monkeypatch.delenv("NNAUDIO_B200_RADIX")
    monkeypatch.setenv("NNAUDIO_B200_EXPERIMENTAL", "0")
    seen.clear()
    nb.STFT(n_fft=512, hop_length=128, verbose=False)._bases(radix_ok=True)
    assert seen == [_C.LAYOUT_DENSE]
