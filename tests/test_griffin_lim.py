"""Griffin_Lim.  The reference module cannot run under torch >= 2.0, so
the pin is the oracle's restatement of its source (oracle.griffin_lim) with the same initial phase;
on top of that, properties any Griffin-Lim must have (spectral convergence, shape)."""
import numpy as np
import pytest
import torch

from helpers import oracle, rel_errors  # noqa: F401 (sets sys.path)
import cpu_kernels

import nnaudio_b200 as nb

CFG = [
    dict(n_fft=256, n_iter=6, hop_length=64),
    dict(n_fft=512, n_iter=4, hop_length=128, win_length=400, window="hamming", momentum=0.5),
    dict(n_fft=256, n_iter=3, hop_length=64, pad_mode="constant"),
]


def _problem(cfg, seed, B=2, T=40):
    rng = np.random.RandomState(seed)
    x = rng.standard_normal((B, cfg["hop_length"] * (T - 1))).astype(np.float32)
    st = nb.STFT(n_fft=cfg["n_fft"], hop_length=cfg["hop_length"], win_length=cfg.get("win_length"),
                 window=cfg.get("window", "hann"), output_format="Magnitude", verbose=False)
    S = oracle.stft(x, st.wsin.numpy(), st.wcos.numpy(), st.stride, True, "reflect", "Magnitude",
                    False, None, np.float64).astype(np.float32)
    ph = rng.standard_normal(S.shape).astype(np.float32)
    return S, ph


def _oracle(cfg, S, ph):
    kw = dict(cfg)
    n_fft = kw.pop("n_fft")
    kw["hop"] = kw.pop("hop_length")
    return oracle.griffin_lim(S, ph, n_fft, **kw)


@pytest.mark.parametrize("cfg", CFG, ids=[f"cfg{i}" for i in range(len(CFG))])
def test_host_loop_matches_oracle(cfg, monkeypatch):
    """CPU: the iteration wiring (which transform is centred, momentum rule, normalisation)."""
    cpu_kernels.install(monkeypatch)
    S, ph = _problem(cfg, 7)
    y = nb.Griffin_Lim(**cfg)(torch.from_numpy(S), rand_phase=torch.from_numpy(ph)).numpy()
    want = _oracle(cfg, S, ph)
    assert y.shape == want.shape == (S.shape[0], cfg["hop_length"] * (S.shape[2] - 1))
    emax, el2 = rel_errors(y, want)
    assert emax < 1e-4 and el2 < 1e-4, (emax, el2)


@pytest.mark.parametrize("n_fft,hop,win_length,window,pad_mode", [
    (256, 64, 256, "hann", "reflect"), (512, 128, 400, "hamming", "reflect"),
    (256, 64, 256, "hann", "constant"), (256, 64, 256, "hamming", "reflect")])
def test_oracle_transform_pair_is_pinned_to_torch(n_fft, hop, win_length, window, pad_mode):
    """The reference's loop calls torch.stft / torch.istft; the oracle's restatements of those two
    calls are held to torch itself (the modern complex API of the same functions)."""
    rng = np.random.RandomState(3)
    y = rng.standard_normal((2, hop * 30)).astype(np.float64)
    w = oracle._padded_window(window, win_length, n_fft, np.float64)
    w_t = torch.from_numpy(w[(n_fft - win_length) // 2: (n_fft - win_length) // 2 + win_length].copy())
    X_t = torch.stft(torch.from_numpy(y), n_fft, hop, win_length=win_length, window=w_t, center=True,
                     pad_mode=pad_mode, return_complex=True)
    X_o = oracle.torch_stft_restated(y, n_fft, hop, w, pad_mode)
    assert X_o.shape == tuple(X_t.shape)
    assert np.abs(X_o - X_t.numpy()).max() < 1e-10 * np.abs(X_o).max()
    for center in (True, False):
        if not center and (window == "hann" or win_length < n_fft):
            continue  # torch rejects a window sum-square that touches zero at the uncut edges
        y_t = torch.istft(X_t, n_fft, hop, win_length=win_length, window=w_t, center=center)
        y_o = oracle.torch_istft_restated(X_t.numpy(), n_fft, hop, w, center)
        assert y_o.shape == tuple(y_t.shape)
        assert np.abs(y_o - y_t.numpy()).max() < 1e-10 * np.abs(y_o).max()


def test_module_surface():
    g = nb.Griffin_Lim(n_fft=512)
    assert g.hop_length == 128 and g.win_length == 512 and g.n_iter == 32 and g.momentum == 0.99
    assert tuple(g.w.shape) == (512,)
    assert len(g.state_dict()) == 0, "the reference module has no state_dict entries"
    with pytest.raises(AssertionError, match="batch, freq_bins, timesteps"):
        g(torch.zeros(257, 10))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        g(torch.rand(1, 257, 10))


def _spectral_convergence(cfg, S, y):
    """|| |STFT(y)| - S || / ||S|| with the oracle's STFT (float64)."""
    st = nb.STFT(n_fft=cfg["n_fft"], hop_length=cfg["hop_length"], win_length=cfg.get("win_length"),
                 window=cfg.get("window", "hann"), output_format="Magnitude", verbose=False)
    S2 = oracle.stft(np.asarray(y, dtype=np.float32), st.wsin.numpy(), st.wcos.numpy(), st.stride,
                     True, "reflect", "Magnitude", False, None, np.float64)
    return float(np.linalg.norm(S2 - S) / np.linalg.norm(S))


@pytest.mark.gpu
@pytest.mark.parametrize("cfg", CFG, ids=[f"cfg{i}" for i in range(len(CFG))])
def test_cuda_matches_oracle(cfg):
    """The iteration renormalises the phase (angles / |angles|), which amplifies any rounding
    where |angles| ~ 0 and feeds it to the next iteration: waveforms of two correct
    implementations drift apart by ~1e-2 within a few iterations (measured by injecting 3e-5
    noise into the float64 stand-ins).  So: a loose bound on the waveform, a tight one on what the
    algorithm optimises (spectral convergence), and an exact one on the shape."""
    S, ph = _problem(cfg, 7)
    mod = nb.Griffin_Lim(**cfg).cuda()
    y = mod(torch.from_numpy(S).cuda(), rand_phase=torch.from_numpy(ph).cuda()).cpu().numpy()
    want = _oracle(cfg, S, ph)
    assert y.shape == want.shape and np.isfinite(y).all()
    _, el2 = rel_errors(y, want)
    assert el2 < 5e-2, el2
    sc_gpu, sc_ref = _spectral_convergence(cfg, S, y), _spectral_convergence(cfg, S, want)
    assert abs(sc_gpu - sc_ref) < 0.02 * max(sc_ref, 1e-3) + 2e-3, (sc_gpu, sc_ref)


@pytest.mark.gpu
def test_cuda_spectral_convergence():
    """32 iterations on a real signal's magnitude: the rebuilt waveform's magnitude spectrogram
    approaches the target (spectral convergence far below the random-phase start)."""
    torch.manual_seed(0)
    t = torch.arange(16000, device="cuda") / 16000.0
    x = (torch.sin(2 * np.pi * 440 * t) + 0.5 * torch.sin(2 * np.pi * 1320 * t))[None]
    st = nb.STFT(n_fft=512, hop_length=128, output_format="Magnitude", verbose=False).cuda()
    with torch.no_grad():
        S = st(x)
    def sc(n_iter):
        y = nb.Griffin_Lim(n_fft=512, hop_length=128, n_iter=n_iter).cuda()(S)
        with torch.no_grad():
            S2 = st(y)
        T = min(S.shape[-1], S2.shape[-1])
        return (torch.linalg.norm(S2[..., :T] - S[..., :T]) / torch.linalg.norm(S[..., :T])).item()

    start, end = sc(0), sc(32)   # ~0.65 -> ~0.12 with float64 transforms
    assert end < 0.25 and end < 0.4 * start, (start, end)


@pytest.mark.gpu
def test_device_argument_places_the_module():
    """the reference builds its window on ``device`` (griffin_lim.py:85-87), so
    ``Griffin_Lim(n_fft, device='cuda')(S_cuda)`` works without ``.to()``."""
    cfg = CFG[0]
    S, ph = _problem(cfg, 3)
    mod = nb.Griffin_Lim(**cfg, device="cuda")
    assert mod.w.is_cuda and mod._stft.wsin.is_cuda
    y = mod(torch.from_numpy(S).cuda(), rand_phase=torch.from_numpy(ph).cuda())
    assert y.is_cuda and torch.isfinite(y).all()
