"""Init-time caches of packed bases / tables (host logic, no GPU): rebuilt when the source tensor
changes (load_state_dict, optimiser step), one entry per device, and safe to share between module
replicas on different devices (torch.nn.DataParallel copies __dict__ shallowly)."""
import threading

import numpy as np
import torch

import nnaudio_b200 as nb
from nnaudio_b200 import _C
from nnaudio_b200.features import _common
from nnaudio_b200.features.cqt import _ScaleCache
from nnaudio_b200.features.stft import _InverseAdjoint, _InverseBasis


def test_per_device_cache_semantics():
    c = _common.PerDeviceCache()
    built = []

    def make(tag):
        def build():
            built.append(tag)
            return tag
        return build

    assert c.lookup("cuda:0", 1, make("a0")) == "a0"
    assert c.lookup("cuda:1", 1, make("a1")) == "a1"        # another device: its own entry
    assert c.lookup("cuda:0", 1, make("x")) == "a0"          # still cached, not evicted by cuda:1
    assert c.lookup("cuda:0", 2, make("b0")) == "b0"          # key changed: rebuilt
    assert c.lookup("cuda:1", 1, make("y")) == "a1"
    assert built == ["a0", "a1", "b0"]


def test_lookup_returns_the_callers_own_value_under_threads():
    """Two replicas hammering one shared cache from two threads never see each other's value."""
    c = _common.PerDeviceCache()
    bad = []

    def worker(dev):
        for i in range(2000):
            v = c.lookup(dev, i % 3, lambda: (dev, i % 3))
            if v != (dev, i % 3):
                bad.append((dev, i, v))

    ts = [threading.Thread(target=worker, args=(d,)) for d in ("cuda:0", "cuda:1")]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert not bad


def test_every_cache_rebuilds_on_version_change_and_keeps_devices_apart(monkeypatch):
    calls = []

    def fake(name):
        def f(*args, **kw):
            calls.append(name)
            return (name, len(calls))
        return f

    for name in ("pack_basis", "pack_adjoint_basis", "build_filterbank_table", "pack_fir", "pack_istft_basis"):
        monkeypatch.setattr(_C, name, fake(name))
    w = torch.zeros(4, 8)
    w_meta = torch.zeros(4, 8, device="meta")
    for cache, args, args_meta in (
            (_common.PackedBasis(), (w, w), (w_meta, w_meta)),
            (_common.AdjointBasis(), (w, w), (w_meta, w_meta)),
            (_common.FilterbankTable(), (w,), (w_meta,)),
            (_common.PackedFir(), (w, 2), (w_meta, 2)),
            (_InverseBasis(), (w, w, 5, True), (w_meta, w_meta, 5, True))):
        calls.clear()
        a = cache.get(*args)
        assert cache.get(*args) is a and len(calls) == 1          # cached
        m = cache.get(*args_meta)
        assert m is not a and len(calls) == 2                      # second device: own entry
        assert cache.get(*args) is a and len(calls) == 2           # first device not evicted
        w.add_(1.0)                                                # in-place update bumps _version
        b = cache.get(*args)
        assert b is not a and len(calls) == 3                      # rebuilt after the change
        assert cache.get(*args_meta) is m and len(calls) == 3


def test_scale_and_inverse_adjoint_caches(monkeypatch):
    monkeypatch.setattr(_C, "pack_basis", lambda a, b: ("packed", a.shape))
    sc = _ScaleCache()
    lengths = torch.tensor([4.0, 9.0, 16.0])
    s1 = sc.get(lengths, 0.5)
    assert torch.equal(s1, torch.tensor([1.0, 1.5, 2.0])) and sc.get(lengths, 0.5) is s1
    assert torch.equal(sc.get(lengths, 1.0), torch.tensor([2.0, 3.0, 4.0]))
    lengths.mul_(4.0)
    assert torch.equal(sc.get(lengths, 1.0), torch.tensor([4.0, 6.0, 8.0]))

    n = 8
    kc = torch.randn(n, n)
    ks = torch.randn(n, n)
    win = torch.rand(n)
    adj = _InverseAdjoint()
    w_re, w_im, packed = adj.get(kc, ks, win, True)
    assert w_re.shape == (n // 2 + 1, n) and packed[0] == "packed"
    full_re = (kc * (win / n)[:, None]).t()
    assert torch.allclose(w_re[1], full_re[1] + full_re[n - 1]) and torch.allclose(w_re[0], full_re[0])
    assert adj.get(kc, ks, win, True)[0] is w_re
    assert adj.get(kc, ks, win, False)[0].shape == (n, n)


def test_tap_support_cache_follows_the_buffers():
    q = nb.CQT1992v2(sr=22050, fmin=220, n_bins=12, verbose=False)
    b0, e0 = q._tap_support()
    assert q._tap_support()[0] is b0 and (e0 > b0).all()
    lens = (e0 - b0).astype(np.int64)
    assert (np.diff(lens) <= 0).all(), "wavelets shorten with frequency"
    with torch.no_grad():
        q.cqt_kernels_real.zero_()
        q.cqt_kernels_imag.zero_()
    b1, e1 = q._tap_support()
    assert (b1 == 0).all() and (e1 == 0).all()


def test_modules_deepcopy_and_pickle_like_plain_nn_modules():
    """Users copy / torch.save whole modules (the reference's are plain nn.Modules): the host-side
    caches must not get in the way, and the copy must own independent buffers."""
    import copy
    import io
    import pickle

    mods = [nb.STFT(n_fft=256, iSTFT=True, verbose=False), nb.MelSpectrogram(sr=16000, n_fft=256, n_mels=20, verbose=False),
            nb.MFCC(sr=16000, n_fft=256, n_mels=20, n_mfcc=8, verbose=False),
            nb.Gammatonegram(sr=16000, n_fft=256, n_bins=8, verbose=False),
            nb.CQT1992v2(sr=22050, fmin=220, n_bins=12, verbose=False), nb.CQT2010v2(sr=22050, n_bins=24, fmin=220, verbose=False),
            nb.VQT(sr=22050, n_bins=24, fmin=220, gamma=3, verbose=False), nb.iSTFT(n_fft=256, verbose=False),
            nb.CQT2010(sr=22050, n_bins=24, fmin=220, verbose=False), nb.Griffin_Lim(n_fft=256, n_iter=2)]
    for m in mods:
        c = copy.deepcopy(m)
        p = pickle.loads(pickle.dumps(m))
        buf = io.BytesIO()
        torch.save(m, buf)
        for other in (c, p):
            assert type(other) is type(m)
            sd, so = m.state_dict(), other.state_dict()
            assert sd.keys() == so.keys()
            for k in sd:
                assert torch.equal(sd[k], so[k]) and (sd[k].numel() == 0 or sd[k].data_ptr() != so[k].data_ptr())
