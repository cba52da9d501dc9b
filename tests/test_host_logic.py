"""Host-side behaviour of the nn.Module boundary that needs no GPU: the
reference's exception types (SURVEY.md §8 b1 'Conventions'), buffer names,
attribute surface, octave planning, and the loud no-CPU-fallback failure."""
import warnings

import numpy as np
import pytest
import torch

import nnaudio_b200 as nb
from nnaudio_b200.features.cqt import _octave_plan
from nnaudio_b200.features._common import broadcast_dim, tap_support


def test_broadcast_dim_rules():
    assert broadcast_dim(torch.zeros(7)).shape == (1, 7)
    assert broadcast_dim(torch.zeros(2, 7)).shape == (2, 7)
    assert broadcast_dim(torch.zeros(2, 1, 7)).shape == (2, 7)
    with pytest.raises(ValueError, match="Only support input with shape"):
        broadcast_dim(torch.zeros(1, 1, 2, 7))


def test_cpu_tensor_fails_loudly_no_fallback():
    m = nb.STFT(n_fft=256, verbose=False)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.randn(1, 4000))


def test_reference_exceptions_raised_before_the_c_call():
    m = nb.STFT(n_fft=512, verbose=False)
    with pytest.raises(AssertionError, match="shorter than reflect padding"):
        m(torch.randn(1, 100))
    with pytest.raises(ValueError):
        m(torch.randn(1, 4000), output_format="nope")
    q = nb.CQT1992v2(sr=22050, fmin=220, n_bins=12, verbose=False)
    with pytest.raises(ValueError, match="normalization_type"):
        q(torch.randn(1, 4000), normalization_type="bogus")
    with pytest.raises(RuntimeError, match="Padding size"):
        q(torch.randn(1, 100))
    with pytest.raises(ValueError, match="Nyquist"):
        nb.CQT1992v2(sr=8000, fmin=220, n_bins=84, verbose=False)
    with pytest.raises(ValueError, match="Nyquist"):
        nb.CQT2010v2(sr=8000, n_bins=96, verbose=False)


def test_forward_only_guard():
    m = nb.STFT(n_fft=256, trainable=True, verbose=False)
    assert isinstance(m.wsin, torch.nn.Parameter) and m.wsin.requires_grad
    # the training path exists (tests/test_backward.py, GPU); on CPU it still fails loudly
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.randn(1, 4000))
    # the pyramid trains too (octave loop over the same kernels); CPU tensors are still refused
    qt = nb.CQT2010v2(trainable=True, verbose=False)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        qt(torch.randn(1, 40000))
    # trainable inverse kernels / window have no dW path: refuse instead of silently not training
    inv = nb.iSTFT(n_fft=256, trainable_kernels=True, verbose=False)
    with pytest.raises(NotImplementedError, match="forward-only"):
        inv(torch.zeros(1, 256, 10, 2))


def test_attribute_surface_matches_reference():
    s = nb.STFT(n_fft=512, hop_length=128, verbose=False)
    for a in ("stride", "n_fft", "pad_amount", "freq_bins", "bins2freq", "bin_list", "win_length",
              "center", "pad_mode", "trainable", "output_format"):
        assert hasattr(s, a), a
    assert "n_fft=512" in s.extra_repr()
    mel = nb.MelSpectrogram(verbose=False)
    assert isinstance(mel.stft, nb.STFT) and mel.power == 2.0 and mel.stride == 512
    mf = nb.MFCC(sr=16000, verbose=False)
    assert mf.n_mfcc == 20 and mf.m_mfcc == 20 and mf.top_db == 80.0
    assert "_dct_rows" not in mf.state_dict()
    c = nb.CQT2010v2(n_bins=88, verbose=False)
    assert c.n_octaves == 8 and c.n_fft == 256 and c.earlydownsample is False
    assert issubclass(nb.CQT, nb.CQT1992v2)
    v = nb.VQT(gamma=3, verbose=False)
    assert v.n_filters == 12 and v.gamma == 3


def test_octave_plan_lengths_and_fallback():
    # cfg4 level lengths of SURVEY.md §8 row a14
    T, flags = _octave_plan(661500, 512, [256] * 8, "reflect")
    assert T == 1292 and not any(flags)
    T, flags = _octave_plan(6000, 512, [256] * 7, "reflect")
    assert T == 12 and flags == [False] * 6 + [True]
    with pytest.raises(RuntimeError):
        _octave_plan(6001, 500, [256] * 7, "reflect")  # octave frame counts diverge


def test_tap_support_of_cqt_bank():
    q = nb.CQT1992v2(sr=22050, fmin=220, n_bins=24, verbose=False)
    kb, ke = q._tap_support()
    lens = q.lenghts.numpy().astype(int)
    assert ((ke - kb) <= lens).all() and ((ke - kb) >= lens - 2).all()
    b, e = tap_support(np.array([[0, 0, 1, 2, 0], [0, 0, 0, 0, 0]]))
    assert (b.tolist(), e.tolist()) == ([2, 0], [4, 0])


def test_legacy_spectrogram_alias_warns():
    import importlib, sys
    sys.modules.pop("nnaudio_b200.Spectrogram", None)
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        mod = importlib.import_module("nnaudio_b200.Spectrogram")
    assert mod.STFT is nb.STFT and any("deprecated" in str(x.message) for x in w)


def test_nnaudio_import_shim():
    """`from nnAudio import features` resolves to this engine when <repo>/shim is on sys.path."""
    import importlib, os, sys
    from helpers import ROOT
    sys.path.insert(0, os.path.join(ROOT, "shim"))
    for k in [k for k in sys.modules if k == "nnAudio" or k.startswith("nnAudio.")]:
        del sys.modules[k]
    try:
        feats = importlib.import_module("nnAudio.features")
        assert feats.MelSpectrogram is nb.MelSpectrogram and feats.CQT1992v2 is nb.CQT1992v2
        # the reference's per-file import paths (`from nnAudio.features.stft import STFT`, ...)
        for sub, names in (("stft", ["STFT", "iSTFT"]), ("mel", ["MelSpectrogram", "MFCC"]),
                           ("gammatone", ["Gammatonegram"]), ("vqt", ["VQT"]), ("griffin_lim", ["Griffin_Lim"]),
                           ("cfp", ["CFP", "Combined_Frequency_Periodicity"]),
                           ("cqt", ["CQT1992v2", "CQT2010v2", "CQT", "CQT1992", "CQT2010"])):
            m = importlib.import_module("nnAudio.features." + sub)
            for n in names:
                assert getattr(m, n) is getattr(nb.features, n), (sub, n)
    finally:
        sys.path.remove(os.path.join(ROOT, "shim"))
        for k in [k for k in sys.modules if k == "nnAudio" or k.startswith("nnAudio.")]:
            del sys.modules[k]


def test_batches_beyond_the_per_call_limit_are_chunked(monkeypatch):
    """More than _C.MAX_BATCH clips: several C calls on the same stream, concatenated."""
    from nnaudio_b200 import _C

    monkeypatch.setattr(_C, "MAX_BATCH", 3)
    calls = []

    @_C._batch_chunked
    def fake_forward(x, gain):
        calls.append(x.shape[0])
        return x * gain

    x = torch.arange(16.0).reshape(8, 2)
    assert torch.equal(fake_forward(x, 2.0), x * 2.0) and calls == [3, 3, 2]
    for name in ("stft_forward", "stft_filterbank_forward", "mfcc_forward", "cqt1992v2_forward",
                 "cqt_pyramid_forward"):
        assert hasattr(getattr(_C, name), "__wrapped__"), name


def _ref_errors():
    import json
    import os
    from helpers import GOLDEN
    with open(os.path.join(GOLDEN, "ref_errors.json")) as f:
        return json.load(f)


from cases import ERROR_CASES  # noqa: E402


@pytest.mark.parametrize("case", ERROR_CASES, ids=[c[0] for c in ERROR_CASES])
def test_same_exception_type_as_the_reference(case):
    """Malformed constructor arguments / inputs raise the exception TYPE the reference raises
    (recorded from the unmodified reference in tests/golden/ref_errors.json), and they do so before
    the C call — so the check also holds for CPU tensors."""
    from helpers import build
    cid, cls, ctor, call = case
    want = _ref_errors()[cid]
    assert want != "ok"
    exc = {"AssertionError": AssertionError, "ValueError": ValueError, "RuntimeError": RuntimeError,
           "NameError": NameError}[want]
    with pytest.raises(exc) as info, warnings.catch_warnings():
        warnings.simplefilter("ignore")
        mod = build(cls, ctor)
        if call[0] == "forward":
            mod(torch.zeros(call[1]), **call[2])
        elif call[0] == "inverse":
            mod.inverse(torch.zeros(call[1]), **call[2])
    assert "no CPU fallback" not in str(info.value), "the reference-type error must come first"


def _ref_attributes():
    import json
    import os
    from helpers import GOLDEN
    with open(os.path.join(GOLDEN, "ref_attributes.json")) as f:
        return json.load(f)


from cases import CASES as _CASES, DESIGN_CASES, attribute_surface  # noqa: E402

_ATTR_CASES = [(c[0], c[1], c[2]) for c in _CASES] + list(DESIGN_CASES)


@pytest.mark.parametrize("case", _ATTR_CASES, ids=[c[0] for c in _ATTR_CASES])
def test_public_attribute_surface_matches_reference(case):
    """Every public non-tensor attribute of the reference module (n_fft, stride, frequencies,
    kernel_width, downsample_factor, ...) exists on ours with the same value (fixture recorded from
    the unmodified reference); ours may carry more."""
    from helpers import build
    cid, cls, ctor = case
    ours = attribute_surface(build(cls, ctor))
    for name, want in _ref_attributes()[cid].items():
        assert name in ours, f"{cls}.{name} missing"
        got = ours[name]
        if isinstance(want, float) and isinstance(got, (int, float)):
            assert got == pytest.approx(want, rel=1e-9, abs=1e-12), name
        elif isinstance(want, list) and want and want[0] == "ndarray":
            assert got[:3] == want[:3] and got[3] == pytest.approx(want[3], rel=1e-9), name
        else:
            assert got == want, (name, got, want)


def test_signatures_match_the_reference():
    """Constructor / forward / inverse parameter names, order and defaults of every class, against
    the signatures recorded from the unmodified reference (tests/golden/ref_signatures.json).  The
    only extension: Griffin_Lim.forward's optional trailing ``rand_phase``."""
    import inspect
    import json
    import os
    from helpers import GOLDEN

    with open(os.path.join(GOLDEN, "ref_signatures.json")) as f:
        ref = json.load(f)
    extra = {("Griffin_Lim", "forward"): [["rand_phase", "None"]]}
    for cls, methods in ref.items():
        klass = getattr(nb.features, cls)
        for meth, want in methods.items():
            params = list(inspect.signature(getattr(klass, meth)).parameters.values())[1:]
            got = [[q.name, None if q.default is inspect.Parameter.empty else repr(q.default)] for q in params]
            assert got == want + extra.get((cls, meth), []), (cls, meth, got, want)
