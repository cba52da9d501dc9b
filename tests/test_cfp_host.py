"""Combined_Frequency_Periodicity / CFP on the CPU: init-time design bit-identical to the unmodified
reference (fixtures: tests/golden/make_golden_cfp.py), attribute / signature / exception-type parity, the
oracle restatement against the reference's outputs, and the whole host layer of OUR modules — the
half-vector cosine-transform formulation, the mirrored cut-off weights, the mean removal, the frame crop
and the three contraction calls — with float64 stand-ins for the C wrappers (tests/cpu_kernels.py)
against the same outputs.  The kernels behind the calls are checked by tests/test_zz_gpu_cfp.py."""
import hashlib
import inspect
import json
import os
import warnings

import numpy as np
import pytest
import torch

from helpers import GOLDEN, build, oracle, rel_errors
from cases import CFP_CASES, CFP_DESIGN_CASES, CFP_ERROR_CASES, attribute_surface, make_input
import cpu_kernels

import nnaudio_b200 as nb


def _meta():
    with open(os.path.join(GOLDEN, "ref_cfp.json")) as f:
        return json.load(f)


def _outputs():
    return dict(np.load(os.path.join(GOLDEN, "ref_cfp.npz")))


def _sha(t):
    return hashlib.sha256(np.ascontiguousarray(t.detach().cpu().numpy()).tobytes()).hexdigest()


ALL_CTORS = [(c[0], c[1], c[2]) for c in CFP_CASES] + list(CFP_DESIGN_CASES)


@pytest.mark.parametrize("case", ALL_CTORS, ids=[c[0] for c in ALL_CTORS])
def test_buffers_bit_identical_to_reference(case):
    cid, cls, ctor = case
    mod = build(cls, ctor)
    want = _meta()["buffers"][cid]
    got = {k: [list(v.shape), _sha(v)] for k, v in mod.state_dict().items()}
    assert got == want


@pytest.mark.parametrize("case", CFP_DESIGN_CASES, ids=[c[0] for c in CFP_DESIGN_CASES])
def test_attribute_surface_matches_reference(case):
    cid, cls, ctor = case
    assert attribute_surface(build(cls, ctor)) == _meta()["attributes"][cid]


@pytest.mark.parametrize("cls", ["Combined_Frequency_Periodicity", "CFP"])
def test_signatures_match_reference(cls):
    klass = getattr(nb.features, cls)
    for meth, want in _meta()["signatures"][cls].items():
        got = [[q.name, None if q.default is inspect.Parameter.empty else repr(q.default)]
               for q in list(inspect.signature(getattr(klass, meth)).parameters.values())[1:]]
        assert got == want, (cls, meth)


def _run_oracle(mod, x, drop, dtype=np.float64):
    return oracle.cfp(x, mod.h.numpy(), mod.freq2logfreq_matrix.numpy(), mod.quef2logfreq_matrix.numpy(),
                      mod.N, mod.hop_length, mod.g, mod.tc_idx, mod.fc_idx, mod.HighFreqIdx, mod.HighQuefIdx,
                      drop_edge_frames=drop, dtype=dtype)


@pytest.mark.parametrize("case", CFP_CASES, ids=[c[0] for c in CFP_CASES])
def test_oracle_reproduces_reference(case):
    cid, cls, ctor, inp = case
    mod = build(cls, ctor)
    outs = _run_oracle(mod, make_input(inp), cls == "Combined_Frequency_Periodicity")
    ref = _outputs()
    n = 4 if cls == "Combined_Frequency_Periodicity" else 1
    for i in range(n):
        emax, el2 = rel_errors(outs[i], ref[f"{cid}|{i}"])
        assert emax < 2e-5 and el2 < 2e-5, (cid, i, emax, el2)


@pytest.mark.parametrize("case", CFP_CASES, ids=[c[0] for c in CFP_CASES])
def test_host_layer_reproduces_reference(case, monkeypatch):
    cpu_kernels.install(monkeypatch)
    cid, cls, ctor, inp = case
    mod = build(cls, ctor)
    with torch.no_grad(), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        y = mod(torch.from_numpy(make_input(inp)))
    ys = y if isinstance(y, tuple) else (y,)
    ref = _outputs()
    for i, t in enumerate(ys):
        want = ref[f"{cid}|{i}"]
        assert tuple(t.shape) == want.shape, (cid, i)
        emax, el2 = rel_errors(t.numpy(), want)
        assert emax < 2e-5 and el2 < 2e-5, (cid, i, emax, el2)
    # the attribute the reference sets in forward
    assert attribute_surface(mod) == _meta()["attributes"][cid]


@pytest.mark.parametrize("case", CFP_ERROR_CASES, ids=[c[0] for c in CFP_ERROR_CASES])
def test_malformed_use_raises_the_reference_exception_type(case, monkeypatch):
    cpu_kernels.install(monkeypatch)
    cid, cls, ctor, shape = case
    want = _meta()["errors"][cid]
    assert want != "ok"
    with pytest.raises(Exception) as ei:
        mod = build(cls, ctor)
        with torch.no_grad():
            mod(torch.zeros(shape))
    assert type(ei.value).__name__ == want, (cid, type(ei.value).__name__, want)


def test_stft_geometry_matches_torch_stft_alignment():
    """Frame t of torch.stft(n_fft=N, win_length=W, center=True) reads x[t hop - d + m]; the framed
    kernel with K taps centred by K // 2 and the window shifted j taps must read the same samples."""
    from nnaudio_b200.features.cfp import _stft_geometry

    for N, W in [(8000, 2049), (4000, 1025), (5333, 1500), (11025, 2049), (8000, 8000), (8000, 7999), (64, 3)]:
        left, K, j = _stft_geometry(N, W)
        d = N // 2 - left
        assert K % 64 == 0 and j >= 0 and j + W <= K
        assert K // 2 - j == d           # first window tap of frame t sits at x[t hop - d]


def test_forward_refuses_gradients_and_cpu_tensors():
    mod = build("CFP", {})
    with pytest.raises(NotImplementedError):
        mod(torch.zeros(1, 4000, requires_grad=True))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        with torch.no_grad():
            mod(torch.zeros(1, 4000))
