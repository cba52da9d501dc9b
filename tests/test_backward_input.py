"""Gradients w.r.t. the input waveform (the dX half) against the
reference's autograd (x.grad recorded by tests/golden/make_golden.py on CPU)."""
import warnings

import numpy as np
import pytest
import torch

from helpers import build, ref_outputs, rel_errors
from cases import GRAD_CASES, loss_weights, make_input

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", GRAD_CASES, ids=[c[0] for c in GRAD_CASES])
def test_input_gradient_matches_reference_autograd(case):
    cid, cls, ctor, inp, kw = case
    mod = build(cls, ctor).cuda()
    x = torch.from_numpy(make_input(inp)).cuda().requires_grad_(True)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        y = mod(x, **kw)
    w = torch.from_numpy(loss_weights(cid, tuple(y.shape))).cuda()
    (y * w).sum().backward()
    torch.cuda.synchronize()
    want = ref_outputs()["grad|" + cid]
    got = x.grad.cpu().numpy()
    assert got.shape == want.shape
    emax, el2 = rel_errors(got, want)
    tol = 4e-4 if cls == "MFCC" else 1e-4
    assert emax < tol and el2 < tol, (cid, emax, el2)
    # the differentiable path and the fused inference path must agree on the forward value
    with torch.no_grad():
        y0 = mod(x.detach(), **kw)
    e2, _ = rel_errors(y.detach().cpu().numpy(), y0.cpu().numpy())
    assert e2 < 1e-4


def test_gradient_flows_to_upstream_module():
    """A waveform produced by a trainable layer receives gradients through the spectrogram."""
    lin = torch.nn.Linear(64, 4096).cuda()
    spec = build("MelSpectrogram", dict(sr=16000, n_fft=512, hop_length=128, n_mels=40)).cuda()
    z = torch.randn(3, 64, device="cuda")
    loss = spec(lin(z)).log1p().mean()
    loss.backward()
    assert lin.weight.grad is not None and torch.isfinite(lin.weight.grad).all()
    assert lin.weight.grad.abs().max() > 0
