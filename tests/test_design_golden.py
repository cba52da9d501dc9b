"""L2 (init-time design) parity: every buffer our modules register must be
bit-identical to the reference's (sha256 recorded by tests/golden/make_golden.py
from the unmodified reference) — the drop-in contract of SURVEY.md §8(b1)."""
import hashlib

import numpy as np
import pytest

from helpers import CASES, build, ref_buffers
from cases import DESIGN_CASES


def _sha(t):
    return hashlib.sha256(np.ascontiguousarray(t.detach().cpu().numpy()).tobytes()).hexdigest()


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_buffers_bit_identical_to_reference(case):
    cid, cls, ctor, _, _ = case
    mod = build(cls, ctor)
    ours = {k: v for k, v in mod.state_dict().items() if v is not None}
    want = ref_buffers()[cid]
    assert sorted(ours) == sorted(want), "state_dict keys differ from the reference"
    for k, (shape, digest) in want.items():
        assert list(ours[k].shape) == shape, k
        assert _sha(ours[k]) == digest, f"{cid}: buffer {k} is not bit-identical"


@pytest.mark.parametrize("case", DESIGN_CASES, ids=[c[0] for c in DESIGN_CASES])
def test_constructor_sweep_buffers_bit_identical_to_reference(case):
    """Wider constructor sweep (windows, frequency scales, htk / norm, bins per octave, gamma ...):
    same keys, shapes and bytes as the reference's state_dict."""
    cid, cls, ctor = case
    mod = build(cls, ctor)
    ours = {k: v for k, v in mod.state_dict().items() if v is not None}
    want = ref_buffers()[cid]
    assert sorted(ours) == sorted(want), "state_dict keys differ from the reference"
    for k, (shape, digest) in want.items():
        assert list(ours[k].shape) == shape, k
        assert _sha(ours[k]) == digest, f"{cid}: buffer {k} is not bit-identical"
