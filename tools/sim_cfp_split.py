#!/usr/bin/env python
"""Predict on the CPU what the split-bf16 tensor-core contractions do to the CFP outputs: the host layer
of nnaudio_b200.features.cfp runs with tests/cpu_kernels.py stand-ins whose contraction emulates the
three-term bf16 split (tools/sim_split_bf16.py), and every output is compared with the unmodified
reference's fixture (tests/golden/ref_cfp.npz).  Used to set the tolerances of tests/test_zz_gpu_cfp.py
before spending GPU time.  Test infrastructure only.

    python tools/sim_cfp_split.py [--no-demean]
"""
import os
import sys
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden"), os.path.join(ROOT, "tools")):
    sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from helpers import GOLDEN, build, rel_errors  # noqa: E402
import cpu_kernels  # noqa: E402
from cases import CFP_CASES, make_input  # noqa: E402
from sim_split_bf16 import _Patch, framed_split_bf16  # noqa: E402
from nnaudio_b200 import _C  # noqa: E402


def split_forward(x, k_real, k_imag, packed, kb, ke, hop, center, pad_mode, scale, scale_all, fmt, eps,
                  path=None):
    c = framed_split_bf16(x, k_real, k_imag, hop, center, pad_mode).float().double()  # fp32 accumulator
    return cpu_kernels._format(cpu_kernels._scaled(c, scale, scale_all), fmt, eps)


def main():
    cpu_kernels.install(_Patch())
    _C.cqt1992v2_forward = split_forward
    ref = dict(np.load(os.path.join(GOLDEN, "ref_cfp.npz")))
    for cid, cls, ctor, inp in CFP_CASES:
        mod = build(cls, ctor)
        with torch.no_grad(), warnings.catch_warnings():
            warnings.simplefilter("ignore")
            y = mod(torch.from_numpy(make_input(inp)))
        ys = y if isinstance(y, tuple) else (y,)
        print(cid, ["%.1e/%.1e" % rel_errors(t.numpy(), ref[f"{cid}|{i}"]) for i, t in enumerate(ys)])


if __name__ == "__main__":
    main()
