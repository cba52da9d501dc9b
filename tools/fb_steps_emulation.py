#!/usr/bin/env python
"""CPU replay of the static action list of the banded-filterbank epilogue (fb_steps_kernel +
the FMT 5 path of framed_tcb_kernel): for every bin the two running sums flush / accumulate exactly
as the kernel does, over arbitrary bin ranges (tile x warp parts); the result must equal fb @ P, and
a filter no wider than a range must receive at most two partial sums."""
import sys, os
import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from nnaudio_b200 import design  # noqa: E402


def build_steps(fb):
    n_fb, F = fb.shape
    steps, cur_a, cur_b = [], -1, -1
    for k in range(F + 160):
        wa = wb = 0.0
        fa = fb_ = -1
        if k < F:
            nz = np.nonzero(fb[:, k])[0]
            assert len(nz) <= 2
            used_a = used_b = False
            for j in nz:
                if j == cur_a: wa, used_a = fb[j, k], True
                elif j == cur_b: wb, used_b = fb[j, k], True
            for j in nz:
                if j in (cur_a, cur_b): continue
                if not used_a:
                    if cur_a >= 0: fa = cur_a
                    cur_a, wa, used_a = j, fb[j, k], True
                else:
                    if cur_b >= 0: fb_ = cur_b
                    cur_b, wb, used_b = j, fb[j, k], True
        steps.append((wa, wb, fa, fb_, cur_a, cur_b))
    return steps


def run_ranges(steps, P, F, n_fb, ranges):
    out = np.zeros(n_fb)
    partials = np.zeros(n_fb, dtype=int)
    for lo, hi in ranges:
        ma = mb = 0.0
        ca = cb = -1
        for k in range(lo, hi):
            if k >= F: break
            wa, wb, fa, fb_, cur_a, cur_b = steps[k]
            if fa >= 0:
                out[fa] += ma; partials[fa] += ma != 0; ma = 0.0
            if fb_ >= 0:
                out[fb_] += mb; partials[fb_] += mb != 0; mb = 0.0
            ma += wa * P[k]; mb += wb * P[k]
            ca, cb = cur_a, cur_b
        if ca >= 0: out[ca] += ma; partials[ca] += ma != 0
        if cb >= 0: out[cb] += mb; partials[cb] += mb != 0
    return out, partials


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    for sr, n_fft, n_mels, htk, norm in ((22050, 2048, 128, False, 1), (16000, 2048, 128, False, 1),
                                          (22050, 512, 40, True, None), (44100, 1024, 64, False, 1)):
        fb = np.asarray(design.mel_filterbank(sr, n_fft, n_mels=n_mels, htk=htk, norm=norm), dtype=np.float64)
        n_fb, F = fb.shape
        steps = build_steps(fb)
        width = max((np.ptp(np.nonzero(fb[j])[0]) + 1) if fb[j].any() else 0 for j in range(n_fb))
        P = rng.random(F)
        want = fb @ P
        for nb in (96, 112, 128, 88):
            # ranges as the kernel cuts them: tiles of nb-2 outputs, two chunk-aligned parts per tile
            ranges, k0 = [], 0
            while k0 < F:
                mid = k0 + 8 * ((nb // 8) // 2) - 2
                ranges += [(k0, mid), (mid, k0 + nb - 2)]
                k0 += nb - 2
            got, partials = run_ranges(steps, P, F, n_fb, ranges)
            err = np.abs(got - want).max() / np.abs(want).max()
            minlen = min(hi - lo for lo, hi in ranges)
            print(f"sr {sr} n_fft {n_fft} mels {n_mels}: width {width} nb {nb} min range {minlen} "
                  f"err {err:.1e} max partials {partials.max()}")
            assert err < 1e-12
            if minlen >= width:
                assert partials.max() <= 2
