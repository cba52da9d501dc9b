#!/usr/bin/env python
"""CPU mirror of fir_decimate_adjoint_kernel's index arithmetic (branch radix2-wip): per 1024-input
block the staged g window [j_lo, j_lo + g_len) and every shared-memory index a thread forms are
checked for range, and the result is compared with autograd through conv1d."""
import numpy as np
import torch


def emulate(g, fir, factor, L):
    B, T = g.shape
    taps = fir.size
    half = (taps - 1) // 2
    g_len = (1023 + taps - 1) // factor + 3
    dx = np.zeros((B, L))
    for i0 in range(0, L, 1024):
        lo = i0 + half - (taps - 1)
        j_lo = 0 if lo <= 0 else (lo + factor - 1) // factor
        gs = np.zeros((B, g_len))
        for idx in range(g_len):
            j = j_lo + idx
            if j < T:
                gs[:, idx] = g[:, j]
        for i in range(i0, min(i0 + 1024, L)):
            top = i + half
            ja = top - (taps - 1)
            ja = 0 if ja <= 0 else (ja + factor - 1) // factor
            jb = min(top // factor, T - 1)
            acc = np.zeros(B)
            for j in range(ja, jb + 1):
                assert 0 <= j - j_lo < g_len, (i, j, j_lo, g_len)
                assert 0 <= top - factor * j < taps, (i, j, top)
                acc += gs[:, j - j_lo] * fir[top - factor * j]
            dx[:, i] = acc
    return dx


def main():
    rng = np.random.RandomState(0)
    for taps, factor, L in ((256, 2, 3000), (256, 2, 2049), (256, 4, 5000), (255, 3, 1500), (16, 5, 1100),
                            (9, 1, 1030), (256, 1, 1500)):
        half = (taps - 1) // 2
        fir = rng.standard_normal(taps)
        x = torch.from_numpy(rng.standard_normal((2, L))).requires_grad_(True)
        y = torch.nn.functional.conv1d(x[:, None, :], torch.from_numpy(fir)[None, None, :], stride=factor,
                                       padding=half)[:, 0, :]
        g = rng.standard_normal(tuple(y.shape))
        (y * torch.from_numpy(g)).sum().backward()
        got = emulate(g, fir, factor, L)
        err = np.abs(got - x.grad.numpy()).max()
        print(f"taps {taps} factor {factor} L {L}: Ly {y.shape[1]}  max err {err:.1e}")
        assert err < 1e-10


if __name__ == "__main__":
    main()
