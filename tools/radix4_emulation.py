#!/usr/bin/env python
"""CPU emulation of the EXPERIMENTAL radix-4 kernel's layout and butterflies (branch radix2-wip):
mirrors pack_basis_radix_kernel<4> and epilogue_tile_radix4 (same row / column / slot formulas)
against the oracle STFT.  Segment r (sample phase n = 4 m + r) holds N/8 bins as
[re half | negated-im half] per 64-column tile; the (k = 0, im) slot of segment r carries the real
number that determines the sub-DFT's Nyquist bin U_r[N/8] = e^{-i pi r / 4} s_r:
    r = 0: re(U)   (im = 0)        r = 1: re(U)   (im = -re)
    r = 2: im(U)   (re = 0)        r = 3: re(U)   (im =  re)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import nnaudio_oracle as oracle  # noqa: E402
import nnaudio_b200 as nb  # noqa: E402

BNS = 64  # columns per segment and tile: 32 bins x (re, im)


def pack(w_re, w_im, K):
    rows_seg, nyq, half = K // 4, K // 8, BNS // 2
    packed = np.zeros((4, rows_seg, K // 4))
    r_idx = np.arange(rows_seg)
    tile, within = np.divmod(r_idx, BNS)
    part, j = np.divmod(within, half)
    k = tile * half + j
    for seg in range(4):
        cols = np.arange(seg, K, 4)
        rows = np.where(part[:, None] == 0, w_re[k][:, cols], -w_im[k][:, cols])
        slot = (part == 1) & (k == 0)
        rows[slot] = -w_im[nyq, cols] if seg == 2 else w_re[nyq, cols]
        packed[seg] = rows
    return packed


def mul_mi(z):  # -i * z
    return z.imag - 1j * z.real


def mul_pi(z):  # +i * z
    return -z.imag + 1j * z.real


def emulate(x, w_re, w_im, K, hop, pad, pad_mode, T):
    F = K // 2 + 1
    NH, NQ = K // 2, K // 4
    xp = oracle.pad_signal(x.astype(np.float64), pad, pad_mode)
    packed = pack(w_re, w_im, K)
    hop4, k4, half = hop // 4, K // 4, BNS // 2
    B = x.shape[0]
    out = np.zeros((B, F, T), dtype=np.complex128)
    for b in range(B):
        acc = []
        for seg in range(4):
            plane = xp[b, seg::4]
            frames = np.stack([plane[t * hop4: t * hop4 + k4] for t in range(T)])
            acc.append(frames @ packed[seg].T)                        # (T, rows_seg)
        for n_tile in range((K // 4) // BNS):
            cols = slice(n_tile * BNS, (n_tile + 1) * BNS)
            for c in range(half):
                k = n_tile * half + c
                re = [acc[s][:, cols][:, c] for s in range(4)]
                im = [acc[s][:, cols][:, half + c] for s in range(4)]
                if k == 0:
                    a = re                                             # U_r[0] are real
                    out[b, 0] = a[0] + a[1] + a[2] + a[3]
                    out[b, NQ] = (a[0] - a[2]) + 1j * (a[3] - a[1])
                    out[b, NH] = a[0] - a[1] + a[2] - a[3]
                    s = im                                             # the packed Nyquist numbers
                    v = [s[0] + 0j, s[1] - 1j * s[1], 0 + 1j * s[2], s[3] + 1j * s[3]]
                    out[b, K // 8] = v[0] + v[1] + v[2] + v[3]
                    out[b, NQ + K // 8] = v[0] + mul_mi(v[1]) - v[2] + mul_pi(v[3])
                else:
                    u = [re[s] + 1j * im[s] for s in range(4)]
                    out[b, k] = u[0] + u[1] + u[2] + u[3]
                    out[b, NQ + k] = u[0] + mul_mi(u[1]) - u[2] + mul_pi(u[3])
                    out[b, NQ - k] = np.conj(u[0]) + mul_mi(np.conj(u[1])) - np.conj(u[2]) + mul_pi(np.conj(u[3]))
                    out[b, NH - k] = np.conj(u[0] - u[1] + u[2] - u[3])
    return out


def main():
    rng = np.random.RandomState(0)
    for n_fft, hop, window, pad_mode in ((512, 256, "hann", "reflect"), (2048, 512, "hann", "reflect"),
                                         (1024, 256, "hamming", "constant")):
        st = nb.STFT(n_fft=n_fft, hop_length=hop, window=window, pad_mode=pad_mode, verbose=False)
        wr, wi = st.wcos.numpy()[:, 0].astype(np.float64), st.wsin.numpy()[:, 0].astype(np.float64)
        x = rng.standard_normal((2, hop * 12)).astype(np.float32)
        want = oracle.stft(x, st.wsin.numpy(), st.wcos.numpy(), hop, True, pad_mode, "Complex", False, None,
                           np.float64)
        want = want[..., 0] + 1j * want[..., 1]
        got = emulate(x, wr, wi, n_fft, hop, n_fft // 2, pad_mode, want.shape[-1])
        err = np.abs(got - want).max() / np.abs(want).max()
        print(f"n_fft {n_fft} hop {hop} {window}/{pad_mode}: radix-4 layout vs oracle STFT  {err:.2e}")
        assert err < 1e-6, err


if __name__ == "__main__":
    main()
