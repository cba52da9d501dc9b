#!/usr/bin/env python
"""CPU emulation of the EXPERIMENTAL radix-2 kernel's index math (branch radix2-wip), mirroring
pad_split_radix2_kernel, pack_basis_radix2_kernel and epilogue_tile_radix2 in tc_kernels.cu line
by line (same row / column / plane formulas), against the oracle STFT of the same module.
It cannot check the tcgen05 mechanics — only that the data layout and the butterfly are right.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import nnaudio_oracle as oracle  # noqa: E402
import nnaudio_b200 as nb  # noqa: E402

R2_BN = int(__import__('os').environ.get('NNAB_RADIX_BN', '128'))


def pack_basis_radix2(w_re, w_im, K):
    """-> packed[seg][row][k2] (fp32; the hi/lo split is orthogonal to the layout)."""
    rows_seg, kpad2, half, nyq = K // 2, K // 2, R2_BN // 2, K // 4
    packed = np.zeros((2, rows_seg, kpad2), dtype=np.float64)
    for seg in range(2):
        for r in range(rows_seg):
            tile, within = divmod(r, R2_BN)
            part, j = divmod(within, half)
            k = tile * half + j
            for k2 in range(kpad2):
                n = 2 * k2 + seg
                v = 0.0
                if k < nyq and n < K:
                    if part == 0:
                        v = w_re[k, n]
                    elif k != 0:
                        v = -w_im[k, n]
                    else:
                        v = w_re[nyq, n] if seg == 0 else -w_im[nyq, n]
                packed[seg, r, k2] = v
    return packed


def pack_basis_radix2_fast(w_re, w_im, K):
    """Vectorised equivalent of the loop above (checked against it at a small size)."""
    rows_seg, half, nyq = K // 2, R2_BN // 2, K // 4
    packed = np.zeros((2, rows_seg, K // 2), dtype=np.float64)
    r = np.arange(rows_seg)
    tile, within = np.divmod(r, R2_BN)
    part, j = np.divmod(within, half)
    k = tile * half + j
    for seg in range(2):
        cols = np.arange(seg, K, 2)
        rows = np.where(part[:, None] == 0, w_re[k][:, cols], -w_im[k][:, cols])
        slot = (part == 1) & (k == 0)
        rows[slot] = w_re[nyq, cols] if seg == 0 else -w_im[nyq, cols]
        packed[seg] = rows
    return packed


def planes_radix2(xp):
    """plane_r[i] = xpad[2 i + r]"""
    return xp[..., 0::2], xp[..., 1::2]


def emulate(x, w_re, w_im, K, hop, pad, pad_mode, T, fast=True):
    F = K // 2 + 1
    xp = oracle.pad_signal(x.astype(np.float64), pad, pad_mode)
    planes = planes_radix2(xp)
    packed = (pack_basis_radix2_fast if fast else pack_basis_radix2)(w_re, w_im, K)
    hop2, k2, half = hop // 2, K // 2, R2_BN // 2
    n_tiles = (K // 2) // R2_BN
    B = x.shape[0]
    out = np.zeros((B, F, T), dtype=np.complex128)
    NH = F - 1
    for b in range(B):
        frames = [np.stack([planes[seg][b, t * hop2: t * hop2 + k2] for t in range(T)]) for seg in range(2)]
        acc = [frames[seg] @ packed[seg].T for seg in range(2)]      # (T, rows_seg) per segment
        for n_tile in range(n_tiles):
            cols = slice(n_tile * R2_BN, (n_tile + 1) * R2_BN)
            s0, u = acc[0][:, cols], acc[1][:, cols]                # TMEM: [S0 re|S0 im|U re|U im]
            for c in range(half):
                k = n_tile * half + c
                ar, ai, br, bi = s0[:, c], s0[:, half + c], u[:, c], u[:, half + c]
                if k == 0:
                    out[b, 0] = ar + br
                    out[b, NH] = ar - br
                    out[b, NH // 2] = ai + 1j * bi
                else:
                    out[b, k] = (ar + br) + 1j * (ai + bi)
                    out[b, NH - k] = (ar - br) + 1j * (bi - ai)
    return out


def main():
    rng = np.random.RandomState(0)
    # the element-wise pack loop vs its vectorised form
    st = nb.STFT(n_fft=512, hop_length=128, verbose=False)
    wr, wi = st.wcos.numpy()[:, 0].astype(np.float64), st.wsin.numpy()[:, 0].astype(np.float64)
    assert np.array_equal(pack_basis_radix2(wr, wi, 512), pack_basis_radix2_fast(wr, wi, 512))
    for n_fft, hop, window, pad_mode in ((512, 128, "hann", "reflect"), (2048, 512, "hann", "reflect"),
                                         (1024, 256, "hamming", "constant")):
        st = nb.STFT(n_fft=n_fft, hop_length=hop, window=window, pad_mode=pad_mode, verbose=False)
        wr, wi = st.wcos.numpy()[:, 0].astype(np.float64), st.wsin.numpy()[:, 0].astype(np.float64)
        x = rng.standard_normal((2, hop * 20)).astype(np.float32)
        want = oracle.stft(x, st.wsin.numpy(), st.wcos.numpy(), hop, True, pad_mode, "Complex", False,
                           None, np.float64)
        want = want[..., 0] + 1j * want[..., 1]
        T = want.shape[-1]
        got = emulate(x, wr, wi, n_fft, hop, n_fft // 2, pad_mode, T)
        err = np.abs(got - want).max() / np.abs(want).max()
        print(f"n_fft {n_fft} hop {hop} {window}/{pad_mode}: radix-2 layout vs oracle STFT  {err:.2e}")
        assert err < 1e-6, err   # limited by the fp32 basis: the identities hold to its rounding


if __name__ == "__main__":
    main()
