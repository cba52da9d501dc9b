#!/usr/bin/env python
"""Executable spec of the block-partial ("sliding") STFT kernel (float64, CPU).

For a periodic Hann window of length N = n_fft and hop = N / R the R hop-sized blocks of a frame
are shared with its neighbours, so the contraction is done ONCE per block instead of once per frame:

    Z_g[k]  = sum_{n < hop} x[g*hop + n] * exp(-2 pi i k n / N)            (un-windowed, K = hop)
    hann[m] = 1/2 - 1/4 e^{+i theta m} - 1/4 e^{-i theta m},  theta = 2 pi / N
    =>  X_t[k] = sum_{j < R} c_k^j * V_j(Z_{t+j})[k],      c_k = exp(-2 pi i k / R)
        V_j(Z)[k] = 1/2 Z[k] - 1/4 w^j Z[k-1] - 1/4 w^-j Z[k+1],           w = exp(+2 pi i / R)

i.e. R times fewer MACs than the dense (frames x n_fft) contraction; the window becomes a 3-tap
filter along the bin axis and the frame sum a combination of R consecutive block rows (for R = 2, 4
only sign flips and re/im swaps).  Bins -1 and N/2+1 (neighbours of the edge bins) are conjugates
of bins 1 and N/2-1; the kernel simply packs them as extra basis rows.
"""
import numpy as np


def stft_dense(x, n_fft, hop):
    pad = n_fft // 2
    xp = np.pad(x, pad, mode="reflect")
    T = (len(xp) - n_fft) // hop + 1
    win = 0.5 - 0.5 * np.cos(2 * np.pi * np.arange(n_fft) / n_fft)
    frames = np.stack([xp[t * hop: t * hop + n_fft] * win for t in range(T)])
    return np.fft.rfft(frames, axis=1)          # (T, F) = re - i*im in the reference's convention


def stft_block(x, n_fft, hop):
    R = n_fft // hop
    assert R * hop == n_fft
    pad = n_fft // 2
    xp = np.pad(x, pad, mode="reflect")
    T = (len(xp) - n_fft) // hop + 1
    n_blocks = T + R - 1
    F = n_fft // 2 + 1
    k = np.arange(-1, F + 1)                     # packed rows: bins -1 .. F
    basis = np.exp(-2j * np.pi * np.outer(k, np.arange(hop)) / n_fft)        # (F+2, hop)
    blocks = xp[: n_blocks * hop].reshape(n_blocks, hop)
    Z = blocks @ basis.T                          # (n_blocks, F+2): the only contraction, K = hop
    Z0, Zm, Zp = Z[:, 1:-1], Z[:, :-2], Z[:, 2:]
    w = np.exp(2j * np.pi / R)
    kk = np.arange(F)
    c = np.exp(-2j * np.pi * kk / R)
    X = np.zeros((T, F), dtype=complex)
    for j in range(R):
        V = 0.5 * Z0 - 0.25 * w ** j * Zm - 0.25 * w ** (-j) * Zp
        X += (c ** j)[None, :] * V[j: j + T]
    return X


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    for n_fft, hop in ((2048, 512), (512, 256), (1024, 128), (2048, 1024), (256, 64)):
        x = rng.standard_normal(hop * 37 + 11)
        a, b = stft_dense(x, n_fft, hop), stft_block(x, n_fft, hop)
        err = np.abs(a - b).max() / np.abs(a).max()
        print(f"n_fft {n_fft} hop {hop} (R = {n_fft // hop}): max-rel {err:.2e}")
        assert err < 1e-12
