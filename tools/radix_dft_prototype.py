#!/usr/bin/env python
"""Numerical prototype (CPU, numpy) of the next step for the STFT-family kernel: one or two
decimation-in-time steps in front of the dense DFT contraction.

Today the kernel contracts each n_fft-sample frame with the full windowed basis:
N x (N + 2) MACs per frame (N = n_fft, one-sided, re + im).  Splitting the frame into its R
sample phases (n = R m + r) gives R windowed sub-DFTs of length N/R whose real-input Hermitian
symmetry leaves N/(2R) + 1 unique bins each:

    S_r[k] = sum_m x[R m + r] w[R m + r] exp(-2 pi i k m / (N/R)),      k = 0 .. N/(2R)
    X[k]   = sum_r exp(-2 pi i r k / N) S_r[k mod N/R]                   (S_r[N/R - k] = conj S_r[k])

i.e. R GEMMs with K = N/R and N/(2R)+1 complex columns: N (N/R + 2) MACs per frame, 1/R of
the dense count, plus R complex multiply-adds per output bin in the epilogue (done from TMEM in
registers: all R sub-DFT accumulators of a frame sit in the same TMEM lane).  The A operand of
sub-DFT r is the r-th sample phase of the signal, which TMA can read from R de-interleaved planes
as long as R divides the hop.

This script checks the algebra against the dense contraction in float64 and measures what the
shorter contractions do to the split-bf16 (3-term) error.  It is a design aid, not product code.

    python tools/radix_dft_prototype.py [--n-fft 2048] [--hop 512] [--radix 2 4]
"""
import argparse

import numpy as np


def bf16(v):
    """Round-to-nearest-even float32 -> bfloat16 -> float32 (numpy)."""
    u = np.asarray(v, dtype=np.float32).view(np.uint32)
    r = ((u >> 16) & 1) + 0x7FFF
    return ((u + r) & 0xFFFF0000).astype(np.uint32).view(np.float32)


def split3(a, b):
    """a @ b.T with the 3-term bf16 hi/lo split, products accumulated in float64 (the tensor
    core accumulates in fp32; float64 here isolates the operand-rounding error)."""
    ah = bf16(a); al = bf16(a - ah)
    bh = bf16(b); bl = bf16(b - bh)
    f = np.float64
    return ah.astype(f) @ bh.astype(f).T + al.astype(f) @ bh.astype(f).T + ah.astype(f) @ bl.astype(f).T


def dense_basis(N, window):
    k = np.arange(N // 2 + 1)[:, None]
    n = np.arange(N)[None, :]
    ang = 2 * np.pi * k * n / N
    return (np.cos(ang) * window).astype(np.float32), (np.sin(ang) * window).astype(np.float32)


def radix_bases(N, R, window):
    """Per sample phase r: (cos, sin) rows for k = 0 .. N/(2R), over m = 0 .. N/R - 1."""
    Ns = N // R
    k = np.arange(Ns // 2 + 1)[:, None]
    m = np.arange(Ns)[None, :]
    ang = 2 * np.pi * k * m / Ns
    return [((np.cos(ang) * window[r::R]).astype(np.float32),
             (np.sin(ang) * window[r::R]).astype(np.float32)) for r in range(R)]


def combine(S, N, R):
    """Epilogue: X[k] for k = 0 .. N/2 from the R half-spectra S[r] (complex, (T, N/(2R)+1))."""
    Ns = N // R
    k = np.arange(N // 2 + 1)
    kk = k % Ns
    X = np.zeros((S[0].shape[0], N // 2 + 1), dtype=np.complex128)
    for r in range(R):
        full = np.where(kk <= Ns // 2, S[r][:, np.minimum(kk, Ns // 2)],
                        np.conj(S[r][:, np.minimum(Ns - kk, Ns // 2)]))
        X += np.exp(-2j * np.pi * r * k / N)[None, :] * full
    return X


def radix2_kernel_layout_check(frames, N, window, exact):
    """Executable spec of the planned R = 2 kernel data layout.

    Basis per sample phase r (even / odd samples), 2 * (N/4) rows of length N/2:
        row 2k   : w[2m + r] *  cos(2 pi k m / (N/2))        k = 0 .. N/4 - 1
        row 2k+1 : w[2m + r] * -sin(2 pi k m / (N/2))        k = 1 .. N/4 - 1
        row 1    : w[2m + r] * (-1)^m      <- the sub-DFT's Nyquist bin (k = N/4, real) stored in
                                              the always-zero imaginary slot of k = 0
    so N/4 (re, im) column pairs cover all N/4 + 1 unique bins and tile evenly (512 -> 8 x 64).
    Epilogue per frame and k (all four accumulators sit in the same TMEM lane):
        U = W_N^k * S1[k]                    X[k]       = S0[k] + U
                                             X[N/2 - k] = conj(S0[k] - U)
        k = 0:  X[0] = S0re + S1re,  X[N/2] = S0re - S1re,
                X[N/4] = S0nyq - i * S1nyq   (W_N^{N/4} = -i)
    Returns max |X - exact| / max |exact| in float64."""
    f = np.float64
    Ns, Q = N // 2, N // 4
    m = np.arange(Ns)
    acc = []
    for r in range(2):
        rows = np.zeros((2 * Q, Ns), dtype=f)
        wr = window[r::2]
        for k in range(Q):
            rows[2 * k] = wr * np.cos(2 * np.pi * k * m / Ns)
            rows[2 * k + 1] = -wr * np.sin(2 * np.pi * k * m / Ns)
        rows[1] = wr * (-1.0) ** m
        acc.append(frames[:, r::2].astype(f) @ rows.astype(np.float32).astype(f).T)   # (T, 2Q)
    T = frames.shape[0]
    X = np.zeros((T, N // 2 + 1), dtype=np.complex128)
    S0, S1 = acc
    for k in range(1, Q):
        s0 = S0[:, 2 * k] + 1j * S0[:, 2 * k + 1]
        s1 = S1[:, 2 * k] + 1j * S1[:, 2 * k + 1]
        u = np.exp(-2j * np.pi * k / N) * s1
        X[:, k] = s0 + u
        X[:, N // 2 - k] = np.conj(s0 - u)
    X[:, 0] = S0[:, 0] + S1[:, 0]
    X[:, N // 2] = S0[:, 0] - S1[:, 0]
    X[:, Q] = S0[:, 1] - 1j * S1[:, 1]
    return np.abs(X - exact).max() / np.abs(exact).max()


def radix4_module_basis_check(frames, N, wc, ws, exact):
    """R = 4 with the module's own rows (twiddles already inside the odd phases):
        U_r[k] = sum_m x[4m + r] (w_cos - i w_sin)[k][4m + r],      k = 0 .. N/8
        X[k]        =       U0 +   U1 + U2 +   U3        X[N/4 + k] = U0 - i U1 - U2 + i U3
        X[N/4 - k]  = conj(U0) - i conj(U1) - conj(U2) + i conj(U3)
        X[N/2 - k]  = conj(U0 - U1 + U2 - U3)
    — additions, sign flips and re/im swaps only."""
    f = np.float64
    Q = N // 8
    U = [frames[:, r::4].astype(f) @ wc[: Q + 1, r::4].astype(f).T
         - 1j * (frames[:, r::4].astype(f) @ ws[: Q + 1, r::4].astype(f).T) for r in range(4)]
    X = np.zeros_like(exact)
    k = np.arange(Q + 1)
    X[:, k] = U[0] + U[1] + U[2] + U[3]
    X[:, N // 4 + k] = U[0] - 1j * U[1] - U[2] + 1j * U[3]
    X[:, N // 4 - k] = np.conj(U[0]) - 1j * np.conj(U[1]) - np.conj(U[2]) + 1j * np.conj(U[3])
    X[:, N // 2 - k] = np.conj(U[0] - U[1] + U[2] - U[3])
    return np.abs(X - exact).max() / np.abs(exact).max()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n-fft", type=int, default=2048)
    ap.add_argument("--hop", type=int, default=512)
    ap.add_argument("--frames", type=int, default=256)
    ap.add_argument("--radix", type=int, nargs="+", default=[2, 4, 8])
    args = ap.parse_args()
    N, hop, T = args.n_fft, args.hop, args.frames
    rng = np.random.RandomState(0)
    x = rng.standard_normal(hop * (T - 1) + N).astype(np.float32)
    frames = np.lib.stride_tricks.sliding_window_view(x, N)[::hop][:T]
    from scipy.signal import get_window
    window = get_window("hann", N, fftbins=True)

    wc, ws = dense_basis(N, window)
    f = np.float64
    exact = frames.astype(f) @ wc.astype(f).T - 1j * (frames.astype(f) @ ws.astype(f).T)
    scale = np.abs(exact).max()
    d3 = split3(frames, wc) - 1j * split3(frames, ws)
    print(f"n_fft {N}, hop {hop}, {T} frames of white noise; errors are max|d| / max|X|")
    print(f"dense   : MACs/frame {N * (N + 2):>9d}   split-bf16 error {np.abs(d3 - exact).max() / scale:.2e}")
    for R in args.radix:
        if N % (2 * R) or hop % R:
            print(f"radix {R}: needs 2R | n_fft and R | hop — skipped")
            continue
        bases = radix_bases(N, R, window)
        S64, S3 = [], []
        for r, (bc, bs) in enumerate(bases):
            a = np.ascontiguousarray(frames[:, r::R])
            S64.append(a.astype(f) @ bc.astype(f).T - 1j * (a.astype(f) @ bs.astype(f).T))
            S3.append(split3(a, bc) - 1j * split3(a, bs))
        X64, X3 = combine(S64, N, R), combine(S3, N, R)
        macs = R * (N // R) * 2 * (N // (2 * R) + 1)
        print(f"radix {R:<2d}: MACs/frame {macs:>9d} ({N * (N + 2) / macs:.2f}x fewer)   "
              f"algebra error {np.abs(X64 - exact).max() / scale:.1e}   "
              f"split-bf16 error {np.abs(X3 - exact).max() / scale:.2e}   "
              f"epilogue {R} complex MADs per bin")
    print(f"radix 4 with the module's own basis rows (add / swap butterflies only): "
          f"error {radix4_module_basis_check(frames, N, wc, ws, exact):.1e}")
    print(f"radix 2 kernel layout (Nyquist packed into the k = 0 imaginary slot, butterfly epilogue): "
          f"error {radix2_kernel_layout_check(frames, N, window, exact):.1e}")


if __name__ == "__main__":
    main()
