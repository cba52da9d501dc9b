"""Host-side replay of the column arithmetic behind the dense-filterbank path (csrc/tcb_kernels.cu FMT_PLANES
epilogue, csrc/simt_kernels.cu fb_tile_bank_kernel, csrc/nnab_api.cu fb_planes_layout): the block-partial kernel
stores the power spectrum of a frame in ITS tile order — tile n, packed column i <-> FFT bin n (nb - 2) + i - 2, the
first two columns of every tile and the bins >= F as zeros — and the bank is re-indexed to the same order, filters
[0, fh) in the real rows, [fh, 2 fh) negated in the imaginary rows (the contraction returns -sum x w_im).  The same
integer arithmetic in NumPy must reproduce ``fb @ power`` for every tile geometry the launcher can choose."""
import numpy as np
import pytest


def choose_nb(F):
    """block_choose_nb (tcb_kernels.cu): fewest padded columns, wider tiles on ties."""
    best, best_cost = 32, 1 << 30
    for nb in range(128, 31, -8):
        tiles = (F + nb - 3) // (nb - 2)
        cost = tiles * (nb + 6)
        if cost < best_cost:
            best_cost, best = cost, nb
    return best, (F + best - 3) // (best - 2)


def planes_row(power, nb, n_tiles, kp):
    """What the FMT_PLANES epilogue writes for one frame (values only; the bf16 hi/lo split is checked on the GPU)."""
    F = power.shape[0]
    row = np.full(kp, np.nan)                 # columns nobody writes stay NaN here ...
    for n in range(n_tiles):
        for c in range(nb // 8):
            for e in range(8):
                k = n * (nb - 2) + 8 * c - 2 + e
                ok = not (c == 0 and e < 2) and k < F
                row[nb * n + 8 * c + e] = power[k] if ok else 0.0
    row[nb * n_tiles:] = 0.0                  # ... except the tail the launcher clears (cudaMemset2DAsync)
    return row


def tile_bank(fb, nb, n_tiles, kp):
    n_fb, F = fb.shape
    fh = (n_fb + 1) // 2
    w_re, w_im = np.zeros((fh, kp)), np.zeros((fh, kp))
    for j in range(fh):
        for col in range(kp):
            n, i = divmod(col, nb)
            if n < n_tiles and i >= 2:
                k = n * (nb - 2) + i - 2
                if k < F:
                    w_re[j, col] = fb[j, k]
                    if j + fh < n_fb:
                        w_im[j, col] = -fb[j + fh, k]
    return w_re, w_im, fh


@pytest.mark.parametrize("F,n_fb", [(1025, 64), (513, 32), (257, 33), (129, 7), (2049, 96), (65, 1)])
def test_tile_ordered_planes_times_reindexed_bank_is_the_filterbank_product(F, n_fb):
    rng = np.random.RandomState(F + n_fb)
    nb, n_tiles = choose_nb(F)
    kp = (nb * n_tiles + 63) // 64 * 64
    assert nb % 8 == 0 and n_tiles * (nb - 2) >= F and (nb * 2) % 16 == 0   # every chunk starts 16-byte aligned
    power = rng.rand(F)
    fb = rng.standard_normal((n_fb, F))
    row = planes_row(power, nb, n_tiles, kp)
    assert np.isfinite(row).all(), "every column of a plane row is written (or cleared) before the second launch"
    w_re, w_im, fh = tile_bank(fb, nb, n_tiles, kp)
    re, im = w_re @ row, -(w_im @ row)         # complex contraction: (sum x w_re, -sum x w_im)
    out = np.zeros(n_fb)
    out[:fh] = re                               # FMT_REALPAIR: re -> row f, im -> row f + fh (when it exists)
    out[fh:] = im[: n_fb - fh]
    np.testing.assert_allclose(out, fb @ power, rtol=1e-12, atol=1e-12)
