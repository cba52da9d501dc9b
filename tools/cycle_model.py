#!/usr/bin/env python
"""Back-of-the-envelope cycle model of the framed-contraction mainloop (CTA-pair kernel), fitted to
the two measured points of round 1 (1-CTA N=208: 75.9 % tensor pipe; pair N=208: 93.3 %):

  per K16 step and SM:  MMA time      = 3 passes * N/2 clk        (4096 dense bf16 MAC/clk/SM)
                        smem reads    = 3 * (A 128x16 + B N/2 x16) bf16      (pair: half of B per SM)
                        smem writes   = TMA fill of A hi+lo and B hi+lo for that step
                        smem port     = 128 B/clk
  step time = max(MMA time, (reads + writes) / 128)

Prints clk per OUTPUT BIN (128-frame rows) for today's dense kernel and the decimation-in-time
variants, so tile shapes can be compared before spending GPU time.  Epilogue cost is a rough
per-column constant taken from the dense kernel (hidden when TMEM is double-buffered).
"""


def step(n_cols, a_from_tmem=False):
    mma = 3 * n_cols / 2
    a_bytes = 128 * 16 * 2
    b_bytes = (n_cols // 2) * 16 * 2
    # A through TMEM: tcgen05.cp reads the hi and lo tiles from shared memory once each (2 reads)
    # instead of the three MMA passes reading them (hi twice, lo once)
    reads = (2 if a_from_tmem else 3) * a_bytes + 3 * b_bytes
    writes = 2 * a_bytes + 2 * b_bytes
    return max(mma, (reads + writes) / 128.0), mma, (reads + writes) / mma


EPI_CLK_PER_COL = 45.0  # ~ 9.4 k clk for 208 columns of 128 rows (fits under the dense mainloop)


def variant(name, n_fft, radix, n_cols, double_buffered, a_from_tmem=False):
    k_steps = n_fft // radix // 16
    t, mma, bpc = step(n_cols, a_from_tmem)
    mainloop = radix * k_steps * t
    bins_k = n_cols // 2                      # sub-DFT bins per tile
    out_bins = bins_k * radix                 # every sub-DFT bin yields `radix` output bins
    epi = EPI_CLK_PER_COL * n_cols * radix
    tile = max(mainloop, epi) if double_buffered else mainloop + epi
    print(f"{name:44s} step {t:6.0f} clk (MMA {mma:4.0f}, smem {bpc * mma / t / 1:5.0f} B/clk needed {bpc:5.0f})  "
          f"tile {tile / 1e3:6.1f} k clk / {out_bins:4d} bins = {tile / out_bins:6.1f} clk/bin")
    return tile / out_bins


def main():
    n = 2048
    base = variant("dense, N=208, TMEM double-buffered (today)", n, 1, 208, True)
    rows = [
        ("radix 2, N=128 x2 segments, double-buffered", 2, 128, True, False),
        ("radix 2, N=256 x2 segments, single-buffered", 2, 256, False, False),
        ("radix 4, N=64 x4 segments, double-buffered", 4, 64, True, False),
        ("radix 4, N=128 x4 segments, single-buffered", 4, 128, False, False),
        ("radix 2, N=128 x2, A operand from TMEM", 2, 128, True, True),
        ("radix 4, N=64 x4, A operand from TMEM", 4, 64, True, True),
    ]
    for name, r, nc, db, at in rows:
        v = variant(name, n, r, nc, db, at)
        print(f"{'':44s} -> {base / v:.2f}x vs today")


if __name__ == "__main__":
    main()
