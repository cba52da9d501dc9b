#!/usr/bin/env python
"""Host-side check of the per-K-block-width plan (branch radix2-wip): for the cfg3 CQT1992v2 bank,
every non-zero tap lies in a block wide enough to include its bin, widths are non-increasing in the
visiting order, the split-K chunks partition the list, and the modelled cost vs the dense kernel."""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import nnaudio_b200 as nb  # noqa: E402
from nnaudio_b200 import _C  # noqa: E402
from nnaudio_b200.features._common import tap_support  # noqa: E402


def plan(kb, ke, F, K, chunks):
    L = _C.lib()
    order = (ctypes.c_int32 * 512)()
    groups = (ctypes.c_int32 * 512)()
    cb = (ctypes.c_int32 * 17)()
    nb_, nc = ctypes.c_int32(0), ctypes.c_int32(0)
    rc = L.nnab_debug_varn_plan(kb.ctypes.data_as(ctypes.c_void_p), ke.ctypes.data_as(ctypes.c_void_p), F, K,
                                chunks, order, groups, cb, ctypes.byref(nb_), ctypes.byref(nc))
    assert rc == 0, rc
    n = nb_.value
    return np.array(order[:n]), np.array(groups[:n]), np.array(cb[: nc.value + 1])


def main():
    for cfg in (dict(sr=44100, n_bins=84, hop_length=512), dict(sr=22050, fmin=220, n_bins=48, hop_length=256)):
        mod = nb.CQT1992v2(verbose=False, **cfg)
        kr, ki = mod.cqt_kernels_real.numpy()[:, 0], mod.cqt_kernels_imag.numpy()[:, 0]
        F, K = kr.shape
        kb, ke = tap_support((kr != 0) | (ki != 0))
        for chunks in (1, 6, 16):
            order, groups, cb = plan(kb, ke, F, K, chunks)
            assert (np.diff(groups) <= 0).all(), "widest block first"
            assert len(set(order.tolist())) == len(order)
            assert cb[0] == 0 and cb[-1] == len(order) and (np.diff(cb) > 0).all()
            width = dict(zip(order.tolist(), groups.tolist()))
            nz = (kr != 0) | (ki != 0)
            for f in range(F):
                for blk in np.unique(np.nonzero(nz[f])[0] // 64):
                    assert width.get(int(blk), 0) * 8 > f, (f, blk)
            cost = np.maximum(16 * groups, 64)
            per_chunk = [int(cost[cb[c]:cb[c + 1]].sum()) for c in range(len(cb) - 1)]
            dense = 16 * ((F + 7) // 8) * ((K + 63) // 64)
            print(f"{cfg['n_bins']} bins, K {K}: {len(order)} active blocks of {(K + 63) // 64}; "
                  f"MMA columns {int((16 * groups).sum())} vs dense {dense} ({dense / (16 * groups).sum():.2f}x), "
                  f"modelled cost {int(cost.sum())} ({dense / cost.sum():.2f}x); chunks {per_chunk}")


if __name__ == "__main__":
    main()
