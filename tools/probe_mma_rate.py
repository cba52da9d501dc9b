#!/usr/bin/env python
"""GPU probe: clock cycles per K block (4 K16 slices) of tcgen05.mma (bf16, smem operands) vs N and
the operand pattern (csrc/tc_probe.cu:probe_mma_rate_kernel).  Writes gpurun_out/probe_mma_rate.json."""
import ctypes, json, os, sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nnaudio_b200 import _C  # noqa: E402

lib = _C.lib()
lib.nnab_probe_mma_rate.restype = ctypes.c_int
lib.nnab_probe_mma_rate.argtypes = [ctypes.c_int] * 4 + [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
sms = torch.cuda.get_device_properties(0).multi_processor_count
out = torch.zeros(2 * sms, dtype=torch.int64, device="cuda")
iters = 2000
K, SH3, SHV, T2, T3, BROT, ACC2, TWO, SAME_A, TERM_MAJOR, COMMIT = 1, 2, 4, 8, 16, 32, 64, 128, 256, 512, 1024
PATTERNS = [
    ("same A slice, 1 MMA per slice", 0),
    ("K16 slices in turn", K),
    ("slices in turn, A 3 rows down", K | SH3),
    ("slices in turn, A row offset varies", K | SHV),
    ("2 terms (a_lo, a_hi)", K | T2),
    ("3 terms (a_lo, a_hi, a_hi)", K | T3),
    ("3 terms, row offset varies", K | T3 | SHV),
    ("3 terms, row offset varies, B stages rotate", K | T3 | SHV | BROT),
    ("3 terms, row offset varies, B rotates, 2 accumulators", K | T3 | SHV | BROT | ACC2),
    ("same A slice, 3 terms", T3),
    ("2 MMAs per slice, both a_hi", K | T2 | SAME_A),
    ("3 MMAs per slice, all a_hi", K | T3 | SAME_A),
    ("3 terms, term-major (12 MMAs, A changes every time)", K | T3 | TERM_MAJOR),
    ("1 term, two accumulators alternating", K | ACC2),
    ("TWO issuers: 1 term each", K | TWO),
    ("TWO issuers: 3 terms each, row offset varies", K | T3 | SHV | TWO),
    ("3 terms, row offset varies, B rotates + commit per K block", K | T3 | SHV | BROT | COMMIT),
    ("TWO issuers: 3 terms, offsets vary, B rotates + commit per K block", K | T3 | SHV | BROT | TWO | COMMIT),
    ("1 term + commit per K block", K | COMMIT),
]
res = []
for cg in (2,):
    for name, flags in PATTERNS:
        line = []
        for n in (16, 64, 128, 192, 256):
            mmas = {0: 4, 1: 8, 2: 12}[(flags >> 3) & 3]
            for rep in range(2):
                out.zero_()
                rc = lib.nnab_probe_mma_rate(cg, n, flags, iters, out.data_ptr(), sms - sms % cg,
                                             torch.cuda.current_stream().cuda_stream)
                assert rc == 0, (rc, cg, flags, n)
                torch.cuda.synchronize()
            c = float(np.median(out[: sms // cg].cpu().numpy().astype(np.float64)))
            if flags & TWO:  # both issuers ran `iters` K blocks concurrently: cycles per K block of the SM
                c = float(np.median(np.maximum(out[: sms // cg].cpu().numpy(), out[sms // cg: 2 * (sms // cg)].cpu().numpy()))) / 2
            res.append(dict(cta_group=cg, pattern=name, flags=flags, n=n, mmas_per_kblock=mmas,
                            cycles_per_kblock=c / iters, cycles_per_mma=c / iters / mmas))
            line.append("N%d %.0f" % (n, c / iters))
        print("cg %d %-58s clk / K block: %s" % (cg, name, "  ".join(line)), flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "probe_mma_rate.json"), "w"), indent=1)
