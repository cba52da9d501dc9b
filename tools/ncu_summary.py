#!/usr/bin/env python
"""Summarise an `ncu --set full` capture (run here, no GPU needed): key raw metrics of every profiled
launch plus the most-stalled instructions of the source page.  Usage:
    python tools/ncu_summary.py gpurun_out/r02_cfg2_block.ncu-rep > profiles/r02_ncu_cfg2_block.txt"""
import csv
import io
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
    "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "launch__block_size", "launch__cluster_dim_x", "launch__shared_mem_per_block_dynamic",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__inst_executed_pipe_tensor_subpipe_hmma.sum",
]


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main(rep):
    rows = page(rep, "raw")
    hdr, units = rows[0], rows[1]
    print(f"# {rep}: ncu --set full --clock-control none (cold caches, serialised launches: read shares, not absolutes)")
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print(f"\n## launch {d.get('ID')}: {d.get('Kernel Name', '')[:110]}")
        for i, h in enumerate(hdr):
            if h in WANT:
                print(f"{h:78s} {r[i]:>16s} {units[i]}")
    src = page(rep, "source")
    if len(src) > 2:
        h = src[1]
        ix = {k: i for i, k in enumerate(h)}
        data = []
        for r in src[2:]:  # a report with several launches repeats the two header rows: keep the first
            if r and r[0] == "Kernel Name":
                break
            if len(r) == len(h):
                data.append(r)
        stalls = [k for k in h if k.startswith("stall_") and "Not Issued" not in k]
        tot = sum(int(r[ix["# Samples"]] or 0) for r in data) or 1
        agg = sorted(((sum(int(r[ix[k]] or 0) for r in data), k) for k in stalls), reverse=True)[:8]
        print("\n## warp-stall samples, whole kernel (first profiled launch):",
              ", ".join(f"{k[6:]} {100 * v / tot:.0f}%" for v, k in agg))
        print("## most-sampled instructions (samples, executed, SASS, top stall)")
        for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]] or 0))[:16]:
            st = max(((int(r[ix[k]] or 0), k) for k in stalls))
            print(f"{r[ix['# Samples']]:>6s} {r[ix['Instructions Executed']]:>9s}  {r[ix['Source']][:64]:64s} {st[1][6:]}")


if __name__ == "__main__":
    main(sys.argv[1])
