#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
echo "== tests"; timeout 900 python -m pytest tests -m gpu -q --timeout 300 -x -k "mel or mfcc or MFCC or Mel or repeatable or cfg2 or cfg5 or cqt1992 or CQT1992 or cfg3 or sweep" 2>&1 | tail -6
q() { timeout 200 python bench.py --workload $1 --steps 50 --warmup 5 --no-cpu-baseline --no-e2e --no-workloads --no-reference-gpu > gpurun_out/q_$2.json 2>> gpurun_out/q_err.txt
  python -c "
import json; d=json.load(open('gpurun_out/q_$2.json')); r=d['roofline']; print('$2 ms %.4f frac %.3f pipe %.3f launch %.4f share %.2f' % (d['ms_per_step'], r['frac'], r['tensor_pipe']['frac'], r['avg_launch_ms'], r['share_of_step']))"; }
q cfg2 cfg2; q cfg5 cfg5; q cfg3 cfg3_tall
NNAB_TALL=0 q cfg3 cfg3_varn
tail -3 gpurun_out/q_err.txt
