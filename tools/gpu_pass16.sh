#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
echo "== tests"; timeout 900 python -m pytest tests -m gpu -q --timeout 300 -k "mel or mfcc or MFCC or Mel or repeatable or cfg2 or cfg5 or cqt2010 or vqt or VQT or CQT2010 or cfg4 or sweep-cqt-2010 or gamma or host_pipeline or output_into" 2>&1 | tail -4
q() { timeout 200 python bench.py --workload $1 --steps 50 --warmup 5 --no-cpu-baseline --no-e2e --no-workloads --no-reference-gpu > gpurun_out/q_$2.json 2>> gpurun_out/q_err.txt
  python -c "
import json; d=json.load(open('gpurun_out/q_$2.json')); r=d['roofline']; print('$2 ms %.4f frac %.3f pipe %.3f launch %.4f' % (d['ms_per_step'], r['frac'], r['tensor_pipe']['frac'], r['avg_launch_ms']))"; }
q cfg2 cfg2; q cfg5 cfg5; q cfg4 cfg4; q stft2048 stft2048
tail -3 gpurun_out/q_err.txt
