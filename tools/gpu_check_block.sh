#!/usr/bin/env bash
# GPU pass over the block-partial STFT kernel: full -m gpu suite, then benches (block on / off).
set -u
mkdir -p gpurun_out
echo "== full gpu suite"; timeout 900 python -m pytest tests -m gpu -q --timeout 300 -x 2>&1 | tail -15
for wl in stft2048 cfg2 cfg5; do
  timeout 200 python bench.py --workload $wl --steps 20 --warmup 5 --no-cpu-baseline --no-e2e \
      > gpurun_out/blk_$wl.json 2>> gpurun_out/blk_err.txt
  python - <<PY
import json
d=json.load(open("gpurun_out/blk_$wl.json")); r=d["roofline"]
print("$wl block  ms/step %.4f value %.3e frac %.3f launch_ms %.4f share %.2f launches %d" % (d["ms_per_step"], d["value"], r["frac"] or 0, r["avg_launch_ms"], r["share_of_step"] or 0, d["gpu_launches"]))
PY
done
NNAUDIO_B200_BLOCK=0 timeout 200 python bench.py --workload cfg2 --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/blk_off_cfg2.json 2>> gpurun_out/blk_err.txt
python -c "
import json
d=json.load(open('gpurun_out/blk_off_cfg2.json')); print('cfg2 dense ms/step %.4f' % d['ms_per_step'])"
tail -5 gpurun_out/blk_err.txt
