#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
echo "== pyramid tests"; timeout 900 python -m pytest tests -m gpu -q --timeout 300 -k "cqt2010 or vqt or VQT or CQT2010 or cfg4 or sweep-cqt-2010 or gamma or pyramid" 2>&1 | tail -8
q() { timeout 200 python bench.py --workload $1 --steps 30 --warmup 5 --no-cpu-baseline --no-e2e --no-workloads --no-reference-gpu > gpurun_out/q_$2.json 2>> gpurun_out/q_err.txt
  python -c "
import json; d=json.load(open('gpurun_out/q_$2.json')); r=d['roofline']; print('$2 ms %.4f hbm_frac %.3f launches/step %.1f' % (d['ms_per_step'], r['hbm']['frac'], d['gpu_launches']/d['steps']))"; }
q cfg4 cfg4_pyr2
NNAB_PYRAMID2=0 q cfg4 cfg4_pyr1
echo "== launch list cfg4"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 40 --csv --log-file gpurun_out/r02_launches_cfg4.csv python bench.py --workload cfg4 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-workloads --no-reference-gpu > /dev/null 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/r02_launches_cfg4.csv')))
hdr=[r for r in rows if r and r[0]=='ID'][0]; i=rows.index(hdr)
ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
for r in rows[i+2:i+2+40]:
    if len(r)>vi: print(r[ki][:70], r[vi])
PY
tail -3 gpurun_out/q_err.txt
