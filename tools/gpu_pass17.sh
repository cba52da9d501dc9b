#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
echo "== pyramid tests"; timeout 600 python -m pytest tests -m gpu -q --timeout 300 -k "cqt2010 or vqt or VQT or CQT2010 or cfg4 or sweep-cqt-2010 or gamma or repeatable" 2>&1 | tail -4
timeout 200 python bench.py --workload cfg4 --steps 50 --warmup 5 --no-cpu-baseline --no-e2e --no-workloads --no-reference-gpu > gpurun_out/q_cfg4.json 2>> gpurun_out/q_err.txt; python -c "
import json; d=json.load(open('gpurun_out/q_cfg4.json')); print('cfg4 ms %.4f hbm frac %.3f' % (d['ms_per_step'], d['roofline']['hbm']['frac']))"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 36 --csv --log-file gpurun_out/r02_launches_cfg4.csv python bench.py --workload cfg4 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-workloads --no-reference-gpu > /dev/null 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/r02_launches_cfg4.csv')))
hdr=[r for r in rows if r and r[0]=='ID'][0]; i=rows.index(hdr)
ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
agg={}
for r in rows[i+2:]:
    if len(r)>vi:
        k=r[ki].split('(')[0][:40]; agg.setdefault(k,[]).append(float(r[vi])/1000)
for k,v in agg.items(): print('%-42s n=%2d total %.0f us  [%s]' % (k,len(v),sum(v),' '.join('%.0f'%x for x in v[:9])))
PY
tail -3 gpurun_out/q_err.txt
