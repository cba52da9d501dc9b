#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
echo "== tests"; timeout 900 python -m pytest tests -m gpu -q --timeout 300 -k "cqt or CQT or vqt or VQT or cfg3 or cfg4 or sweep or gamma or repeatable" 2>&1 | tail -4
q() { timeout 200 python bench.py --workload $1 --steps 50 --warmup 5 --no-cpu-baseline --no-e2e --no-workloads --no-reference-gpu > gpurun_out/q_$1.json 2>> gpurun_out/q_err.txt
  python -c "
import json; d=json.load(open('gpurun_out/q_$1.json')); r=d['roofline']; print('$1 ms %.4f frac %.3f pipe %.3f hbm %.3f' % (d['ms_per_step'], r['tensor_algorithmic_frac'] or 0, r['tensor_pipe']['frac'] or 0, r['hbm']['frac']))"; }
q cfg3; q cfg4
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
