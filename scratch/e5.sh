#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "eval or topk or tie" 2>&1 | tail -6
for w in eval eval70k; do
timeout 300 python bench.py --workload $w --no-cpu-baseline > gpurun_out/e5_$w.json 2>gpurun_out/e5_$w.err; tail -2 gpurun_out/e5_$w.err
python -c "
import json; d=json.load(open('gpurun_out/e5_$w.json')); print('$w', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['clocks'])"
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 200 --csv --log-file gpurun_out/e5_launches.csv python bench.py --workload eval --no-cpu-baseline --steps 8 --warmup 3 > gpurun_out/e5_ncu.log 2>&1
python - <<PY
import csv,collections
rows=list(csv.reader(open('gpurun_out/e5_launches.csv')))
for i,r in enumerate(rows):
    if 'Kernel Name' in r: h=r; start=i; break
idx={k:j for j,k in enumerate(h)}
agg=collections.OrderedDict()
for r in rows[start+1:]:
    if len(r)<len(h) or r[idx['Metric Name']]!='gpu__time_duration.sum': continue
    n=r[idx['Kernel Name']][:60]; v=float(r[idx['Metric Value']].replace(',',''))
    a=agg.setdefault(n,[0,0.0]); a[0]+=1; a[1]+=v
for n,(c,t) in agg.items(): print(f'{c:4d} {t/c/1000:10.1f} us {n}')
PY
