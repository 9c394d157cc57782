#!/bin/bash
# round 2, session 3: pair kernel for the evaluation chunk maxima + two-level selection + coalesced exact scoring
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
T=${TAG:-e1}
timeout 900 python -m pytest tests -m gpu -x -q -k "eval or topk or tie" > gpurun_out/${T}_tests.txt 2>&1
tail -5 gpurun_out/${T}_tests.txt
for w in eval eval70k; do
timeout 300 python bench.py --workload $w --no-cpu-baseline > gpurun_out/${T}_$w.json 2>gpurun_out/${T}_$w.err; tail -2 gpurun_out/${T}_$w.err
python -c "
import json; d=json.load(open('gpurun_out/${T}_$w.json')); print('$w', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['metrics_sample'])"
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --workload eval --no-cpu-baseline --steps 4 --warmup 3 > gpurun_out/${T}_ncu.log 2>&1
python - <<PY
import csv,collections
rows=list(csv.reader(open('gpurun_out/${T}_launches.csv')))
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
