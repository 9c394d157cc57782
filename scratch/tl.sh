#!/bin/bash
# serial per-kernel durations (us) of the head kernels from REC_TIMELINE: scratch/tl.sh <workload> [env...]
wl=$1; shift
env "$@" REC_TIMELINE=1 REC_NO_OVERLAP=1 timeout 150 python bench.py --workload $wl --steps 50 --warmup 5 --no-cpu-baseline 2>&1 | grep timeline | head -24 | python -c "
import sys
prev=0.0
out=[]
for line in sys.stdin:
    f=line.split()
    if f[1]!='end': continue
    t=float(f[2]); name=f[-1]
    out.append('%s=%.0f' % (name, t-prev)); prev=t
print('$wl $*', ' '.join(o for o in out if 'heads' in o), 'total=%.0f' % prev)
"
