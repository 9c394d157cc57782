import sys
lines=[l for l in open(sys.argv[1]) if 'timeline]   end' in l]
# keep only the LAST dumped step
steps=[i for i,l in enumerate(open(sys.argv[1]).read().splitlines()) if '[timeline] step' in l]
prev=0; agg={}; order=[]
for l in lines:
    f=l.split(); t=float(f[2]); name=f[-1]
    if t < prev: prev=0; agg={}; order=[]
    d=t-prev; prev=t
    if name not in agg: agg[name]=[0,0]; order.append(name)
    agg[name][0]+=d; agg[name][1]+=1
for k in order: print(f"{k:28s} n={agg[k][1]:3d} total={agg[k][0]:8.1f} us")
print('launches', sum(v[1] for v in agg.values()), 'serial us', prev)
