mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/nsga2_launches.csv python scripts/bench_nsga2.py 65536 > gpurun_out/nsga2_ncu.log 2>&1
python - <<PY
import csv, collections
rows=list(csv.reader(open('gpurun_out/nsga2_launches.csv')))
i=[k for k,r in enumerate(rows) if 'Kernel Name' in r][0]
hdr=rows[i]; kn=hdr.index('Kernel Name'); mv=hdr.index('Metric Value'); mu=hdr.index('Metric Unit')
agg=collections.defaultdict(lambda:[0,0.0])
for r in rows[i+1:]:
    if len(r)<=mv: continue
    v=float(r[mv].replace(',',''))
    u=r[mu]
    if u in ('ns','nsecond'): v/=1e3
    elif u in ('ms','msecond'): v*=1e3
    elif u in ('s','second'): v*=1e6
    a=agg[r[kn][:80]]; a[0]+=1; a[1]+=v
tot=sum(t for c,t in agg.values())
print("total %.2f ms over %d launches"%(tot/1e3,sum(c for c,t in agg.values())))
for k,(c,t) in sorted(agg.items(), key=lambda kv:-kv[1][1])[:22]:
    print(f"{t/1e3:9.3f} ms {c:6d} launches {t/c:8.2f} us avg  {k}")
PY
