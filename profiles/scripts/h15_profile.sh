#!/bin/bash
# H-Codec-1.5 adaptive leg: bench line, then the per-launch list of one step (ncu, serialised; compare shares)
tag=${1:-h15}
mkdir -p gpurun_out
(timeout 600 python bench.py --workload h15 --steps 3 --warmup 1 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?")
tail -c 600 gpurun_out/${tag}_bench.err; cat gpurun_out/${tag}_bench.json | head -c 1500; echo
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --workload h15 --steps 2 --warmup 1 > gpurun_out/${tag}_ncu.log 2>&1; echo "ncu rc=$?"
python - <<PY
import csv, collections, re
rows=[r for r in csv.reader(open("gpurun_out/${tag}_launches.csv")) if len(r)>5 and r[0].isdigit()]
agg=collections.defaultdict(lambda:[0,0.0])
for r in rows:
    name=re.sub(r"\(.*","",r[4]); 
    try: v=float(r[-1].replace(",",""))
    except: continue
    agg[name][0]+=1; agg[name][1]+=v
tot=sum(v[1] for v in agg.values())
print("launches",len(rows),"total ms",tot/1e6)
for k,v in sorted(agg.items(), key=lambda kv:-kv[1][1])[:25]:
    print(f"{100*v[1]/tot:6.2f}% {v[1]/1e6:9.3f} ms {v[0]:6d}  {k[:110]}")
PY
