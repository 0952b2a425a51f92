#!/bin/bash
# usage (on the GPU box): scripts/ncu_variants.sh v1 v2 ...  -> median gpu__time_duration per variant/kernel
for v in "$@"; do
  export CTD_LIB=ctdirect.jl_b200/lib/variants/libctd_$v.so
  KTIME_K=40 python scripts/ktime.py > gpurun_out/kt_$v.log 2>&1 || { echo "$v plain run failed"; tail -3 gpurun_out/kt_$v.log; continue; }
  KTIME_K=40 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none --csv --log-file gpurun_out/ncu_$v.csv python scripts/ktime.py > /dev/null 2>&1
  python - <<PY
import csv, collections
rows=[r for r in csv.reader(open("gpurun_out/ncu_$v.csv")) if len(r)>10 and r[0].isdigit()]
d=collections.defaultdict(list); ins={}
for r in rows:
    if r[-3]=="gpu__time_duration.sum": d[r[4]].append(float(r[-1]))
    else: ins[r[4]]=float(r[-1].replace(",",""))
out=[]
for k,v in d.items():
    if "ctd_eval" not in k: continue
    v=sorted(v[10:]); f=k.split(",")[1].split(">")[0].strip()
    out.append(f"F={f}:{v[len(v)//2]/1e3:.2f}us/{ins.get(k,0)/1e6:.2f}Mi")
print("$v", " ".join(out))
PY
done
