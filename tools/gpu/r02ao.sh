#!/bin/bash
# round 2, GPU call AO (1 GPU): cost of a grid change after the radix-sort rebuild; where the live adaptive step spends its time (launch list under ncu)
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
python tools/rebuild_cost.py 2 9 2 3 2>&1 | tail -n 1
python tools/rebuild_cost.py 4 8 3 3 2>&1 | tail -n 1
python tools/rebuild_cost.py 6 7 1 2 2>&1 | tail -n 1
examples/live_burgers_adapt -NM 9 -N0 2 -steps 10 -gen 1 2>&1 | grep "wall per step"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/r02ao_live_launches.csv examples/live_burgers_adapt -NM 9 -N0 2 -steps 10 -gen 1 > $O/r02ao_live_under_ncu.log 2>&1
python - <<'PY'
import csv, collections, re
rows=[r for r in csv.reader(open('gpurun_out/r02ao_live_launches.csv')) if len(r)>10]
h=rows[0]; ix={c:i for i,c in enumerate(h)}
agg=collections.OrderedDict()
for r in rows[1:]:
    n=re.sub(r"\(.*","",r[ix["Kernel Name"]]).replace("void ","")
    a=agg.setdefault(n,[0,0.0]); a[0]+=1; a[1]+=float(r[ix["Metric Value"]])
tot=sum(v[1] for v in agg.values()); cnt=sum(v[0] for v in agg.values())
print("launches", cnt, "gpu time ms %.2f"%(tot/1e6))
for n,(c,t) in sorted(agg.items(), key=lambda x:-x[1][1])[:12]: print("  %-60s %5d %9.1f us  avg %.1f us"%(n[:60],c,t/1e3,t/1e3/c))
PY
