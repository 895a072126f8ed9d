#!/bin/bash
# round 2, GPU call E (1 GPU): live-reference drop-in run, per-kernel breakdown of the cfg5 stage
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
examples/live_burgers_adapt -NM 6 -N0 2 -steps 10 > $O/r02e_live_n6.log 2>&1; tail -6 $O/r02e_live_n6.log
examples/live_burgers_adapt -NM 9 -N0 2 -steps 10 > $O/r02e_live_n9.log 2>&1; tail -4 $O/r02e_live_n9.log
timeout 900 python -m pytest tests/test_gpu_stage.py -x -q -m gpu > $O/r02e_pytest.log 2>&1; tail -3 $O/r02e_pytest.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $O/r02e_launches_cfg5.csv python bench.py --no-cpu --no-secondary --steps 1 --warmup 3 --no-graph > $O/r02e_under_ncu.log 2>&1
python - <<'PY'
import csv, collections, re
rows=[r for r in csv.reader(open('gpurun_out/r02e_launches_cfg5.csv')) if len(r)>10]
hdr=rows[0]; ix={h:i for i,h in enumerate(hdr)}
data=rows[1:]
# keep the last stage only: find launches after the last rk_stage_kernel but one
names=[r[ix['Kernel Name']] for r in data]
rk=[i for i,n in enumerate(names) if 'rk_stage' in n]
print('launches', len(data), 'rk positions', rk[-3:])
seg=data[rk[-2]+1:rk[-1]+1] if len(rk)>=2 else data
agg=collections.OrderedDict()
for r in seg:
    n=re.sub(r'\(.*','',r[ix['Kernel Name']]); t=float(r[ix['Metric Value']])
    k=(n, r[ix['Grid Size']].split(',')[1].strip() if ',' in r[ix['Grid Size']] else '')
    a=agg.setdefault(n,[0,0.0]); a[0]+=1; a[1]+=t
tot=sum(v[1] for v in agg.values())
for n,v in sorted(agg.items(), key=lambda x:-x[1][1]): print('%-70s %4d launches %10.1f us %5.1f%%'%(n[:70], v[0], v[1]/1e3, 100*v[1]/tot))
print('stage total (serialised, cold) %.2f ms over %d launches'%(tot/1e6, len(seg)))
PY
