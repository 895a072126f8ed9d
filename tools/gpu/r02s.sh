#!/bin/bash
# round 2, GPU call S: whole GPU suite; lean vs column kernel on every cfg5 shape (thresholds of the auto mode); bench with the auto mode
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
PKG=$PWD/adaptive-multiresolution-dg_b200
( time timeout 1500 python -m pytest tests -x -q -m gpu > $O/r02s_pytest.log 2>&1 ) 2>&1 | grep real
tail -4 $O/r02s_pytest.log
ST=tools/sweep_time.py
: > $O/r02s_sweeps.jsonl
for k in 5 8; do
  timeout 300 python $ST --workload cfg5 --kernel $k --lus 2 --tag k$k >> $O/r02s_sweeps.jsonl 2>>$O/r02s_err.log
  timeout 300 python $ST --workload cfg5 --kernel $k --lus 0,1 --dims 0,2,4 --shapes "b>a,a>b" --tag k$k >> $O/r02s_sweeps.jsonl 2>>$O/r02s_err.log
  timeout 300 python $ST --workload cfg5 --kernel $k --lus 1 --acc 1 --dims 0,2,4 --shapes "b>a,a>b" --tag k$k >> $O/r02s_sweeps.jsonl 2>>$O/r02s_err.log
done
AMDG_LIB=$PKG/libamdg_b200_touch.so timeout 300 python $ST --workload cfg5 --kernel 8 --lus 2 --tag k8touch >> $O/r02s_sweeps.jsonl 2>>$O/r02s_err.log
for up in 4 16; do AMDG_COL_UPC=$up timeout 300 python $ST --workload cfg5 --kernel 8 --lus 2 --dims 0,3 --shapes "b>a,b>b" --tag k8upc$up >> $O/r02s_sweeps.jsonl 2>>$O/r02s_err.log; done
python - <<'PY'
import json,collections
T=collections.defaultdict(dict)
for l in open('gpurun_out/r02s_sweeps.jsonl'):
    d=json.loads(l); T[(d['workload'],d['shape'],d['t'],d['lu'],d['acc'])][d['tag']]=d['us']
for k,v in sorted(T.items()):
    print(k, '  '.join('%s:%.1f'%(tag,us) for tag,us in sorted(v.items(), key=lambda x:x[1])))
PY
for k in 0 5 8; do
  python bench.py --no-cpu --no-secondary --steps 10 --kernel $k > $O/r02s_bench_k$k.json 2>>$O/r02s_err.log
  python - <<PY
import json
try:
    d=json.load(open('$O/r02s_bench_k$k.json')); c=d['config']
    print('kernel $k: cfg5 stage ms', d['ms_per_step'], 'value %.3e'%d['value'], 'launches', c['launches_per_stage'], 'parity', c['parity_rel_l2'], 'roof', d['roofline']['frac'], d['roofline']['us_per_launch'], 'e2e ms', d['e2e']['ms_per_step'])
except Exception as e: print('kernel $k: no bench', e)
PY
done
AMDG_LIB=$PKG/libamdg_b200_touch.so python bench.py --no-cpu --no-secondary --steps 10 --kernel 0 > $O/r02s_bench_touch.json 2>>$O/r02s_err.log
python -c "
import json
d=json.load(open('$O/r02s_bench_touch.json')); print('touch lib, kernel 0: cfg5 stage ms', d['ms_per_step'])
"
grep -v "^frame" $O/r02s_err.log | tail -5
