#!/bin/bash
# round 2, GPU call L: column kernel with cp.async-based L1 prefetch
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
TAG=${1:-v4}
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_stage.py -x -q -m gpu -k "8" > $O/r02l_pytest.log 2>&1
tail -3 $O/r02l_pytest.log
ST=tools/sweep_time.py
: > $O/r02l_sweeps.jsonl
python $ST --workload cfg2 --kernel 8 --tag $TAG >> $O/r02l_sweeps.jsonl 2>>$O/r02l_err.log
python $ST --workload cfg5 --kernel 8 --lus 2 --dims 0,3,5 --tag $TAG >> $O/r02l_sweeps.jsonl 2>>$O/r02l_err.log
for ct in 3 6; do
  AMDG_COL_CTAS=$ct python $ST --workload cfg2 --kernel 8 --lus 1,2 --dims 1 --tag ctas$ct >> $O/r02l_sweeps.jsonl 2>>$O/r02l_err.log
  AMDG_COL_CTAS=$ct python $ST --workload cfg5 --kernel 8 --lus 2 --dims 0,3 --shapes "b>a" --tag ctas$ct >> $O/r02l_sweeps.jsonl 2>>$O/r02l_err.log
done
for nc in 1 2; do
  AMDG_COL_NC=$nc python $ST --workload cfg2 --kernel 8 --lus 1,2 --dims 1 --tag nc$nc >> $O/r02l_sweeps.jsonl 2>>$O/r02l_err.log
  AMDG_COL_NC=$nc python $ST --workload cfg5 --kernel 8 --lus 2 --dims 0,3 --shapes "b>a" --tag nc$nc >> $O/r02l_sweeps.jsonl 2>>$O/r02l_err.log
done
python - <<'PY'
import json,collections
T=collections.defaultdict(dict)
for l in open('gpurun_out/r02l_sweeps.jsonl'):
    d=json.loads(l); T[(d['workload'],d['shape'],d['t'],d['lu'])][d['tag']]=d['us']
for k,v in T.items():
    print(k, '  '.join('%s:%.1f'%(tag,us) for tag,us in sorted(v.items(), key=lambda x:x[1])))
PY
ncu --set full --clock-control none --import-source on -k regex:sweep_col -s 8 -c 1 -o /tmp/l_cfg2 -f python $ST --workload cfg2 --kernel 8 --lus 1 --dims 1 > $O/r02l_ncu1.log 2>&1
ncu -i /tmp/l_cfg2.ncu-rep --page details > $O/r02l_cfg2_details.txt
ncu -i /tmp/l_cfg2.ncu-rep --page source --csv > $O/r02l_cfg2_source.csv 2>/dev/null
grep -v "^frame" $O/r02l_err.log | tail -5
