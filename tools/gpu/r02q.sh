#!/bin/bash
# round 2, GPU call Q: persistent staged column kernel (variant 9): parity, timings, buffer-size x CTA matrix
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_stage.py -x -q -m gpu -k "9" > $O/r02q_pytest.log 2>&1
tail -4 $O/r02q_pytest.log
ST=tools/sweep_time.py
: > $O/r02q_sweeps.jsonl
timeout 200 python $ST --workload cfg2 --kernel 9 --tag k9 >> $O/r02q_sweeps.jsonl 2>>$O/r02q_err.log
timeout 200 python $ST --workload cfg5 --kernel 9 --lus 2 --dims 0,3,5 --tag k9 >> $O/r02q_sweeps.jsonl 2>>$O/r02q_err.log
for cfg in "24 2" "24 3" "24 4" "40 1" "40 2" "64 1" "88 1" "16 4" "16 6"; do
  set -- $cfg
  AMDG_SC_CAP_KB=$1 AMDG_SC_CTAS=$2 timeout 200 python $ST --workload cfg2 --kernel 9 --lus 1,2 --dims 1 --tag cap$1x$2 >> $O/r02q_sweeps.jsonl 2>>$O/r02q_err.log
  AMDG_SC_CAP_KB=$1 AMDG_SC_CTAS=$2 timeout 200 python $ST --workload cfg5 --kernel 9 --lus 2 --dims 0,3 --shapes "b>a" --tag cap$1x$2 >> $O/r02q_sweeps.jsonl 2>>$O/r02q_err.log
done
timeout 200 python $ST --workload cfg2 --kernel 9 --acc 1 --lus 1 --dims 0,3 --tag k9_acc >> $O/r02q_sweeps.jsonl 2>>$O/r02q_err.log
python - <<'PY'
import json,collections
T=collections.defaultdict(dict)
for l in open('gpurun_out/r02q_sweeps.jsonl'):
    d=json.loads(l); T[(d['workload'],d['shape'],d['t'],d['lu'],d['acc'])][d['tag']]=d['us']
for k,v in T.items():
    print(k, '  '.join('%s:%.1f'%(tag,us) for tag,us in sorted(v.items(), key=lambda x:x[1])))
PY
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sweep_sc -s 8 -c 1 -o /tmp/q_cfg2 -f python $ST --workload cfg2 --kernel 9 --lus 2 --dims 1 > $O/r02q_ncu1.log 2>&1
ncu -i /tmp/q_cfg2.ncu-rep --page details > $O/r02q_cfg2_details.txt
ncu -i /tmp/q_cfg2.ncu-rep --page source --csv > $O/r02q_cfg2_source.csv 2>/dev/null
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sweep_sc -s 8 -c 1 -o /tmp/q_cfg5 -f python $ST --workload cfg5 --kernel 9 --lus 2 --dims 0 --shapes "b>a" > $O/r02q_ncu2.log 2>&1
ncu -i /tmp/q_cfg5.ncu-rep --page details > $O/r02q_cfg5_details.txt
ncu -i /tmp/q_cfg5.ncu-rep --page source --csv > $O/r02q_cfg5_source.csv 2>/dev/null
grep -v "^frame" $O/r02q_err.log | tail -5
