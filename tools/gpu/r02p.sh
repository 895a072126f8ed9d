#!/bin/bash
# round 2, GPU call P: staged column kernel (variant 9): parity, timings, piece-size matrix
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_stage.py -x -q -m gpu -k "9" > $O/r02p_pytest.log 2>&1
tail -4 $O/r02p_pytest.log
ST=tools/sweep_time.py
: > $O/r02p_sweeps.jsonl
timeout 300 python $ST --workload cfg2 --kernel 9 --tag k9 >> $O/r02p_sweeps.jsonl 2>>$O/r02p_err.log
timeout 300 python $ST --workload cfg5 --kernel 9 --lus 2 --dims 0,3,5 --tag k9 >> $O/r02p_sweeps.jsonl 2>>$O/r02p_err.log
for cap in 24 56 72; do
  AMDG_SC_CAP_KB=$cap timeout 300 python $ST --workload cfg2 --kernel 9 --lus 1,2 --dims 1 --tag cap$cap >> $O/r02p_sweeps.jsonl 2>>$O/r02p_err.log
  AMDG_SC_CAP_KB=$cap timeout 300 python $ST --workload cfg5 --kernel 9 --lus 2 --dims 0,3 --shapes "b>a" --tag cap$cap >> $O/r02p_sweeps.jsonl 2>>$O/r02p_err.log
done
for mt in 12 24; do
  AMDG_SC_MAX_TGT=$mt timeout 300 python $ST --workload cfg2 --kernel 9 --lus 1,2 --dims 1 --tag mt$mt >> $O/r02p_sweeps.jsonl 2>>$O/r02p_err.log
  AMDG_SC_MAX_TGT=$mt timeout 300 python $ST --workload cfg5 --kernel 9 --lus 2 --dims 0,3 --shapes "b>a" --tag mt$mt >> $O/r02p_sweeps.jsonl 2>>$O/r02p_err.log
done
for nc in 1 2; do
  AMDG_COL_NC=$nc timeout 300 python $ST --workload cfg2 --kernel 9 --lus 1,2 --dims 1 --tag nc$nc >> $O/r02p_sweeps.jsonl 2>>$O/r02p_err.log
  AMDG_COL_NC=$nc timeout 300 python $ST --workload cfg5 --kernel 9 --lus 2 --dims 0,3 --shapes "b>a" --tag nc$nc >> $O/r02p_sweeps.jsonl 2>>$O/r02p_err.log
done
timeout 300 python $ST --workload cfg2 --kernel 9 --acc 1 --lus 1 --dims 0,3 --tag k9_acc >> $O/r02p_sweeps.jsonl 2>>$O/r02p_err.log
python - <<'PY'
import json,collections
T=collections.defaultdict(dict)
for l in open('gpurun_out/r02p_sweeps.jsonl'):
    d=json.loads(l); T[(d['workload'],d['shape'],d['t'],d['lu'],d['acc'])][d['tag']]=d['us']
for k,v in T.items():
    print(k, '  '.join('%s:%.1f'%(tag,us) for tag,us in sorted(v.items(), key=lambda x:x[1])))
PY
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sweep_sc -s 8 -c 1 -o /tmp/p_cfg2 -f python $ST --workload cfg2 --kernel 9 --lus 2 --dims 1 > $O/r02p_ncu1.log 2>&1
ncu -i /tmp/p_cfg2.ncu-rep --page details > $O/r02p_cfg2_details.txt
ncu -i /tmp/p_cfg2.ncu-rep --page source --csv > $O/r02p_cfg2_source.csv 2>/dev/null
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sweep_sc -s 8 -c 1 -o /tmp/p_cfg5 -f python $ST --workload cfg5 --kernel 9 --lus 2 --dims 0 --shapes "b>a" > $O/r02p_ncu2.log 2>&1
ncu -i /tmp/p_cfg5.ncu-rep --page details > $O/r02p_cfg5_details.txt
ncu -i /tmp/p_cfg5.ncu-rep --page source --csv > $O/r02p_cfg5_source.csv 2>/dev/null
grep -v "^frame" $O/r02p_err.log | tail -5
