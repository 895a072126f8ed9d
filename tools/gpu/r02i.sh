#!/bin/bash
# round 2, GPU call I: parity and first timings of the column kernel (variant 8)
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_stage.py -x -q -m gpu -k "8" > $O/r02i_pytest.log 2>&1
tail -6 $O/r02i_pytest.log
ST=tools/sweep_time.py
python $ST --workload cfg2 --kernel 8 --tag k8 > $O/r02i_sweeps.jsonl 2>$O/r02i_err.log
python $ST --workload cfg5 --kernel 8 --lus 2 --dims 0,3,5 --tag k8 >> $O/r02i_sweeps.jsonl 2>>$O/r02i_err.log
for nc in 1 4; do AMDG_COL_NC=$nc python $ST --workload cfg2 --kernel 8 --lus 2 --dims 0,3 --tag k8_nc$nc >> $O/r02i_sweeps.jsonl 2>>$O/r02i_err.log; done
for up in 4 16; do AMDG_COL_UPC=$up python $ST --workload cfg2 --kernel 8 --lus 2 --dims 0,3 --tag k8_upc$up >> $O/r02i_sweeps.jsonl 2>>$O/r02i_err.log; done
for hv in 12 48; do AMDG_COL_HEAVY=$hv python $ST --workload cfg2 --kernel 8 --lus 2 --dims 0,3 --tag k8_heavy$hv >> $O/r02i_sweeps.jsonl 2>>$O/r02i_err.log; done
python $ST --workload cfg2 --kernel 8 --acc 1 --lus 1 --dims 0,3 --tag k8_acc >> $O/r02i_sweeps.jsonl 2>>$O/r02i_err.log
for nc in 2 3; do AMDG_COL_NC=$nc python $ST --workload cfg5 --kernel 8 --lus 2 --dims 0 --shapes "b>a,b>b" --tag k8_nc$nc >> $O/r02i_sweeps.jsonl 2>>$O/r02i_err.log; done
cat $O/r02i_sweeps.jsonl
python bench.py --no-cpu --steps 10 --kernel 8 > $O/r02i_bench_k8.json 2>>$O/r02i_err.log
python - <<PY
import json
try:
    d=json.load(open('$O/r02i_bench_k8.json')); c=d['config']; s=d.get('secondary',{})
    print('kernel 8: cfg5 stage ms', d['ms_per_step'], 'value %.3e'%d['value'], 'launches', c['launches_per_stage'], 'parity', c['parity_rel_l2'], 'roof', d['roofline']['frac'], d['roofline']['us_per_launch'], 'e2e ms', d['e2e']['ms_per_step'])
    if s: print('   cfg2 ms', s['ms_per_step'], 'value %.3e'%s['value'], 'roof', s['roofline']['frac'], s['roofline']['us_per_launch'], 'e2e ms', s['e2e']['ms_per_step'])
except Exception as e: print('kernel 8: no bench', e)
PY
ncu --set full --clock-control none --import-source on -k regex:sweep_col -s 8 -c 1 -o /tmp/i_cfg2 -f python $ST --workload cfg2 --kernel 8 --lus 2 --dims 1 > $O/r02i_ncu1.log 2>&1
ncu -i /tmp/i_cfg2.ncu-rep --page details > $O/r02i_cfg2_details.txt
ncu -i /tmp/i_cfg2.ncu-rep --page source --csv > $O/r02i_cfg2_source.csv 2>/dev/null
ncu --set full --clock-control none --import-source on -k regex:sweep_col -s 8 -c 1 -o /tmp/i_cfg5 -f python $ST --workload cfg5 --kernel 8 --lus 2 --dims 0 --shapes "b>a" > $O/r02i_ncu2.log 2>&1
ncu -i /tmp/i_cfg5.ncu-rep --page details > $O/r02i_cfg5_details.txt
ncu -i /tmp/i_cfg5.ncu-rep --page source --csv > $O/r02i_cfg5_source.csv 2>/dev/null
grep -v "^frame" $O/r02i_err.log | tail -5
