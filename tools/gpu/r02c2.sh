#!/bin/bash
# round 2, GPU call C2 (1 GPU): streaming kernel with work lists in shared memory; bench contract with the cfg5 stage as default workload
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_stage.py -x -q -m gpu -k "7" > $O/r02c2_pytest.log 2>&1
tail -4 $O/r02c2_pytest.log
ST=tools/sweep_time.py
python $ST --workload cfg2 --kernel 7 --tag ws2 > $O/r02c2_sweeps.jsonl 2>$O/r02c2_err.log
for it in 4 8; do AMDG_WS_ITEMS=$it python $ST --workload cfg2 --kernel 7 --lus 2 --tag ws2_items$it >> $O/r02c2_sweeps.jsonl 2>>$O/r02c2_err.log; done
AMDG_DIR_MAXM=1 python $ST --workload cfg2 --kernel 7 --lus 2 --dims 0 --tag ws2_maxm1 >> $O/r02c2_sweeps.jsonl 2>>$O/r02c2_err.log
AMDG_DIR_MAXM=8 python $ST --workload cfg2 --kernel 7 --lus 2 --dims 0 --tag ws2_maxm8 >> $O/r02c2_sweeps.jsonl 2>>$O/r02c2_err.log
AMDG_DIR_MINM=16 python $ST --workload cfg2 --kernel 7 --lus 2 --dims 0 --tag ws2_m16plus >> $O/r02c2_sweeps.jsonl 2>>$O/r02c2_err.log
python $ST --workload cfg5 --kernel 7 --lus 2 --dims 0,3,5 --tag ws2 >> $O/r02c2_sweeps.jsonl 2>>$O/r02c2_err.log
cat $O/r02c2_sweeps.jsonl
python bench.py --no-cpu --no-secondary --steps 10 > $O/r02c2_bench_cfg5_quick.json 2>>$O/r02c2_err.log
python -c "
import json
d=json.load(open('$O/r02c2_bench_cfg5_quick.json')); print('cfg5 stage ms', d['ms_per_step'], 'value', d['value'], 'launches/stage', d['config']['launches_per_stage'], 'parity', d['config']['parity_rel_l2'], 'roof', d['roofline']['frac'], d['roofline']['us_per_launch'], 'e2e ms', d['e2e']['ms_per_step'])
"
python bench.py --no-cpu --no-secondary --steps 10 --kernel 7 > $O/r02c2_bench_cfg5_k7.json 2>>$O/r02c2_err.log
python -c "
import json
d=json.load(open('$O/r02c2_bench_cfg5_k7.json')); print('cfg5 k7 stage ms', d['ms_per_step'], 'parity', d['config']['parity_rel_l2'])
"
python bench.py --workload cfg4 --no-cpu --steps 10 > $O/r02c2_bench_cfg4.json 2>>$O/r02c2_err.log
python -c "
import json
d=json.load(open('$O/r02c2_bench_cfg4.json')); print('cfg4 stage ms', d['ms_per_step'], d['value'], d['config']['launches_per_stage'])
"
( time python bench.py > $O/r02c2_bench_default.json 2>>$O/r02c2_err.log ) 2>&1 | grep real
cat $O/r02c2_bench_default.json | head -c 7000; echo
ncu --set full --clock-control none --import-source on -k regex:sweep_ws -s 8 -c 2 -o $O/r02c2_ws_full python $ST --workload cfg2 --kernel 7 --lus 2 --dims 0 > $O/r02c2_ncu.log 2>&1
grep -v "^frame" $O/r02c2_err.log | tail -5
