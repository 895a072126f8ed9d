#!/bin/bash
# round 2, GPU call B: parity and timings of the warp-specialised streaming kernel (variant 7)
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "kernel_variants and 7" > $O/r02b_pytest1.log 2>&1
tail -5 $O/r02b_pytest1.log
if ! grep -q "passed" $O/r02b_pytest1.log || grep -q "failed" $O/r02b_pytest1.log; then echo "first parity test failed: stop"; exit 1; fi
timeout 1500 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "live or batch or random or full_size" > $O/r02b_pytest2.log 2>&1
tail -8 $O/r02b_pytest2.log
ST=tools/sweep_time.py
python $ST --workload cfg2 --kernel 7 --tag ws > $O/r02b_sweeps.jsonl 2>$O/r02b_err.log
AMDG_TC_PDL=0 python $ST --workload cfg2 --kernel 7 --lus 2 --tag ws_nopdl >> $O/r02b_sweeps.jsonl 2>>$O/r02b_err.log
for it in 6 24; do AMDG_WS_ITEMS=$it python $ST --workload cfg2 --kernel 7 --lus 2 --tag ws_items$it >> $O/r02b_sweeps.jsonl 2>>$O/r02b_err.log; done
for mm in 1 8; do AMDG_DIR_MAXM=$mm python $ST --workload cfg2 --kernel 7 --lus 2 --dims 0,3 --tag ws_maxm$mm >> $O/r02b_sweeps.jsonl 2>>$O/r02b_err.log; done
AMDG_DIR_MINM=16 python $ST --workload cfg2 --kernel 7 --lus 2 --dims 0,3 --tag ws_m16plus >> $O/r02b_sweeps.jsonl 2>>$O/r02b_err.log
python $ST --workload cfg2 --kernel 7 --acc 1 --lus 1 --tag ws_acc >> $O/r02b_sweeps.jsonl 2>>$O/r02b_err.log
python $ST --workload cfg5 --kernel 7 --lus 2 --dims 0,3,5 --tag ws >> $O/r02b_sweeps.jsonl 2>>$O/r02b_err.log
python bench.py --workload cfg2 --kernel 7 --no-cpu > $O/r02b_bench_cfg2_k7.json 2>>$O/r02b_err.log
python bench.py --workload cfg5 --kernel 7 --no-cpu --steps 5 > $O/r02b_bench_cfg5_k7.json 2>>$O/r02b_err.log
ncu --set full --clock-control none --import-source on -k regex:sweep_ws -s 8 -c 2 -o $O/r02b_ws_full python $ST --workload cfg2 --kernel 7 --lus 2 --dims 0 > $O/r02b_ncu.log 2>&1
cat $O/r02b_sweeps.jsonl
python -c "
import json
for f in ('cfg2','cfg5'):
    try:
        d=json.load(open('$O/r02b_bench_%s_k7.json'%f)); print(f, d['ms_per_step'], d['value'], d['roofline']['frac'], d['roofline']['us_per_launch'])
    except Exception as e: print(f, 'no bench', e)
"
tail -5 $O/r02b_err.log
