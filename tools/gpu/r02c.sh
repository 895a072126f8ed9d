#!/bin/bash
# round 2, GPU call C (1 GPU): pieces of the batched stage program, bench contract with the cfg5 stage as default workload
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
timeout 1200 python -m pytest tests/test_gpu_stage.py -x -q -m gpu > $O/r02c_pytest.log 2>&1
tail -15 $O/r02c_pytest.log
python bench.py --no-cpu --no-secondary --steps 10 > $O/r02c_bench_cfg5_quick.json 2>$O/r02c_err.log
tail -3 $O/r02c_err.log
python -c "
import json
d=json.load(open('$O/r02c_bench_cfg5_quick.json')); print('cfg5 stage ms', d['ms_per_step'], 'value', d['value'], 'launches/stage', d['config']['launches_per_stage'], 'parity', d['config']['parity_rel_l2'], 'roof', d['roofline']['frac'], d['roofline']['us_per_launch'], 'e2e ms', d['e2e']['ms_per_step'])
"
python bench.py --no-cpu --no-secondary --steps 10 --kernel 7 > $O/r02c_bench_cfg5_k7.json 2>>$O/r02c_err.log
python -c "
import json
d=json.load(open('$O/r02c_bench_cfg5_k7.json')); print('cfg5 k7 stage ms', d['ms_per_step'], 'parity', d['config']['parity_rel_l2'])
"
python bench.py --workload cfg4 --no-cpu --steps 10 > $O/r02c_bench_cfg4.json 2>>$O/r02c_err.log
python -c "
import json
d=json.load(open('$O/r02c_bench_cfg4.json')); print('cfg4 stage ms', d['ms_per_step'], d['value'], d['config']['launches_per_stage'])
"
( time python bench.py > $O/r02c_bench_default.json 2>>$O/r02c_err.log ) 2>&1 | grep real
( time python bench.py --impl reference > $O/r02c_bench_ref.json 2>>$O/r02c_err.log ) 2>&1 | grep real
cat $O/r02c_bench_default.json | head -c 6000; echo
cat $O/r02c_bench_ref.json
tail -5 $O/r02c_err.log
