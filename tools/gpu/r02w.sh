#!/bin/bash
# round 2, GPU call W (2 GPUs): parallel streams per schedule level -- parity and timing at N = 1 and 2; cfg1 / cfg3 workloads
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_stage.py -x -q -m gpu > $O/r02w_pytest.log 2>&1; tail -2 $O/r02w_pytest.log
for ns in 0 1 2 3; do
  AMDG_STAGE_STREAMS=$ns python bench.py --no-cpu --no-secondary --steps 10 > $O/r02w_bench_n1_s$ns.json 2>>$O/r02w_err.log
  python -c "
import json
d=json.load(open('$O/r02w_bench_n1_s$ns.json')); print('streams $ns N=1 stage ms %.3f'%d['ms_per_step'], 'parity', d['config']['parity_rel_l2'], 'roof', d['roofline']['frac'])
"
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655 tests/dist_check.py > $O/r02w_dist_check.log 2>&1
tail -3 $O/r02w_dist_check.log
for ns in 0 2; do
  AMDG_STAGE_STREAMS=$ns timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2966$ns bench.py --gpus 2 --steps 10 --warmup 3 > $O/r02w_bench_n2_s$ns.json 2>>$O/r02w_err.log
  python -c "
import json
d=json.loads([l for l in open('$O/r02w_bench_n2_s$ns.json') if l.startswith('{')][-1]); print('streams $ns N=2 stage ms %.3f'%d['ms_per_step'], 'parity', d['config']['parity_rel_l2'], 'timeouts', d['config']['barrier_timeouts'])
"
done
for wl in cfg1 cfg3; do
  python bench.py --workload $wl --steps 20 > $O/r02w_bench_$wl.json 2>>$O/r02w_err.log
  python -c "
import json
d=json.load(open('$O/r02w_bench_$wl.json')); print('$wl stage ms %.4f'%d['ms_per_step'], 'value %.3e'%d['value'], 'e2e %.3e'%d['e2e']['value'], 'roof us', d['roofline']['us_per_launch'], 'cpu', d.get('cpu_baseline',{}).get('value'))
"
done
grep -v "^frame\|OMP_NUM\|^\*\*\*\|^$\|NCCL version" $O/r02w_err.log | tail -6
