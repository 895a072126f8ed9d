#!/bin/bash
# round 2, GPU call AE (2 GPUs): the partitioned stage with the RK combination fused (parity at N = 2 inside dist_check and bench), GPU suite incl. the 2-GPU test
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29811 tests/dist_check.py > $O/r02ae_dist_check.log 2>&1; tail -n 4 $O/r02ae_dist_check.log
for f in 0 1; do
  extra=""; [ $f = 0 ] && extra="--no-fuse-rk"
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29820+f)) bench.py --gpus 2 --steps 10 --warmup 3 $extra > $O/r02ae_bench_n2_f$f.json 2>$O/r02ae_err_n2_f$f.log
  python - <<PY
import json
try:
    d=json.loads([l for l in open('$O/r02ae_bench_n2_f$f.json') if l.startswith('{')][-1]); c=d['config']
    print('N=2 fuse $f stage ms %.3f'%d['ms_per_step'], 'value %.3e'%d['value'], 'launches', c['launches_per_stage'], 'parity', c['parity_rel_l2'], 'timeouts', c['barrier_timeouts'], 'e2e ms %.3f'%d['e2e']['ms_per_step'], c.get('rk_update'))
except Exception as e:
    print('N=2 failed', e)
PY
done
for f in 0 1; do
  extra=""; [ $f = 0 ] && extra="--no-fuse-rk"
  python bench.py --no-cpu --no-secondary --steps 10 $extra > $O/r02ae_bench_n1_f$f.json 2>$O/r02ae_err_n1_f$f.log
  python -c "
import json
d=json.loads([l for l in open('$O/r02ae_bench_n1_f$f.json') if l.startswith('{')][-1]); print('N=1 fuse $f stage ms', d['ms_per_step'], 'value %.3e'%d['value'], 'launches', d['config']['launches_per_stage'], 'parity', d['config']['parity_rel_l2'], 'e2e', d['e2e']['ms_per_step'])
"
done
timeout 600 python -m pytest tests/test_gpu_stage.py tests/test_gpu_parity.py -x -q -m gpu -k "stage_program or multi_gpu or fibre" 2>&1 | tail -n 3
timeout 20 nvidia-smi --query-gpu=name,memory.used --format=csv,noheader
