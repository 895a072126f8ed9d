#!/bin/bash
# round 2, GPU call M (8 GPUs): strong scaling of the fibre-partitioned cfg5 stage at 8, 4, 2 ranks (the driver's launch line), parity at every N,
# device time per kind of operation (eager launches, max over ranks)
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
K=${1:-0}
nvidia-smi topo -m > $O/r02m_topo.txt 2>&1
for n in 8 4 2; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29700+n)) bench.py --gpus $n --steps 10 --warmup 3 --kernel $K --breakdown > $O/r02m_bench_n$n.json 2>$O/r02m_err_n$n.log
  python - <<PY
import json
try:
    d=json.loads([l for l in open('$O/r02m_bench_n$n.json') if l.startswith('{')][-1]); c=d['config']
    print('N=$n stage ms %.3f'%d['ms_per_step'], 'value %.3e'%d['value'], 'launches', c['launches_per_stage'], 'barriers', c['barriers_per_stage'], 'parity', c['parity_rel_l2'], 'timeouts', c['barrier_timeouts'], 'exchange MB %.1f'%(c['exchange_bytes_per_stage_all_ranks']/1e6), 'local', c['max_local_elements'], 'e2e ms %.3f'%d['e2e']['ms_per_step'], 'roof', d['roofline'].get('frac'))
    print('   breakdown', {k:round(v,3) for k,v in c.get('breakdown_ms_eager_max_over_ranks',{}).items()})
except Exception as e:
    print('N=$n failed', e)
PY
  grep -v "^frame\|OMP_NUM\|^\*\*\*\|^$\|NCCL version" $O/r02m_err_n$n.log | tail -4
done
python bench.py --no-cpu --no-secondary --steps 10 --kernel $K --breakdown > $O/r02m_bench_n1.json 2>$O/r02m_err_n1.log
python -c "
import json
d=json.load(open('$O/r02m_bench_n1.json')); print('N=1 stage ms', d['ms_per_step'], 'value %.3e'%d['value'], d['config'].get('breakdown_ms_eager_max_over_ranks'))
"
