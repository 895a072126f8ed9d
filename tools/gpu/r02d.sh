#!/bin/bash
# round 2, GPU call D (2 GPUs): fibre-partitioned stage -- parity at 2 ranks, strong scaling 1 -> 2
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
nvidia-smi topo -m > $O/r02d_topo.txt 2>&1
python bench.py --no-cpu --no-secondary --steps 10 > $O/r02d_bench_n1.json 2>$O/r02d_err1.log
python -c "
import json
d=json.load(open('$O/r02d_bench_n1.json')); print('N=1 stage ms', d['ms_per_step'], 'launches', d['config']['launches_per_stage'], 'parity', d['config']['parity_rel_l2'])
"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655 tests/dist_check.py > $O/r02d_dist_check.log 2>&1
tail -6 $O/r02d_dist_check.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29656 bench.py --gpus 2 --steps 10 --warmup 3 > $O/r02d_bench_n2.json 2>$O/r02d_err2.log
python -c "
import json
d=json.load(open('$O/r02d_bench_n2.json')); c=d['config']; print('N=2 stage ms', d['ms_per_step'], 'value', d['value'], 'launches', c['launches_per_stage'], 'barriers', c['barriers_per_stage'], 'parity', c['parity_rel_l2'], 'timeouts', c['barrier_timeouts'], 'exchange MB', c['exchange_bytes_per_stage_all_ranks']/1e6, 'local', c['max_local_elements'], 'e2e ms', d['e2e']['ms_per_step'])
"
grep -v "^frame" $O/r02d_err2.log | tail -8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29657 bench.py --gpus 2 --steps 10 --warmup 3 --no-graph > $O/r02d_bench_n2_nograph.json 2>>$O/r02d_err2.log
python -c "
import json
d=json.load(open('$O/r02d_bench_n2_nograph.json')); print('N=2 no graph stage ms', d['ms_per_step'])
"
