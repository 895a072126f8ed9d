#!/bin/bash
# round 2, GPU call AH (1 GPU): column kernel A/B -- resident CTAs the register allocation aims at (4 / 5 / 6) and the operator block of the next entry
# prefetched with its source; the cfg5 stage and three single sweeps per variant
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
V=adaptive-multiresolution-dg_b200/csrc/build/variants
for v in m4_b0 m5_b0 m5_b1 m4_b1 m6_b0 m4_b0; do
  AMDG_LIB=$PWD/$V/libamdg_$v.so python bench.py --no-cpu --no-secondary --steps 10 > $O/r02ah_bench_$v.json 2>$O/r02ah_err_$v.log
  python -c "
import json
d=json.loads([l for l in open('$O/r02ah_bench_$v.json') if l.startswith('{')][-1]); r=d['roofline']; print('$v stage ms %.3f'%d['ms_per_step'], 'parity', d['config']['parity_rel_l2'], 'roofline frac %.4f'%r['frac'], 'pass ms', r.get('pass_ms'))
"
  AMDG_LIB=$PWD/$V/libamdg_$v.so python tools/sweep_time.py --workload cfg5 --kernel 0 --lus 0,2 --dims 0,3 --shapes "b>a" 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('   $v', {k:d[k] for k in d if k in ('t','lu','us','frac','kf','kt')})
"
done
timeout 20 nvidia-smi --query-gpu=name,memory.used --format=csv,noheader
