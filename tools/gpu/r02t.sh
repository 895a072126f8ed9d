#!/bin/bash
# round 2, GPU call T: column kernel with two source buffers in turn (no register copies) against the committed form
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
PKG=$PWD/adaptive-multiresolution-dg_b200
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_stage.py -x -q -m gpu -k "8 or 0" > $O/r02t_pytest.log 2>&1
tail -3 $O/r02t_pytest.log
ST=tools/sweep_time.py
: > $O/r02t_sweeps.jsonl
for lib in "" _v1; do
  AMDG_LIB=$PKG/libamdg_b200$lib.so timeout 300 python $ST --workload cfg5 --kernel 8 --lus 2 --dims 0,2,4 --tag new$lib >> $O/r02t_sweeps.jsonl 2>>$O/r02t_err.log
  AMDG_LIB=$PKG/libamdg_b200$lib.so timeout 300 python $ST --workload cfg5 --kernel 8 --lus 1 --acc 1 --dims 0,2 --shapes "b>a,a>b" --tag new$lib >> $O/r02t_sweeps.jsonl 2>>$O/r02t_err.log
  AMDG_LIB=$PKG/libamdg_b200$lib.so timeout 300 python $ST --workload cfg2 --kernel 8 --lus 1,2 --dims 1 --tag new$lib >> $O/r02t_sweeps.jsonl 2>>$O/r02t_err.log
  AMDG_LIB=$PKG/libamdg_b200$lib.so python bench.py --no-cpu --no-secondary --steps 10 --kernel 0 > $O/r02t_bench$lib.json 2>>$O/r02t_err.log
done
python - <<'PY'
import json,collections
T=collections.defaultdict(dict)
for l in open('gpurun_out/r02t_sweeps.jsonl'):
    d=json.loads(l); T[(d['workload'],d['shape'],d['t'],d['lu'],d['acc'])][d['tag']]=d['us']
for k,v in sorted(T.items()):
    print(k, '  '.join('%s:%.1f'%(tag,us) for tag,us in sorted(v.items(), key=lambda x:x[1])))
for lib in ('','_v1'):
    d=json.load(open('gpurun_out/r02t_bench%s.json'%lib)); print('lib', lib or 'new', 'stage ms', d['ms_per_step'], 'roof', d['roofline']['frac'])
PY
grep -v "^frame" $O/r02t_err.log | tail -5
