#!/bin/bash
# round 2, GPU call N: column kernel, prefetch distance (units) x resident CTAs
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
PKG=$PWD/adaptive-multiresolution-dg_b200
ST=tools/sweep_time.py
: > $O/r02n_sweeps.jsonl
for v in p1b4 p2b4 p4b4 p2b3 p4b3 p4b2 p8b2; do
  b=${v: -1}
  AMDG_LIB=$PKG/libamdg_b200_$v.so AMDG_COL_CTAS=$b python $ST --workload cfg2 --kernel 8 --lus 1,2 --dims 1 --tag ${v} >> $O/r02n_sweeps.jsonl 2>>$O/r02n_err.log
  AMDG_LIB=$PKG/libamdg_b200_$v.so AMDG_COL_CTAS=$b python $ST --workload cfg5 --kernel 8 --lus 2 --dims 0,3 --shapes "b>a" --tag ${v} >> $O/r02n_sweeps.jsonl 2>>$O/r02n_err.log
  AMDG_LIB=$PKG/libamdg_b200_$v.so AMDG_COL_CTAS=$b AMDG_COL_NC=2 python $ST --workload cfg5 --kernel 8 --lus 2 --dims 0,3 --shapes "b>a" --tag ${v}_nc2 >> $O/r02n_sweeps.jsonl 2>>$O/r02n_err.log
done
python - <<'PY'
import json,collections
T=collections.defaultdict(dict)
for l in open('gpurun_out/r02n_sweeps.jsonl'):
    d=json.loads(l); T[(d['workload'],d['shape'],d['t'],d['lu'])][d['tag']]=d['us']
for k,v in T.items():
    print(k, '  '.join('%s:%.1f'%(tag,us) for tag,us in sorted(v.items(), key=lambda x:x[1])))
PY
grep -v "^frame" $O/r02n_err.log | tail -5
