#!/bin/bash
# round 2, GPU call R: column kernel (variant 8) CTA count / heavy threshold on the cfg5 shapes where it beats the lean kernel
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
ST=tools/sweep_time.py
: > $O/r02r_sweeps.jsonl
for ct in 8 16 28; do for hv in 24 32; do
  AMDG_COL_CTAS=$ct AMDG_COL_HEAVY=$hv timeout 200 python $ST --workload cfg5 --kernel 8 --lus 2 --dims 0,3 --shapes "b>a" --tag c${ct}h$hv >> $O/r02r_sweeps.jsonl 2>>$O/r02r_err.log
  AMDG_COL_CTAS=$ct AMDG_COL_HEAVY=$hv timeout 200 python $ST --workload cfg5 --kernel 8 --lus 2 --dims 0 --shapes "b>b" --tag c${ct}h$hv >> $O/r02r_sweeps.jsonl 2>>$O/r02r_err.log
  AMDG_COL_CTAS=$ct AMDG_COL_HEAVY=$hv timeout 200 python $ST --workload cfg5 --kernel 8 --lus 2 --dims 5 --shapes "a>b" --tag c${ct}h$hv >> $O/r02r_sweeps.jsonl 2>>$O/r02r_err.log
done; done
python - <<'PY'
import json,collections
T=collections.defaultdict(dict)
for l in open('gpurun_out/r02r_sweeps.jsonl'):
    d=json.loads(l); T[(d['workload'],d['shape'],d['t'],d['lu'],d['acc'])][d['tag']]=d['us']
for k,v in T.items():
    print(k, '  '.join('%s:%.1f'%(tag,us) for tag,us in sorted(v.items(), key=lambda x:x[1])))
PY
grep -v "^frame" $O/r02r_err.log | tail -5
