#!/bin/bash
# round 2, GPU call J: column kernel v2 (64-bit per-lane pointers, deeper prefetch, heavy units per 32 columns): parity, then the depth x occupancy x NC matrix
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
PKG=$PWD/adaptive-multiresolution-dg_b200
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_stage.py -x -q -m gpu -k "8" > $O/r02j_pytest.log 2>&1
tail -4 $O/r02j_pytest.log
ST=tools/sweep_time.py
: > $O/r02j_sweeps.jsonl
for v in d1b5 d1b6 d2b4 d2b5 d2b6 d3b4; do
  for nc in 1 2; do
    AMDG_LIB=$PKG/libamdg_b200_$v.so AMDG_COL_NC=$nc python $ST --workload cfg2 --kernel 8 --lus 1,2 --dims 1 --tag ${v}_nc$nc >> $O/r02j_sweeps.jsonl 2>>$O/r02j_err.log
  done
  for nc in 1 2 4; do
    AMDG_LIB=$PKG/libamdg_b200_$v.so AMDG_COL_NC=$nc python $ST --workload cfg5 --kernel 8 --lus 2 --dims 0,3 --shapes "b>a" --tag ${v}_nc$nc >> $O/r02j_sweeps.jsonl 2>>$O/r02j_err.log
  done
done
python - <<'PY'
import json
for l in open('gpurun_out/r02j_sweeps.jsonl'):
    d=json.loads(l); print("%-12s %s %s t=%d %s  %7.2f us  frac %.3f"%(d['tag'],d['workload'],d['shape'],d['t'],d['lu'],d['us'],d['frac']))
PY
grep -v "^frame" $O/r02j_err.log | tail -5
