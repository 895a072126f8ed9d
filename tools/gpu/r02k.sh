#!/bin/bash
# round 2, GPU call K: column kernel v3 (software pipeline across units, L1 prefetch of sources and operator blocks): parity, timings
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
PKG=$PWD/adaptive-multiresolution-dg_b200
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_stage.py -x -q -m gpu -k "8" > $O/r02k_pytest.log 2>&1
tail -4 $O/r02k_pytest.log
ST=tools/sweep_time.py
: > $O/r02k_sweeps.jsonl
python $ST --workload cfg2 --kernel 8 --tag k8v3 >> $O/r02k_sweeps.jsonl 2>>$O/r02k_err.log
python $ST --workload cfg5 --kernel 8 --lus 2 --dims 0,3,5 --tag k8v3 >> $O/r02k_sweeps.jsonl 2>>$O/r02k_err.log
for v in d1b5 d1b6 d2b4; do
  for nc in 1 2; do
    AMDG_LIB=$PKG/libamdg_b200_$v.so AMDG_COL_NC=$nc python $ST --workload cfg2 --kernel 8 --lus 1,2 --dims 1 --tag ${v}_nc$nc >> $O/r02k_sweeps.jsonl 2>>$O/r02k_err.log
  done
  for nc in 1 2 4; do
    AMDG_LIB=$PKG/libamdg_b200_$v.so AMDG_COL_NC=$nc python $ST --workload cfg5 --kernel 8 --lus 2 --dims 0,3 --shapes "b>a" --tag ${v}_nc$nc >> $O/r02k_sweeps.jsonl 2>>$O/r02k_err.log
  done
done
for ct in 2 3 6 8; do
  AMDG_COL_CTAS=$ct python $ST --workload cfg2 --kernel 8 --lus 1,2 --dims 1 --tag ctas$ct >> $O/r02k_sweeps.jsonl 2>>$O/r02k_err.log
  AMDG_COL_CTAS=$ct python $ST --workload cfg5 --kernel 8 --lus 2 --dims 0,3 --shapes "b>a" --tag ctas$ct >> $O/r02k_sweeps.jsonl 2>>$O/r02k_err.log
done
for hv in 8 16 32; do
  AMDG_COL_HEAVY=$hv python $ST --workload cfg2 --kernel 8 --lus 0,2 --dims 1 --tag heavy$hv >> $O/r02k_sweeps.jsonl 2>>$O/r02k_err.log
done
python - <<'PY'
import json
for l in open('gpurun_out/r02k_sweeps.jsonl'):
    d=json.loads(l); print("%-12s %s %s t=%d %s  %7.2f us  frac %.3f"%(d['tag'],d['workload'],d['shape'],d['t'],d['lu'],d['us'],d['frac']))
PY
ncu --set full --clock-control none --import-source on -k regex:sweep_col -s 8 -c 1 -o /tmp/k_cfg2 -f python $ST --workload cfg2 --kernel 8 --lus 1 --dims 1 > $O/r02k_ncu1.log 2>&1
ncu -i /tmp/k_cfg2.ncu-rep --page details > $O/r02k_cfg2_details.txt
ncu -i /tmp/k_cfg2.ncu-rep --page source --csv > $O/r02k_cfg2_source.csv 2>/dev/null
ncu --set full --clock-control none --import-source on -k regex:sweep_col -s 8 -c 1 -o /tmp/k_cfg5 -f python $ST --workload cfg5 --kernel 8 --lus 2 --dims 0 --shapes "b>a" > $O/r02k_ncu2.log 2>&1
ncu -i /tmp/k_cfg5.ncu-rep --page details > $O/r02k_cfg5_details.txt
ncu -i /tmp/k_cfg5.ncu-rep --page source --csv > $O/r02k_cfg5_source.csv 2>/dev/null
grep -v "^frame" $O/r02k_err.log | tail -5
