#!/bin/bash
# round 2, GPU call A: parity of the register-direct kernel + first timings against the lean kernel
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
PKG=adaptive-multiresolution-dg_b200
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/r02a_smi.txt
timeout 1500 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "kernel_variants or live or batch or random or full_size" > $O/r02a_pytest.log 2>&1
tail -15 $O/r02a_pytest.log
ST=tools/sweep_time.py
python $ST --workload cfg2 --kernel 5 --tag lean > $O/r02a_sweeps.jsonl 2>$O/r02a_err.log
python $ST --workload cfg2 --kernel 6 --tag dir_c5 >> $O/r02a_sweeps.jsonl 2>>$O/r02a_err.log
AMDG_LIB=$PWD/$PKG/libamdg_b200_c4.so python $ST --workload cfg2 --kernel 6 --tag dir_c4 >> $O/r02a_sweeps.jsonl 2>>$O/r02a_err.log
AMDG_LIB=$PWD/$PKG/libamdg_b200_c6.so python $ST --workload cfg2 --kernel 6 --tag dir_c6 >> $O/r02a_sweeps.jsonl 2>>$O/r02a_err.log
for c in 64 320; do AMDG_DIR_COST=$c python $ST --workload cfg2 --kernel 6 --lus 2 --tag dir_c5_cost$c >> $O/r02a_sweeps.jsonl 2>>$O/r02a_err.log; done
AMDG_TC_PDL=0 python $ST --workload cfg2 --kernel 6 --lus 2 --tag dir_c5_nopdl >> $O/r02a_sweeps.jsonl 2>>$O/r02a_err.log
for mm in 1 2 4 8; do AMDG_DIR_MAXM=$mm AMDG_DIR_MINM=$mm python $ST --workload cfg2 --kernel 6 --lus 2 --dims 0,3 --tag dir_m$mm >> $O/r02a_sweeps.jsonl 2>>$O/r02a_err.log; done
AMDG_DIR_MINM=16 python $ST --workload cfg2 --kernel 6 --lus 2 --dims 0,3 --tag dir_m16plus >> $O/r02a_sweeps.jsonl 2>>$O/r02a_err.log
python $ST --workload cfg2 --kernel 6 --acc 1 --lus 1 --tag dir_acc >> $O/r02a_sweeps.jsonl 2>>$O/r02a_err.log
python $ST --workload cfg2 --kernel 5 --acc 1 --lus 1 --tag lean_acc >> $O/r02a_sweeps.jsonl 2>>$O/r02a_err.log
python $ST --workload cfg5 --kernel 5 --lus 2 --dims 0,3,5 --tag lean >> $O/r02a_sweeps.jsonl 2>>$O/r02a_err.log
python $ST --workload cfg5 --kernel 6 --lus 2 --dims 0,3,5 --tag dir_c5 >> $O/r02a_sweeps.jsonl 2>>$O/r02a_err.log
AMDG_LIB=$PWD/$PKG/libamdg_b200_c4.so python $ST --workload cfg5 --kernel 6 --lus 2 --dims 0,3,5 --tag dir_c4 >> $O/r02a_sweeps.jsonl 2>>$O/r02a_err.log
for k in 5 6; do
  python bench.py --workload cfg2 --kernel $k --no-cpu > $O/r02a_bench_cfg2_k$k.json 2>>$O/r02a_err.log
  python bench.py --workload cfg5 --kernel $k --no-cpu --steps 5 > $O/r02a_bench_cfg5_k$k.json 2>>$O/r02a_err.log
done
ncu --set full --clock-control none --import-source on -k regex:sweep_dir -s 8 -c 2 -o $O/r02a_dir_full python $ST --workload cfg2 --kernel 6 --lus 2 --dims 0 > $O/r02a_ncu.log 2>&1
cat $O/r02a_sweeps.jsonl | head -100
tail -5 $O/r02a_err.log
