#!/bin/bash
# round 2, GPU call Y (1 GPU): full runs on static grids (convergence table), moments, whole suite
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
for n in 3 4 5 6 7; do
  examples/live_burgers_adapt -NM $n -N0 $n -static 1 -tf 0.05 > $O/r02y_conv_n$n.log 2>&1
  echo "N=$n: $(grep -c '^step' $O/r02y_conv_n$n.log) steps; $(grep 'error vs exact' $O/r02y_conv_n$n.log); $(grep 'LIVE' $O/r02y_conv_n$n.log)"
done
( time timeout 1500 python -m pytest tests -x -q -m gpu > $O/r02y_pytest.log 2>&1 ) 2>&1 | grep real
tail -4 $O/r02y_pytest.log
