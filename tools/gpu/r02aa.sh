#!/bin/bash
# round 2, GPU call AA (1 GPU): whole GPU suite after the grid_set rework (cached neighbour templates, packed table upload), the f4 compositions,
# cost of a grid change, the live adaptive run
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
( time timeout 1500 python -m pytest tests -x -q -m gpu > $O/r02aa_pytest.log 2>&1 ) 2>&1 | grep real; tail -n 5 $O/r02aa_pytest.log
python tools/rebuild_cost.py 2 9 2 3 2>&1 | tail -n 1
python tools/rebuild_cost.py 4 8 3 3 2>&1 | tail -n 1
python tools/rebuild_cost.py 6 7 1 2 2>&1 | tail -n 1
examples/live_burgers_adapt -NM 9 -N0 2 -steps 10 > $O/r02aa_live_n9.log 2>&1; tail -n 4 $O/r02aa_live_n9.log
