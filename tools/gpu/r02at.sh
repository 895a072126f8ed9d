#!/bin/bash
# round 2, GPU call AT (1 GPU): gather kernel with grouped dependent loads -- GPU suite, per-call device time on a tiny grid, the live adaptive run
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
( time timeout 1200 python -m pytest tests -x -q -m gpu > $O/r02at_pytest.log 2>&1 ) 2>&1 | grep real; tail -n 3 $O/r02at_pytest.log
python tools/call_overhead.py 2>&1 | head -n 1 | cut -c1-260
examples/live_burgers_adapt -NM 9 -N0 2 -steps 10 -gen 1 2>&1 | tail -n 3 | tee $O/r02at_live_n9.log
