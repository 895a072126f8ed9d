#!/bin/bash
# round 2, GPU call AV (1 GPU): final GPU suite + smoke on the final tree
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
( time timeout 1200 python -m pytest tests -x -q -m gpu > $O/r02av_pytest.log 2>&1 ) 2>&1 | grep real; tail -n 3 $O/r02av_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 1
