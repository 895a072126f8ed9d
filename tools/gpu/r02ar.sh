#!/bin/bash
# round 2, GPU call AR (1 GPU): GPU suite after the destructor fix
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
( time timeout 1200 python -m pytest tests -x -q -m gpu > $O/r02ar_pytest.log 2>&1 ) 2>&1 | grep real; tail -n 4 $O/r02ar_pytest.log
timeout 20 nvidia-smi --query-gpu=name,memory.used --format=csv,noheader
