#!/bin/bash
# round 2, GPU call AC (1 GPU): A/B of the packed grid-table upload in the live adaptive run
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
for i in 1 2; do
examples/live_burgers_adapt -NM 9 -N0 2 -steps 10 2>&1 | grep "wall per step" | sed "s/^/[packed] /"
AMDG_GRID_UPLOAD_SPLIT=1 examples/live_burgers_adapt -NM 9 -N0 2 -steps 10 2>&1 | grep "wall per step" | sed "s/^/[split ] /"
done
AMDG_VERBOSE=1 examples/live_burgers_adapt -NM 9 -N0 2 -steps 3 > $O/r02ac_verbose.log 2>&1
grep -c . $O/r02ac_verbose.log
timeout 20 nvidia-smi --query-gpu=name,memory.used --format=csv,noheader
