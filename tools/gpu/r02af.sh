#!/bin/bash
# round 2, GPU call AF (1 GPU): which sweep kernel for the small adaptive grids of the live Burgers run (work lists are rebuilt after every grid change)
cd "$GRAFT_REPO_ROOT"
for k in 0 1 0 1 4 8; do
AMDG_KERNEL=$k examples/live_burgers_adapt -NM 9 -N0 2 -steps 10 -gen 1 2>&1 | grep "wall per step\|LIVE" | tr '\n' ' ' | sed "s/^/[kernel $k] /"; echo
done
timeout 20 nvidia-smi --query-gpu=name,memory.used --format=csv,noheader
