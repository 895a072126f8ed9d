#!/bin/bash
# round 2, GPU call AY (1 GPU): adaptive mode restricted to dimensions with short fibres -- live run (with the longest fibres printed), the adaptive-mode test
cd "$GRAFT_REPO_ROOT"
examples/live_burgers_adapt -NM 9 -N0 2 -steps 10 -gen 1 -trace 1 2>&1 | grep -E "longest|wall per step|averages|LIVE" | tail -n 7
timeout 300 python -m pytest tests -x -q -m gpu -k "adaptive_mode or f4_comp" 2>&1 | tail -n 2
python tools/rebuild_cost.py 2 9 2 3 2>&1 | tail -n 1
