#!/bin/bash
# round 2, GPU call AP (1 GPU): host cost per C-ABI call on a tiny grid
cd "$GRAFT_REPO_ROOT"
python tools/call_overhead.py 2>&1 | tail -n 4
