#!/bin/bash
# round 2, GPU call AS (1 GPU): issue vs drain time of the three RK3 stages in the live adaptive run
cd "$GRAFT_REPO_ROOT"
examples/live_burgers_adapt -NM 9 -N0 2 -steps 10 -gen 1 2>&1 | tail -n 3
