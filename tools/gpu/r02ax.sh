#!/bin/bash
# round 2, GPU call AX (1 GPU): the C++ mirror example incl. the DiffusionAlpt check
cd "$GRAFT_REPO_ROOT"
timeout 300 python -m pytest tests -x -q -m gpu -k "cpp_host_mirror" 2>&1 | tail -n 12
