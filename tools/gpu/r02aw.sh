#!/bin/bash
# round 2, GPU call AW (1 GPU): the mirror's linear-operator classes (HyperbolicAlpt, RK3SSP::step_rk) in the advection convergence run
cd "$GRAFT_REPO_ROOT"
examples/live_advection_convergence -Nmin 3 -Nmax 6 2>&1 | tail -n 9
timeout 600 python -m pytest tests -x -q -m gpu -k "live_reference_advection or cpp_host_mirror or live_reference_dropin" 2>&1 | tail -n 2
