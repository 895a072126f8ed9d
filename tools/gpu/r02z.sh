#!/bin/bash
# round 2, GPU call Z (1 GPU): linear advection convergence table (reference arm and device arm), new live tests
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
( time examples/live_advection_convergence -Nmin 3 -Nmax 7 > $O/r02z_advection.log 2>&1 ) 2>&1 | grep real
cat $O/r02z_advection.log
timeout 1500 python -m pytest tests/test_gpu_stage.py -x -q -m gpu -k "live" > $O/r02z_pytest.log 2>&1; tail -4 $O/r02z_pytest.log
