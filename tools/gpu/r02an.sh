#!/bin/bash
# round 2, GPU call AN (1 GPU): run-time knobs of the cfg5 stage at N = 1 (parallel streams per schedule level, units per CTA and heavy-unit threshold of the column kernel)
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
run() { env "$@" python bench.py --no-cpu --no-secondary --steps 10 2>/dev/null | python -c "
import sys, json
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('$*', 'stage ms %.3f'%d['ms_per_step'], 'roof %.4f'%d['roofline']['frac'])
"; }
run AMDG_STAGE_STREAMS=3
run AMDG_STAGE_STREAMS=1
run AMDG_STAGE_STREAMS=2
run AMDG_STAGE_STREAMS=4
run AMDG_STAGE_STREAMS=6
run AMDG_COL_UPC=4
run AMDG_COL_UPC=16
run AMDG_COL_UPC=32
run AMDG_COL_HEAVY=12
run AMDG_COL_HEAVY=32
run AMDG_STAGE_STREAMS=3
