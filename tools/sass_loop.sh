#!/bin/bash
# SASS of the probe instantiations of the column kernel: register counts and opcode histogram of the hottest loop bodies
cd /root/repo/adaptive-multiresolution-dg_b200/csrc
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -DAMDG_COL_PROBE -Xptxas -v -cubin -o /tmp/col_probe.cubin kernels_col.cu 2>&1 | grep -E "registers|spill" 
cuobjdump -sass /tmp/col_probe.cubin > /tmp/col_probe.sass
grep -c . /tmp/col_probe.sass
