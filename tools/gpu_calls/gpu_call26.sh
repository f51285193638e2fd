#!/bin/bash
# round 2, GPU call 26: ncu --set full of the largest non-PatchMatch kernels of the finest level (last launches of a pair)
set -u
mkdir -p gpurun_out
NCT_WLS_LOOP=0 NCT_NL_GRAPH=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"nl_spmv_kernel|nl_update_kernel|nl_pupdate_kernel|bds_feature_error_kernel|knn_grid_kernel|reconstruct_bds_kernel" -s 1290 -c 24 -o gpurun_out/r2_misc_full python tools/one_pair.py 700 1 > gpurun_out/c26.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/c26.log
