#!/bin/bash
# round 2, GPU call 28: ncu --set full of the BDS / k-NN / k-means kernels of one pair
set -u
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"bds_feature_error_kernel|knn_grid_kernel|reconstruct_bds_kernel|km_dist_kernel|inv_|grid_count|grid_fill|cell_masks" -c 60 -o gpurun_out/r2_bds_knn_full python tools/one_pair.py 700 1 > gpurun_out/c28.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/c28.log
