#!/bin/bash
# round 2, GPU call 32: ncu --set full of the finest-level launch of the BDS error, k-NN and non-local matvec kernels
# (call 28 captured 60 launches with sources and overflowed the 64 MiB return limit; here: one launch each, CSV only)
set -u
mkdir -p gpurun_out
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed,dram__throughput.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size,launch__block_size,launch__occupancy_limit_registers,launch__occupancy_limit_shared_mem,smsp__average_warp_latency_issue_stalled_long_scoreboard.pct,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct"
run() {  # name regex skip
  timeout 300 ncu --metrics "$M" --clock-control none -k regex:"$2" --launch-skip "$3" -c 1 --csv --log-file gpurun_out/c32_$1.csv python tools/one_pair.py 700 1 > gpurun_out/c32_$1.log 2>&1
  echo "$1 rc=$?"
}
run bds_err "^bds_feature_error_kernel" 4
run knn_grid "^knn_grid_kernel" 4
run reconstruct "^reconstruct_bds_kernel" 4
run nl_spmv "^nl_spmv_kernel" 400
run nl_update "^nl_update_kernel" 400
ls -la gpurun_out/c32_*
