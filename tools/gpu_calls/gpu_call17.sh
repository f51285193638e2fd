#!/bin/bash
# round 2, GPU call 17: ncu --set full of one WLS iteration (plain launches) incl. the shared-memory bottom kernel with source counters
set -u
mkdir -p gpurun_out
NCT_WLS_LOOP=0 NCT_NL_GRAPH=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"mg_|pcg_" -s 600 -c 28 -o gpurun_out/r2_wls_iter_full python tools/one_pair.py 700 1 > gpurun_out/c17_ncu_wls.log 2>&1; echo "ncu wls rc=$?"; tail -2 gpurun_out/c17_ncu_wls.log
