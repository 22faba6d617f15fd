#!/bin/bash
# round 2, GPU call 12: halo kernel with unrolled MMA issue: conv tests, launch list
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
export DDN_PDL=0 DDN_FUSE_BWD_STATS_MINC=9999
timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "conv2d" > $O/r2c12_ops.log 2>&1; echo "ops rc=$?"; grep -E "passed|failed|FAILED|Error|assert" $O/r2c12_ops.log | tail -4
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r2c12_launches.csv python bench.py --profile-run --steps 1 > $O/r2c12_ncu_bench.log 2>&1; echo "ncu list rc=$?"
grep -E "halo" $O/r2c12_launches.csv | tail -14 | awk -F'","' '{print $5, $(NF)}' | cut -c1-80
