#!/bin/bash
# round 2, GPU call 1: first hardware run of conv_tc_pair_kernel (cta_group::2) + A/B bench, every step under its own timeout
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > $O/r2c1_gpu.txt 2>&1
timeout 300 python bench.py --no-cpu-baseline --steps 5 > $O/r2c1_bench_base.json 2> $O/r2c1_bench_base.err; echo "base rc=$?"
DDN_TC_2CTA=1 timeout 240 python -m pytest tests/test_gpu_ops.py -m gpu -x -q > $O/r2c1_pair1_ops.log 2>&1; rc=$?; echo "pair1 ops rc=$rc"; tail -5 $O/r2c1_pair1_ops.log
if [ $rc -eq 0 ]; then
  DDN_TC_2CTA=1 timeout 300 python -m pytest tests/test_gpu_network.py -m gpu -x -q > $O/r2c1_pair1_net.log 2>&1; echo "pair1 net rc=$?"; tail -5 $O/r2c1_pair1_net.log
  DDN_TC_2CTA=1 timeout 200 python bench.py --no-cpu-baseline --steps 5 > $O/r2c1_bench_pair1.json 2> $O/r2c1_bench_pair1.err; echo "pair1 bench rc=$?"
  DDN_TC_2CTA=2 timeout 240 python -m pytest tests/test_gpu_ops.py -m gpu -x -q > $O/r2c1_pair2_ops.log 2>&1; rc2=$?; echo "pair2 ops rc=$rc2"; tail -3 $O/r2c1_pair2_ops.log
  if [ $rc2 -eq 0 ]; then
    DDN_TC_2CTA=2 timeout 200 python bench.py --no-cpu-baseline --steps 5 > $O/r2c1_bench_pair2.json 2> $O/r2c1_bench_pair2.err; echo "pair2 bench rc=$?"
  fi
fi
nvidia-smi > $O/r2c1_after.txt 2>&1
cat $O/r2c1_bench_*.json | cut -c1-400
