#!/bin/bash
# round 2, GPU call 8: BatchNorm-backward column sums in the data-gradient epilogue, programmatic dependent launch, per-D loss kernels
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q -x > $O/r2c8_ops.log 2>&1; echo "ops rc=$?"; grep -E "passed|failed|FAILED|Error" $O/r2c8_ops.log | tail -8
timeout 900 python -m pytest tests/test_gpu_network.py -m gpu -q -s > $O/r2c8_net.log 2>&1; echo "net rc=$?"; grep -E "rel err|parity|passed|failed|Error|error|FAILED" $O/r2c8_net.log | tail -30
for v in "default:" "nopdl:DDN_PDL=0" "nofuse:DDN_FUSE_BWD_STATS_MINC=9999" "fuse64:DDN_FUSE_BWD_STATS_MINC=64" "neither:DDN_PDL=0 DDN_FUSE_BWD_STATS_MINC=9999"; do
  name=${v%%:*}; envs=${v#*:}
  env $envs timeout 300 python bench.py --quick --steps 10 > $O/r2c8_bench_$name.json 2> $O/r2c8_bench_$name.err; echo "bench $name rc=$?"
  cut -c1-190 $O/r2c8_bench_$name.json
done
LOSS_SWEEP=1000,5000,50000 timeout 300 python scripts/bench_loss.py > $O/r2c8_loss_sweep.json 2> $O/r2c8_loss_sweep.err; echo "loss sweep rc=$?"; tail -n 9 $O/r2c8_loss_sweep.err | cut -c1-420
NCU="ncu --set full --clock-control none --import-source on"
LOSS_SWEEP=5000 LOSS_DIMS=16 timeout 300 $NCU -k regex:loss_lowres -s 6 -c 2 -o $O/r2_prof_loss_c3 -f python scripts/bench_loss.py > $O/r2c8_ncu5.log 2>&1; echo "ncu loss rc=$?"
