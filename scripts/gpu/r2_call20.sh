#!/bin/bash
# round 2, GPU call 20: quad-lane loss kernels: tests + sweep
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q > $O/r2c20_ops.log 2>&1; echo "ops rc=$?"; grep -E "passed|failed|FAILED|Error|assert" $O/r2c20_ops.log | tail -8
LOSS_SWEEP=1000,5000,50000 timeout 300 python scripts/bench_loss.py > $O/r2c20_loss_sweep.json 2> $O/r2c20_loss_sweep.err; echo "loss sweep rc=$?"; tail -n 9 $O/r2c20_loss_sweep.err | cut -c1-330
