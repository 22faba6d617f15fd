#!/bin/bash
# round 2, GPU call 32: loss forward with 8 vs 4 index pairs per lane group: tests + sweep
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "loss or ragged or triplet" > $O/r2c32_ops.log 2>&1; echo "ops rc=$?"; grep -E "passed|failed|FAILED|Error" $O/r2c32_ops.log | tail -3
for it in 8 4 8 4; do
DDN_LOSS_FWD_ITEMS=$it LOSS_SWEEP=5000,50000 LOSS_DIMS=8,16 timeout 300 python scripts/bench_loss.py > $O/r2c32_sweep_$it.json 2> $O/r2c32_sweep_$it.err
echo "items=$it"; grep "fused" $O/r2c32_sweep_$it.err | python -c "
import sys,ast
for l in sys.stdin:
    d=ast.literal_eval(l.strip()); print('  D=%d nm=%d fwd %.1f us (%.0f%%) bwd %.1f us (%.0f%%)'%(d['D'],d['non_matches_per_image'],d['fused']['fwd_us'],100*d['fused']['fwd_frac_of_hbm_peak'],d['fused']['bwd_us'],100*d['fused']['bwd_frac_of_hbm_peak']))"
done
