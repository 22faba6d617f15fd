#!/bin/bash
# round 2, GPU call 36: 32-bit pixel division in the fused loss kernels: loss tests + C3 timings
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
timeout 200 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "loss or ragged or triplet" > $O/r2c36_ops.log 2>&1; echo "ops rc=$?"; tail -n 1 $O/r2c36_ops.log
LOSS_SWEEP=1000,5000,50000 timeout 200 python scripts/bench_loss.py > $O/r2c36_loss_sweep.json 2> $O/r2c36_loss_sweep.err; echo "sweep rc=$?"
grep "fused" $O/r2c36_loss_sweep.err | python -c "
import sys,ast
for l in sys.stdin:
    d=ast.literal_eval(l.strip()); print('  D=%d nm=%d fwd %.1f us (%.0f%%) bwd %.1f us (%.0f%%)'%(d['D'],d['non_matches_per_image'],d['fused']['fwd_us'],100*d['fused']['fwd_frac_of_hbm_peak'],d['fused']['bwd_us'],100*d['fused']['bwd_frac_of_hbm_peak']))"
