#!/bin/bash
# round 2, GPU call 29: compute-sanitizer over the new kernels (small shapes)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "conv2d_tcgen05 and bf16x3 and (case0- or case2- or case5- or case8- or case10- or case13-)" > $O/r2c29_memcheck_conv.log 2>&1; echo "memcheck conv rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|Error" $O/r2c29_memcheck_conv.log | tail -6
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "fused_with_the_upsample" > $O/r2c29_memcheck_loss.log 2>&1; echo "memcheck loss rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|Error" $O/r2c29_memcheck_loss.log | tail -6
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_network.py -m gpu -q -x -k "train_step_small or forward_pair_equals" > $O/r2c29_memcheck_net.log 2>&1; echo "memcheck net rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|Error" $O/r2c29_memcheck_net.log | tail -6
timeout 400 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "conv2d_tcgen05 and bf16x3 and case13-" > $O/r2c29_racecheck_conv.log 2>&1; echo "racecheck conv rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed|hazard|Error" $O/r2c29_racecheck_conv.log | tail -6
