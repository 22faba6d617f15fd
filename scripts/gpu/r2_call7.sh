#!/bin/bash
# round 2, GPU call 7: loss fused with the upsample (tests, C3 sweep, A/B), full test suite, ncu --set full captures of the largest launches
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q > $O/r2c7_ops.log 2>&1; echo "ops rc=$?"; grep -E "passed|failed|FAILED|Error" $O/r2c7_ops.log | tail -8
timeout 900 python -m pytest tests/test_gpu_network.py -m gpu -q -s > $O/r2c7_net.log 2>&1; echo "net rc=$?"; grep -E "rel err|parity|passed|failed|Error|error|FAILED" $O/r2c7_net.log | tail -24
LOSS_SWEEP=1000,5000,50000 timeout 300 python scripts/bench_loss.py > $O/r2c7_loss_sweep.json 2> $O/r2c7_loss_sweep.err; echo "loss sweep rc=$?"; tail -n 9 $O/r2c7_loss_sweep.err | cut -c1-420
timeout 300 python bench.py --quick --steps 10 > $O/r2c7_bench.json 2> $O/r2c7_bench.err; echo "bench rc=$?"
DDN_FUSED_UPSAMPLE_LOSS=0 timeout 300 python bench.py --quick --steps 10 > $O/r2c7_bench_generic_loss.json 2> $O/r2c7_bench_generic_loss.err; echo "bench generic rc=$?"
cut -c1-200 $O/r2c7_bench.json $O/r2c7_bench_generic_loss.json
NCU="ncu --set full --clock-control none --import-source on"
timeout 300 $NCU -k regex:conv_tc_kernel -s 103 -c 1 -o $O/r2_prof_conv_fwd_l4 -f python bench.py --profile-run --steps 1 > $O/r2c7_ncu1.log 2>&1; echo "ncu fwd rc=$?"
timeout 300 $NCU -k regex:conv_tc_kernel -s 107 -c 1 -o $O/r2_prof_conv_dgrad_l4 -f python bench.py --profile-run --steps 1 > $O/r2c7_ncu2.log 2>&1; echo "ncu dgrad rc=$?"
timeout 300 $NCU -k regex:wgrad_tc_kernel -s 36 -c 1 -o $O/r2_prof_wgrad_l4 -f python bench.py --profile-run --steps 1 > $O/r2c7_ncu3.log 2>&1; echo "ncu wgrad rc=$?"
timeout 300 $NCU -k regex:bn_bwd_apply_kernel -s 36 -c 2 -o $O/r2_prof_bn_bwd_apply -f python bench.py --profile-run --steps 1 > $O/r2c7_ncu4.log 2>&1; echo "ncu bn rc=$?"
LOSS_SWEEP=5000 LOSS_DIMS=16 timeout 300 $NCU -k regex:loss_lowres -s 6 -c 2 -o $O/r2_prof_loss_c3 -f python scripts/bench_loss.py > $O/r2c7_ncu5.log 2>&1; echo "ncu loss rc=$?"
ls -la $O/*.ncu-rep
