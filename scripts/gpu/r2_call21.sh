#!/bin/bash
# round 2, GPU call 21: the evidence run of the final build: full GPU test suite, bench (full legs), c5 shard, loss sweep,
# ncu launch list and one `--set full` capture of the largest launch of every kernel class
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --ignore=tests/test_gpu_multi.py > $O/r2c21_tests.log 2>&1; echo "tests rc=$?"; grep -E "passed|failed|FAILED|Error" $O/r2c21_tests.log | tail -5
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/r2c21_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 2 $O/r2c21_smoke.log | cut -c1-200
LOSS_SWEEP=1000,5000,50000 timeout 300 python scripts/bench_loss.py > $O/r2c21_loss_sweep.json 2> $O/r2c21_loss_sweep.err; echo "loss sweep rc=$?"; grep "'D': 16, 'non_matches_per_image': 5000" $O/r2c21_loss_sweep.err | cut -c1-420
timeout 900 python bench.py > $O/r2c21_bench_1gpu.json 2> $O/r2c21_bench_1gpu.err; echo "bench rc=$?"; cut -c1-260 $O/r2c21_bench_1gpu.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/r2c21_bench_ref.json 2> $O/r2c21_bench_ref.err; echo "ref rc=$?"; cut -c1-300 $O/r2c21_bench_ref.json
timeout 600 python bench.py --config c5 --quick > $O/r2c21_bench_c5.json 2> $O/r2c21_bench_c5.err; echo "c5 rc=$?"; cut -c1-260 $O/r2c21_bench_c5.json
timeout 600 python bench.py --config c5 --quick --l2-pixel-loss > $O/r2c21_bench_c5_l2.json 2> $O/r2c21_bench_c5_l2.err; echo "c5 l2 rc=$?"; cut -c1-160 $O/r2c21_bench_c5_l2.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r2_final_launches.csv python bench.py --profile-run --steps 1 > $O/r2c21_ncu_bench.log 2>&1; echo "ncu list rc=$?"
NCU="ncu --set full --clock-control none --import-source on"
# launch indices: one warm-up step precedes the captured one (profile-run --steps 1): skip the first step's launches of each kernel
timeout 300 $NCU -k regex:conv_tc_kernel -s 85 -c 1 -o $O/r2_prof_conv_fwd_l4 -f python bench.py --profile-run --steps 1 > $O/r2c21_ncu1.log 2>&1; echo "ncu fwd rc=$?"
timeout 300 $NCU -k regex:conv_tc_kernel -s 90 -c 1 -o $O/r2_prof_conv_dgrad_l4 -f python bench.py --profile-run --steps 1 > $O/r2c21_ncu2.log 2>&1; echo "ncu dgrad rc=$?"
timeout 300 $NCU -k regex:^wgrad_tc_kernel -s 30 -c 1 -o $O/r2_prof_wgrad_l4 -f python bench.py --profile-run --steps 1 > $O/r2c21_ncu3.log 2>&1; echo "ncu wgrad rc=$?"
timeout 300 $NCU -k regex:bn_bwd_apply_kernel -s 36 -c 2 -o $O/r2_prof_bn_bwd_apply -f python bench.py --profile-run --steps 1 > $O/r2c21_ncu4.log 2>&1; echo "ncu bn rc=$?"
timeout 300 $NCU -k regex:conv64_halo_kernel -s 12 -c 1 -o $O/r2_prof_conv64_halo -f python bench.py --profile-run --steps 1 > $O/r2c21_ncu5.log 2>&1; echo "ncu halo rc=$?"
timeout 300 $NCU -k regex:wgrad64_halo_kernel -s 6 -c 1 -o $O/r2_prof_wgrad64_halo -f python bench.py --profile-run --steps 1 > $O/r2c21_ncu6.log 2>&1; echo "ncu wgrad halo rc=$?"
LOSS_SWEEP=5000 LOSS_DIMS=16 timeout 300 $NCU -k regex:loss_lowres -s 6 -c 2 -o $O/r2_prof_loss_c3 -f python scripts/bench_loss.py > $O/r2c21_ncu7.log 2>&1; echo "ncu loss rc=$?"
ls -la $O/r2_prof_*.ncu-rep
