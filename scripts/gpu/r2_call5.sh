#!/bin/bash
# round 2, GPU call 5: fixed tests (ops incl. ragged / triplet / masked match; network), ncu launch list, two-stream experiment
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q > $O/r2c5_ops.log 2>&1; echo "ops rc=$?"; grep -E "passed|failed|FAILED|Error" $O/r2c5_ops.log | tail -8
timeout 900 python -m pytest tests/test_gpu_network.py -m gpu -q -s -k "well_conditioned or stale or contract" > $O/r2c5_net.log 2>&1; echo "net rc=$?"; grep -E "rel err|passed|failed|Error|error|FAILED" $O/r2c5_net.log | tail -30
timeout 300 python bench.py --quick --steps 10 --two-streams > $O/r2c5_bench_2s.json 2> $O/r2c5_bench_2s.err; echo "bench 2s rc=$?"; tail -n 2 $O/r2c5_bench_2s.err
timeout 300 python bench.py --quick --steps 10 > $O/r2c5_bench.json 2> $O/r2c5_bench.err; echo "bench rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r2c5_launches.csv python bench.py --profile-run --steps 1 > $O/r2c5_ncu_bench.log 2>&1; echo "ncu rc=$?"
cut -c1-200 $O/r2c5_bench_2s.json $O/r2c5_bench.json
