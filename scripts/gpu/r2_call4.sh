#!/bin/bash
# round 2, GPU call 4: occupancy-sized BN grids, fixed well-conditioned gradient test, smoke, full default bench
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -x -q > $O/r2c4_ops.log 2>&1; echo "ops rc=$?"; tail -2 $O/r2c4_ops.log
timeout 900 python -m pytest tests/test_gpu_network.py -m gpu -q -s > $O/r2c4_net.log 2>&1; echo "net rc=$?"; grep -E "rel err|parity|passed|failed|Error|error|FAILED" $O/r2c4_net.log | tail -30
timeout 300 python __graft_entry__.py > $O/r2c4_smoke.log 2>&1; echo "smoke rc=$?"; grep smoke $O/r2c4_smoke.log
timeout 600 python bench.py --steps 10 > $O/r2c4_bench.json 2> $O/r2c4_bench.err; echo "bench rc=$?"
tail -n 3 $O/r2c4_bench.err; cut -c1-300 $O/r2c4_bench.json
timeout 300 python bench.py --quick --steps 10 --two-calls > $O/r2c4_bench_two.json 2> $O/r2c4_bench_two.err; echo "bench two rc=$?"
