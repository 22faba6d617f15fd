#!/bin/bash
# round 2, GPU call 3: network tests incl. the new round-2 parity tests, kernel-level launch list of one step (ncu), quick bench
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_network.py -m gpu -q -s > $O/r2c3_net.log 2>&1; echo "net rc=$?"; grep -E "rel err|parity|passed|failed|Error|error|FAILED" $O/r2c3_net.log | tail -30
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r2c3_launches.csv python bench.py --profile-run --steps 1 > $O/r2c3_ncu_bench.log 2>&1; echo "ncu rc=$?"
timeout 300 python bench.py --quick --steps 10 > $O/r2c3_bench.json 2> $O/r2c3_bench.err; echo "bench rc=$?"
tail -n 3 $O/r2c3_bench.err; cut -c1-400 $O/r2c3_bench.json
