#!/bin/bash
# round 2, GPU call 2: the refactored engine (stats via atomics+ticket, planes-only activations, groups, sub-tile conv kernel with
# N-split tail, batched unpack) -- op tests, network tests, A/B benches
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -x -q > $O/r2c2_ops.log 2>&1; rc=$?; echo "ops rc=$rc"; tail -15 $O/r2c2_ops.log
if [ $rc -ne 0 ]; then
  DDN_TC_PAIR=0 DDN_TC_TAIL=0 timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -x -q > $O/r2c2_ops_nopair.log 2>&1; echo "ops nopair notail rc=$?"; tail -8 $O/r2c2_ops_nopair.log
fi
timeout 600 python -m pytest tests/test_gpu_network.py -m gpu -x -q -s > $O/r2c2_net.log 2>&1; rc2=$?; echo "net rc=$rc2"; grep -E "rel err|parity|passed|failed|Error|error" $O/r2c2_net.log | tail -25
if [ $rc -eq 0 ]; then
  timeout 200 python bench.py --no-cpu-baseline --steps 8 > $O/r2c2_bench_pair.json 2> $O/r2c2_bench_pair.err; echo "bench pair rc=$?"
  timeout 200 python bench.py --no-cpu-baseline --steps 8 --two-calls > $O/r2c2_bench_two.json 2> $O/r2c2_bench_two.err; echo "bench two rc=$?"
  DDN_TC_PAIR=0 timeout 200 python bench.py --no-cpu-baseline --steps 8 > $O/r2c2_bench_pair_1cta.json 2> $O/r2c2_bench_pair_1cta.err; echo "bench 1cta rc=$?"
  DDN_TC_TAIL=0 timeout 200 python bench.py --no-cpu-baseline --steps 8 > $O/r2c2_bench_pair_notail.json 2> $O/r2c2_bench_pair_notail.err; echo "bench notail rc=$?"
fi
tail -3 $O/r2c2_bench_*.err
cat $O/r2c2_bench_*.json | cut -c1-300
