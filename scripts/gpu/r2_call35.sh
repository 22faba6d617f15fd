#!/bin/bash
# round 2, GPU call 35: the tree as committed: full GPU test suite + smoke
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2c35_tests.log 2>&1; echo "tests rc=$?"; tail -n 2 gpurun_out/r2c35_tests.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -n 2 | cut -c1-200
timeout 300 python bench.py --quick > gpurun_out/r2c35_bench.json 2>/dev/null; cut -c1-140 gpurun_out/r2c35_bench.json
