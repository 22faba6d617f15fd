#!/bin/bash
# round 2, GPU call 31: final tree: full GPU test suite, smoke, default bench
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > $O/r2c31_tests.log 2>&1; echo "tests rc=$?"; grep -E "passed|failed|FAILED|Error" $O/r2c31_tests.log | tail -4
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/r2c31_smoke.log 2>&1; echo "smoke rc=$?"; grep -E "smoke" $O/r2c31_smoke.log | cut -c1-220
timeout 600 python bench.py > $O/r2c31_bench.json 2> $O/r2c31_bench.err; echo "bench rc=$?"
python -c "
import json;d=json.load(open('$O/r2c31_bench.json'));r=d['roofline'];print(round(d['value'],1),'e2e',round(d['e2e']['value'],1),'ms',round(d['ms_per_step'],2),d['clocks'],'issued',round(r['issued_frac'],3),'share',round(r['conv_share_of_step'],3),'launches',d['launches_per_step'])"
