#!/bin/bash
# round 2, GPU call 30: CTA pairs for the 128-channel layer: tests + A/B
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k conv2d > $O/r2c30_ops.log 2>&1; echo "ops rc=$?"; grep -E "passed|failed|FAILED|Error|assert" $O/r2c30_ops.log | tail -4
timeout 900 python -m pytest tests/test_gpu_network.py -m gpu -q > $O/r2c30_net.log 2>&1; echo "net rc=$?"; grep -E "passed|failed|FAILED|Error" $O/r2c30_net.log | tail -4
for rep in 1 2 3; do
for v in "pair128:DDN_TC_PAIR128=1" "single128:DDN_TC_PAIR128=0"; do
  name=${v%%:*}; envs=${v#*:}
  env $envs timeout 300 python bench.py --quick --steps 20 > $O/r2c30_ab_${name}_$rep.json 2> $O/r2c30_ab_${name}_$rep.err
  python - "$name" "$rep" "$O/r2c30_ab_${name}_$rep.json" <<'PY'
import json,sys
d=json.load(open(sys.argv[3])); c=d["roofline"]["classes"]
print(sys.argv[1], sys.argv[2], round(d["value"],1), "e2e", round(d["e2e"]["value"],1), d["clocks"]["sm_mhz"], {k[5:-3]:round(v["ms"]/d["steps"],2) for k,v in c.items() if k.startswith("conv")})
PY
done
done
