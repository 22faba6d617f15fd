#!/bin/bash
# round 2, GPU call 33: stem backward without the intermediate gradient: tests + A/B
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_network.py -m gpu -q -s > $O/r2c33_net.log 2>&1; echo "net rc=$?"; grep -E "well-conditioned|parity|passed|failed|FAILED|Error" $O/r2c33_net.log | tail -16
for rep in 1 2; do
for v in "fused:DDN_STEM_BWD_FUSED=1" "separate:DDN_STEM_BWD_FUSED=0"; do
  name=${v%%:*}; envs=${v#*:}
  env $envs timeout 300 python bench.py --quick --steps 20 > $O/r2c33_ab_${name}_$rep.json 2> $O/r2c33_ab_${name}_$rep.err
  python - "$name" "$rep" "$O/r2c33_ab_${name}_$rep.json" <<'PY'
import json,sys
d=json.load(open(sys.argv[3]))
print(sys.argv[1], sys.argv[2], round(d["value"],1), "e2e", round(d["e2e"]["value"],1), d["clocks"]["sm_mhz"], d["launches_per_step"])
PY
done
done
