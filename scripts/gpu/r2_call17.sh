#!/bin/bash
# round 2, GPU call 17: column sums fused in the halo kernel too; multi-seed gradient test; A/B; loss sweep
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_network.py -m gpu -q -s > $O/r2c17_net.log 2>&1; echo "net rc=$?"; grep -E "rel err|parity|passed|failed|Error|error|FAILED" $O/r2c17_net.log | tail -30
for rep in 1 2; do
for v in "fuse64:DDN_FUSE_BWD_STATS_MINC=64" "fuse128:DDN_FUSE_BWD_STATS_MINC=128"; do
  name=${v%%:*}; envs=${v#*:}
  env $envs timeout 300 python bench.py --quick --steps 20 > $O/r2c17_ab_${name}_$rep.json 2> $O/r2c17_ab_${name}_$rep.err
  python - "$name" "$rep" "$O/r2c17_ab_${name}_$rep.json" <<'PY'
import json,sys
d=json.load(open(sys.argv[3])); c=d["roofline"]["classes"]
print(sys.argv[1], sys.argv[2], round(d["value"],1), "e2e", round(d["e2e"]["value"],1), d["clocks"]["sm_mhz"], "host", round(d["host_enqueue_ms_per_step"],2), {k[5:-3]:round(v["ms"]/d["steps"],2) for k,v in c.items() if k.startswith("conv")})
PY
done
done
