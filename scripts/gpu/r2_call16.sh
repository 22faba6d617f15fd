#!/bin/bash
# round 2, GPU call 16: halo weight-gradient kernel for the 64x64 convs, multi-seed gradient test, bench
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k conv2d > $O/r2c16_ops.log 2>&1; echo "ops rc=$?"; grep -E "passed|failed|FAILED|Error|assert" $O/r2c16_ops.log | tail -6
timeout 900 python -m pytest tests/test_gpu_network.py -m gpu -q -s > $O/r2c16_net.log 2>&1; echo "net rc=$?"; grep -E "rel err|parity|passed|failed|Error|error|FAILED" $O/r2c16_net.log | tail -30
for rep in 1 2; do
for v in "fuse128:DDN_FUSE_BWD_STATS_MINC=128" "fuse256:DDN_FUSE_BWD_STATS_MINC=256"; do
  name=${v%%:*}; envs=${v#*:}
  env $envs timeout 300 python bench.py --quick --steps 20 > $O/r2c16_ab_${name}_$rep.json 2> $O/r2c16_ab_${name}_$rep.err
  python - "$name" "$rep" "$O/r2c16_ab_${name}_$rep.json" <<'PY'
import json,sys
d=json.load(open(sys.argv[3])); c=d["roofline"]["classes"]
print(sys.argv[1], sys.argv[2], round(d["value"],1), "e2e", round(d["e2e"]["value"],1), d["clocks"]["sm_mhz"], {k[5:-3]:round(v["ms"]/d["steps"],2) for k,v in c.items() if k.startswith("conv")})
PY
done
done
DDN_FUSE_BWD_STATS_MINC=9999 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r2c16_launches.csv python bench.py --profile-run --steps 1 > $O/r2c16_ncu_bench.log 2>&1; echo "ncu list rc=$?"
grep -c wgrad64_halo $O/r2c16_launches.csv
