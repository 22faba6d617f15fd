#!/bin/bash
# round 2, GPU call 10: halo-tile kernel for the 64->64 convolutions (layer1), probe of stacked taps, A/B
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
timeout 120 scripts/probe_umma_offset > $O/r2c10_probe.txt 2>&1; echo "probe rc=$?"; tail -n 1 $O/r2c10_probe.txt
grep "MN-stack" $O/r2c10_probe.txt | grep -c ok; grep "MN-stack" $O/r2c10_probe.txt | grep MISMATCH | head -20
timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "conv2d" > $O/r2c10_ops.log 2>&1; echo "ops rc=$?"; grep -E "passed|failed|FAILED|Error|assert" $O/r2c10_ops.log | tail -8
timeout 900 python -m pytest tests/test_gpu_network.py -m gpu -q -s > $O/r2c10_net.log 2>&1; echo "net rc=$?"; grep -E "rel err|parity|passed|failed|Error|error|FAILED" $O/r2c10_net.log | tail -30
for rep in 1 2; do
for v in "halo:DDN_PDL=0 DDN_FUSE_BWD_STATS_MINC=9999" "nohalo:DDN_TC_HALO=0 DDN_PDL=0 DDN_FUSE_BWD_STATS_MINC=9999" "halo_pdl_fuse512:DDN_PDL=1 DDN_FUSE_BWD_STATS_MINC=512"; do
  name=${v%%:*}; envs=${v#*:}
  env $envs timeout 300 python bench.py --quick --steps 20 > $O/r2c10_ab_${name}_$rep.json 2> $O/r2c10_ab_${name}_$rep.err
  python - "$name" "$rep" "$O/r2c10_ab_${name}_$rep.json" <<'PY'
import json,sys
d=json.load(open(sys.argv[3])); c=d["roofline"]["classes"]
print(sys.argv[1], sys.argv[2], round(d["value"],1), "e2e", round(d["e2e"]["value"],1), d["clocks"]["sm_mhz"], {k[5:-3]:round(v["ms"]/d["steps"],2) for k,v in c.items() if k.startswith("conv")})
PY
done
done
