#!/bin/bash
# round 2, GPU call 9: tcgen05 descriptor probe (unaligned swizzle-atom starts) + a longer, interleaved A/B of PDL and the fused backward column sums
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
timeout 120 scripts/probe_umma_offset > $O/r2c9_probe.txt 2>&1; echo "probe rc=$?"; tail -n 3 $O/r2c9_probe.txt
grep -c ok $O/r2c9_probe.txt; grep MISMATCH $O/r2c9_probe.txt | head -40
for rep in 1 2 3; do
for v in "neither:DDN_PDL=0 DDN_FUSE_BWD_STATS_MINC=9999" "pdl:DDN_PDL=1 DDN_FUSE_BWD_STATS_MINC=9999" "fuse512:DDN_PDL=0 DDN_FUSE_BWD_STATS_MINC=512" "fuse128:DDN_PDL=0 DDN_FUSE_BWD_STATS_MINC=128" "pdl_fuse512:DDN_PDL=1 DDN_FUSE_BWD_STATS_MINC=512"; do
  name=${v%%:*}; envs=${v#*:}
  env $envs timeout 300 python bench.py --quick --steps 30 > $O/r2c9_ab_${name}_$rep.json 2> $O/r2c9_ab_${name}_$rep.err
  python - "$name" "$rep" "$O/r2c9_ab_${name}_$rep.json" <<'PY'
import json,sys
d=json.load(open(sys.argv[3])); c=d["roofline"]["classes"]
print(sys.argv[1], sys.argv[2], round(d["value"],1), "e2e", round(d["e2e"]["value"],1), d["clocks"]["sm_mhz"], {k[5:-3]:round(v["ms"]/d["steps"],2) for k,v in c.items() if k.startswith("conv")})
PY
done
done
