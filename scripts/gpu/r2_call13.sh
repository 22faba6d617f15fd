#!/bin/bash
# round 2, GPU call 13: A/B of PDL / fused backward sums / halo with an uninstrumented timed region
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
for rep in 1 2; do
for v in "all:DDN_PDL=1 DDN_FUSE_BWD_STATS_MINC=512" "nopdl:DDN_PDL=0 DDN_FUSE_BWD_STATS_MINC=512" "pdl_nofuse:DDN_PDL=1 DDN_FUSE_BWD_STATS_MINC=9999" "pdl_fuse128:DDN_PDL=1 DDN_FUSE_BWD_STATS_MINC=128" "pdl_nohalo:DDN_PDL=1 DDN_FUSE_BWD_STATS_MINC=512 DDN_TC_HALO=0"; do
  name=${v%%:*}; envs=${v#*:}
  env $envs timeout 300 python bench.py --quick --steps 20 > $O/r2c13_ab_${name}_$rep.json 2> $O/r2c13_ab_${name}_$rep.err
  python - "$name" "$rep" "$O/r2c13_ab_${name}_$rep.json" <<'PY'
import json,sys
d=json.load(open(sys.argv[3])); c=d["roofline"]["classes"]
print(sys.argv[1], sys.argv[2], round(d["value"],1), "e2e", round(d["e2e"]["value"],1), d["clocks"]["sm_mhz"], {k[5:-3]:round(v["ms"]/d["steps"],2) for k,v in c.items() if k.startswith("conv")}, d["roofline"]["timed"][-75:])
PY
done
done
