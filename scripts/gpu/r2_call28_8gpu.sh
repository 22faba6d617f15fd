#!/bin/bash
# round 2, GPU call 28 (8 GPUs): final build, configs[1] and configs[4]
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
run() { tag=$1; n=$2; shift 2; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 10 --warmup 3 "$@" > $O/r2c28_bench_$tag.json 2> $O/r2c28_bench_$tag.err; echo "$tag rc=$?"; }
run c2_8 8
run c5_8 8 --config c5
run c2_2 2
timeout 300 python bench.py --quick --steps 10 > $O/r2c28_bench_c2_1.json 2> $O/r2c28_bench_c2_1.err; echo "1gpu rc=$?"
for f in c2_8 c5_8 c2_2 c2_1; do python - <<PY
import json
try:
    d=json.loads(open("$O/r2c28_bench_$f.json").read().strip().splitlines()[-1])
    print("$f", "value %.1f e2e %.1f ms %.2f"%(d["value"], d["e2e"]["value"], d["ms_per_step"]), d.get("allreduce_check"), d["clocks"]["sm_mhz"])
except Exception as e:
    print("$f", "ERR", e)
PY
done
