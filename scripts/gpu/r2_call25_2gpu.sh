#!/bin/bash
# round 2, GPU call 25 (2 GPUs): SMs reserved only while an all-reduce overlaps the backward: A/B
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
timeout 400 python -m pytest tests/test_gpu_multi.py -m gpu -q -s > $O/r2c25_multi.log 2>&1; echo "multi rc=$?"; grep -E "passed|failed|Error" $O/r2c25_multi.log | tail -3
run() { tag=$1; shift; timeout 400 env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --quick > $O/r2c25_bench_$tag.json 2> $O/r2c25_bench_$tag.err; echo "$tag rc=$?"; }
for rep in 1 2; do
run win8_$rep DDN_OVERLAP_RESERVED_SMS=8
run win0_$rep DDN_OVERLAP_RESERVED_SMS=0
run win4_$rep DDN_OVERLAP_RESERVED_SMS=4
run win16_$rep DDN_OVERLAP_RESERVED_SMS=16
run win0cap8_$rep DDN_OVERLAP_RESERVED_SMS=0 NCCL_MAX_CTAS=8
done
timeout 300 python bench.py --quick --steps 10 > $O/r2c25_bench_1gpu.json 2> $O/r2c25_bench_1gpu.err; echo "1gpu rc=$?"
for f in win8_1 win0_1 win4_1 win16_1 win0cap8_1 win8_2 win0_2 win4_2 win16_2 win0cap8_2 1gpu; do python - <<PY
import json
try:
    d=json.loads(open("$O/r2c25_bench_$f.json").read().strip().splitlines()[-1])
    print("$f", "value %.1f e2e %.1f ms %.2f"%(d["value"], d["e2e"]["value"], d["ms_per_step"]), d.get("allreduce_check"))
except Exception as e:
    print("$f", "ERR", e)
PY
done
