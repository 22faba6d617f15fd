#!/bin/bash
# round 2, GPU call 6 (2 GPUs): NCCL correctness of the overlapped all-reduce + A/B overlap vs after-backward
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi -L > $O/r2c22_gpus.txt
timeout 400 python -m pytest tests/test_gpu_multi.py -m gpu -q -s > $O/r2c22_multi.log 2>&1; echo "multi rc=$?"; grep -E "RANK|passed|failed|Error" $O/r2c22_multi.log | tail -8
run() { tag=$1; shift; timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 "$@" > $O/r2c22_bench_$tag.json 2> $O/r2c22_bench_$tag.err; echo "$tag rc=$?"; }
run overlap
run nooverlap --no-overlap
run overlap_two --two-calls
timeout 300 python bench.py --quick --steps 10 > $O/r2c22_bench_1gpu.json 2> $O/r2c22_bench_1gpu.err; echo "1gpu rc=$?"
for f in overlap nooverlap overlap_two 1gpu; do python - <<PY
import json
try:
    d=json.loads(open("$O/r2c22_bench_$f.json").read().strip().splitlines()[-1])
    print("$f", "value %.1f e2e %.1f ms %.2f"%(d["value"], d["e2e"]["value"], d["ms_per_step"]), d.get("allreduce_check"), d.get("allreduce_detail"))
except Exception as e:
    print("$f", "ERR", e)
PY
done
tail -n 5 $O/r2c22_bench_overlap.err
