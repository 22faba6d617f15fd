#!/bin/bash
# round 2, GPU call 26: vectorised fc weight gradient, inference epilogue in the halo kernel: tests + bench
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --ignore=tests/test_gpu_multi.py > $O/r2c26_tests.log 2>&1; echo "tests rc=$?"; grep -E "passed|failed|FAILED|Error" $O/r2c26_tests.log | tail -5
timeout 600 python bench.py --steps 10 > $O/r2c26_bench.json 2> $O/r2c26_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2c26_bench.json"))
print(round(d["value"],1), "e2e", round(d["e2e"]["value"],1), d["clocks"], "adam", round(d["train_step_with_adam"]["value"],1))
f=d["forward_b16"]; print("fwd b16 train %.0f eval %.0f img/s"%(f["train_bn"]["imgs_per_s"], f["eval_bn"]["imgs_per_s"]), f["eval_bn"]["conv_kernels"])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r2c26_launches.csv python bench.py --profile-run --steps 1 > $O/r2c26_ncu_bench.log 2>&1; echo "ncu list rc=$?"
grep -E "fc_wgrad" $O/r2c26_launches.csv | awk -F'","' '{print $5, $NF}' | cut -c1-100
