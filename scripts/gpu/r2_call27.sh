#!/bin/bash
# round 2, GPU call 27: pipelined end-to-end loop, single-pass bf16 probe, final 1-GPU bench lines
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
for rep in 1 2 3; do
timeout 300 python bench.py --quick --steps 10 > $O/r2c27_bench_q$rep.json 2> $O/r2c27_bench_q$rep.err
python -c "
import json;d=json.load(open('$O/r2c27_bench_q$rep.json'));print('quick $rep', round(d['value'],1),'e2e',round(d['e2e']['value'],1),'adam',round(d['train_step_with_adam']['value'],1),d['clocks']['sm_mhz'])"
done
timeout 300 python bench.py --quick --steps 10 --precision bf16 > $O/r2c27_bench_bf16.json 2> $O/r2c27_bench_bf16.err
python -c "
import json;d=json.load(open('$O/r2c27_bench_bf16.json'));r=d['roofline'];print('bf16 single pass', round(d['value'],1),'pairs/s; classes', {k:round(v['flops']/v['ms']/1e9,1) for k,v in r['classes'].items() if 'flops' in v})"
timeout 600 python bench.py > $O/r2c27_bench_1gpu.json 2> $O/r2c27_bench_1gpu.err; echo "full bench rc=$?"
python -c "
import json;d=json.load(open('$O/r2c27_bench_1gpu.json'));print('full', round(d['value'],1),'e2e',round(d['e2e']['value'],1),'adam',round(d['train_step_with_adam']['value'],1),d['clocks'], 'cpu', d['cpu_baseline']['value'])"
