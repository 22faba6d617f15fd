#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out; mkdir -p $O
T='tests/test_gpu_network.py::test_whole_network_gradients_well_conditioned'
for v in "base:" "notail:DDN_TC_TAIL=0" "nopdl:DDN_PDL=0" "nofuse:DDN_FUSE_BWD_STATS_MINC=9999" "fuse64:DDN_FUSE_BWD_STATS_MINC=64"; do
  name=${v%%:*}; envs=${v#*:}
  env $envs timeout 300 python -m pytest "$T" -m gpu -q -s -k "3-2-64-96-train-bf16x3" > $O/r2c15_$name.log 2>&1
  echo "$name: $(grep -E 'well-conditioned|passed|failed' $O/r2c15_$name.log | tr '\n' ' ' | cut -c1-260)"
done
