#!/bin/bash
# round 2, GPU call 11: launch list of the current build + one full capture of the halo-tile kernel
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
export DDN_PDL=0 DDN_FUSE_BWD_STATS_MINC=9999
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r2c11_launches.csv python bench.py --profile-run --steps 1 > $O/r2c11_ncu_bench.log 2>&1; echo "ncu list rc=$?"
NCU="ncu --set full --clock-control none --import-source on"
timeout 300 $NCU -k regex:conv64_halo -s 14 -c 2 -o $O/r2_prof_conv64_halo -f python bench.py --profile-run --steps 1 > $O/r2c11_ncu1.log 2>&1; echo "ncu halo rc=$?"
ls -la $O/*.ncu-rep | tail -3
