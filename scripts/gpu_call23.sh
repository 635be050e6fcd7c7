#!/bin/bash
# source-level ncu of the shipped list build and pair tiles (786 k atoms): where the warp instructions go now
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on --kernel-name 'regex:k_build_lists|k_pair_force' \
    --launch-skip 3 --launch-count 5 -o /tmp/r2x_hot -f python scripts/profile_step.py 8 3 > gpurun_out/r2x_ncu.log 2>&1
( python scripts/ncu_hot.py /tmp/r2x_hot.ncu-rep k_build_lists; python scripts/ncu_hot.py /tmp/r2x_hot.ncu-rep "LJCPot<(int)2"; python scripts/ncu_hot.py /tmp/r2x_hot.ncu-rep "LJCPot<(int)1" ) > gpurun_out/r2x_hot.txt 2>&1
python scripts/ncu_summary.py /tmp/r2x_hot.ncu-rep > gpurun_out/r2x_summary.txt 2>&1
head -50 gpurun_out/r2x_hot.txt
