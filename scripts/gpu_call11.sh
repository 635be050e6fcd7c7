#!/bin/bash
# ncu source-level captures at config-5 size of the list build, the fused inner loop and the packed pair tiles
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
./scripts/micro/ffma2_rate > gpurun_out/r2k_ffma2.log 2>&1; cat gpurun_out/r2k_ffma2.log
timeout 900 python scripts/profile_step.py 14 2 > gpurun_out/r2k_plain_c5.log 2>&1 || exit 1
timeout 1200 ncu --set full --clock-control none --import-source on --kernel-name 'regex:k_build_lists' \
    --launch-skip 0 --launch-count 2 -o gpurun_out/r2k_build_c5 -f python scripts/profile_step.py 14 1 > gpurun_out/r2k_ncu_build.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on --kernel-name 'regex:k_inner|k_pair_force2' \
    --launch-skip 2 --launch-count 4 -o gpurun_out/r2k_inner_pair_c5 -f python scripts/profile_step.py 14 2 > gpurun_out/r2k_ncu_inner_pair.log 2>&1
B2_PAIR_SCALAR=1 timeout 1200 ncu --set full --clock-control none --import-source on --kernel-name 'regex:k_pair_force' \
    --launch-skip 2 --launch-count 2 -o gpurun_out/r2k_pair_scalar_c5 -f python scripts/profile_step.py 14 2 > gpurun_out/r2k_ncu_pair_scalar.log 2>&1
ls -la gpurun_out/*.ncu-rep
