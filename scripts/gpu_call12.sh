#!/bin/bash
# launch times of the pair tiles and the list build for the tuning variants in atomsmm_b200/variants/ (786 k atoms)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
: > gpurun_out/r2l_variants.txt
# the flattened group-level walk first has to reproduce the pair sets and forces
( B2_LIBRARY=$PWD/atomsmm_b200/variants/lib_v5.so timeout 600 python -m pytest tests/test_gpu_forces.py tests/test_gpu_scale.py -m gpu -q -x ) > gpurun_out/r2l_tests_v5.log 2>&1
tail -3 gpurun_out/r2l_tests_v5.log
run() {   # tag, env assignments...
    tag="$1"; shift
    env "$@" timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name 'regex:k_pair_force|k_build_lists|k_cell_scan' \
        --csv --log-file gpurun_out/r2l_$tag.csv python scripts/profile_step.py 8 6 > gpurun_out/r2l_$tag.log 2>&1
    python scripts/variant_times.py $tag gpurun_out/r2l_$tag.csv >> gpurun_out/r2l_variants.txt 2>&1
}
run scalar B2_PAIR_SCALAR=1 B2_LIBRARY=$PWD/atomsmm_b200/variants/lib_v1.so
for v in v1 v2 v3 v4 v5 v6; do run $v B2_LIBRARY=$PWD/atomsmm_b200/variants/lib_$v.so; done
run v5yz10 B2_CELL_YZ=1.0 B2_LIBRARY=$PWD/atomsmm_b200/variants/lib_v5.so
run v5yz035 B2_CELL_YZ=0.35 B2_LIBRARY=$PWD/atomsmm_b200/variants/lib_v5.so
cat gpurun_out/r2l_variants.txt
