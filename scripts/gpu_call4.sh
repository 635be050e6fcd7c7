#!/bin/bash
# round-2 GPU check 4 (one GPU): full GPU test-suite, compute-sanitizer on the smoke workload, ncu at config-5 size
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/r2d_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2d_tests.log
( time timeout 900 /usr/local/cuda/bin/compute-sanitizer --tool memcheck --error-exitcode 7 python __graft_entry__.py smoke ) > gpurun_out/r2d_memcheck.log 2>&1
echo "memcheck rc=$?" >> gpurun_out/r2d_memcheck.log
( time timeout 900 /usr/local/cuda/bin/compute-sanitizer --tool racecheck --error-exitcode 7 python __graft_entry__.py smoke ) > gpurun_out/r2d_racecheck.log 2>&1
echo "racecheck rc=$?" >> gpurun_out/r2d_racecheck.log
# launch list of 8 MD steps of config 5 (4.2 M atoms) and of config 2
timeout 900 python scripts/profile_step.py 14 8 > gpurun_out/r2d_plain_c5.log 2>&1 && \
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2d_launches_c5.csv \
    python scripts/profile_step.py 14 8 > gpurun_out/r2d_ncu_c5.log 2>&1
timeout 900 python scripts/profile_step.py 4 8 > gpurun_out/r2d_plain_c2.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2d_launches_c2.csv \
    python scripts/profile_step.py 4 8 > gpurun_out/r2d_ncu_c2.log 2>&1
# full metric sets: the step's kernels at config-5 size (bandwidth-bound integrator / list kernels, pair kernels)
timeout 1500 ncu --set full --clock-control none \
    --kernel-name 'regex:k_inner|k_vel|k_skin_check|k_pair_force|k_build_lists|k_group_geom|k_cell_sort_pack|k_save_ref|k_pair_band|k_cell_fill' \
    --launch-skip 80 --launch-count 26 -o /tmp/r2d_full_c5 -f python scripts/profile_step.py 14 8 > gpurun_out/r2d_ncu_full_c5.log 2>&1
python scripts/ncu_summary.py /tmp/r2d_full_c5.ncu-rep > gpurun_out/r2d_ncu_full_c5.summary.txt 2>&1
# the pair kernels of the shipped build at config 2 with source correlation
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name 'regex:k_pair_force' \
    --launch-skip 12 --launch-count 3 -o gpurun_out/r2d_pair_c2 -f python scripts/profile_step.py 4 8 > gpurun_out/r2d_ncu_pair_c2.log 2>&1
# skin sensitivity at config 2 (resident rate only)
for skin in 0.08 0.10 0.12 0.15 0.20; do
  B2_SKIN=$skin timeout 300 python bench.py --workload c2 --steps 4 --warmup 3 --no-cpu-baseline --no-parity --no-e2e > gpurun_out/r2d_skin_$skin.json 2>/dev/null
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2d_skin_*.json')):
    try:
        d = json.loads([l for l in open(f) if l.startswith('{')][-1])
        print(f, 'value %.4g' % d['value'], d['roofline']['phases_ms_per_md_step'], d['engine']['list_stats'])
    except Exception as e:
        print(f, 'FAILED', e)
PY
du -sh gpurun_out; ls -la gpurun_out/ | tail -30
tail -5 gpurun_out/r2d_tests.log; tail -4 gpurun_out/r2d_memcheck.log; tail -4 gpurun_out/r2d_racecheck.log
