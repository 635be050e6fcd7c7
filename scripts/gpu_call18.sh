#!/bin/bash
# round-2 final single-GPU evidence: default bench (config 5) with CPU baseline, reference arm, config 2,
# ncu launch list and full metric set of one MD step's kernels at config-5 size (summaries only: gpurun_out is capped at 64 MiB)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
( time timeout 900 python bench.py ) > gpurun_out/r2s_c5.json 2> gpurun_out/r2s_c5.err
( time timeout 900 python bench.py --impl reference ) > gpurun_out/r2s_c5_reference.json 2> gpurun_out/r2s_c5_reference.err
( time timeout 600 python bench.py --workload c2 ) > gpurun_out/r2s_c2.json 2> gpurun_out/r2s_c2.err
python - <<'PY'
import json
for f in ('r2s_c5', 'r2s_c5_reference', 'r2s_c2'):
    try:
        d = json.loads([l for l in open('gpurun_out/%s.json' % f) if l.startswith('{')][-1])
        print(f, 'value %.4g' % d['value'], 'e2e', d.get('e2e') and '%.4g' % d['e2e']['value'], 'parity', d.get('parity') and d['parity'].get('ok'))
    except Exception as e:
        print(f, 'FAILED', e)
PY
: > gpurun_out/r2s_inner_variants.txt
for v in c128 c96 c64 c32; do
    B2_LIBRARY=$PWD/atomsmm_b200/variants/lib_$v.so timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name 'regex:k_inner' \
        --csv --log-file /tmp/r2s_$v.csv python scripts/profile_step.py 8 4 > /tmp/r2s_$v.log 2>&1
    python scripts/variant_times.py $v /tmp/r2s_$v.csv >> gpurun_out/r2s_inner_variants.txt 2>&1
done
cat gpurun_out/r2s_inner_variants.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2s_launches_c5.csv python scripts/profile_step.py 14 8 > gpurun_out/r2s_ncu_launches.log 2>&1
python scripts/summarize_launches.py gpurun_out/r2s_launches_c5.csv > gpurun_out/r2s_launches_c5.summary.txt 2>&1
timeout 1500 ncu --set full --clock-control none \
    --kernel-name 'regex:k_inner|k_vel|k_skin_check|k_pair_force|k_build_lists|k_group_geom|k_cell_sort_pack|k_save_ref|k_pair_band|k_cell_scan' \
    --launch-skip 60 --launch-count 45 -o /tmp/r2s_full_c5 -f python scripts/profile_step.py 14 8 > gpurun_out/r2s_ncu_full_c5.log 2>&1
python scripts/ncu_summary.py /tmp/r2s_full_c5.ncu-rep > gpurun_out/r2s_ncu_full_c5.summary.txt 2>&1
grep -c "====" gpurun_out/r2s_ncu_full_c5.summary.txt; du -sh gpurun_out
