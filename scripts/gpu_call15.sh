#!/bin/bash
# round-2 GPU check 15: float64 energy kernel with the fp32 prefilter, cheaper list emission -- tests, end-to-end breakdown, bench
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q ) > gpurun_out/r2o_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2o_tests.log
tail -6 gpurun_out/r2o_tests.log
timeout 900 python scripts/profile_e2e.py 14 > gpurun_out/r2o_e2e_c5.log 2>&1
grep -v "^ \|^$" gpurun_out/r2o_e2e_c5.log | head -24
( time timeout 900 python bench.py --steps 4 --warmup 3 --no-cpu-baseline ) > gpurun_out/r2o_c5.json 2> gpurun_out/r2o_c5.err
python - <<'PY'
import json
for f in ('r2o_c5',):
    try:
        d = json.loads([l for l in open('gpurun_out/%s.json' % f) if l.startswith('{')][-1])
        print(f, 'value %.4g' % d['value'], 'e2e', d['e2e'] and '%.4g' % d['e2e']['value'], 'parity', d.get('parity') and d['parity'].get('ok'),
              [(k['kernel'], k['avg_launch_us'], k.get('slots_inside_cutoff')) for k in d['roofline']['pair_kernels']], d['roofline']['phases_ms_per_md_step'])
    except Exception as e:
        print(f, 'FAILED', e)
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/r2o_launches_c5.csv python scripts/profile_step.py 14 8 > gpurun_out/r2o_ncu_launches.log 2>&1
python scripts/summarize_launches.py gpurun_out/r2o_launches_c5.csv > gpurun_out/r2o_launches_c5.summary.txt 2>&1; head -30 gpurun_out/r2o_launches_c5.summary.txt
