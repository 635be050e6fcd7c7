#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/r2g_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2g_tests.log
tail -12 gpurun_out/r2g_tests.log
( time timeout 900 python bench.py --steps 4 --warmup 3 --no-cpu-baseline ) > gpurun_out/r2g_c5.json 2> gpurun_out/r2g_c5.err
python - <<'PY'
import json
for f in ('r2g_c5',):
    try:
        d = json.loads([l for l in open('gpurun_out/%s.json' % f) if l.startswith('{')][-1])
        print(f, 'value %.4g' % d['value'], 'e2e %.4g' % d['e2e']['value'], 'parity', d['parity']['ok'],
              'pair us', d['roofline']['avg_launch_us'], d['roofline']['phases_ms_per_md_step'], d['engine']['list_stats'])
    except Exception as e:
        print(f, 'FAILED', e)
PY
