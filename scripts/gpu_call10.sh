#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
./scripts/micro/ffma2_rate > gpurun_out/r2j_ffma2.log 2>&1
cat gpurun_out/r2j_ffma2.log
( time timeout 900 python -m pytest tests -m gpu -q ) > gpurun_out/r2j_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2j_tests.log
tail -8 gpurun_out/r2j_tests.log
( time timeout 900 python bench.py --steps 4 --warmup 3 --no-cpu-baseline ) > gpurun_out/r2j_c5.json 2> gpurun_out/r2j_c5.err
( time B2_PAIR_SCALAR=1 timeout 900 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-parity --no-e2e ) > gpurun_out/r2j_c5_scalar.json 2> gpurun_out/r2j_c5_scalar.err
( time timeout 600 python bench.py --workload c2 --steps 5 --warmup 3 --no-cpu-baseline ) > gpurun_out/r2j_c2.json 2> gpurun_out/r2j_c2.err
python - <<'PY'
import json
for f in ('r2j_c5', 'r2j_c5_scalar', 'r2j_c2'):
    try:
        d = json.loads([l for l in open('gpurun_out/%s.json' % f) if l.startswith('{')][-1])
        print(f, 'value %.4g' % d['value'], 'e2e', d['e2e'] and '%.4g' % d['e2e']['value'], 'parity', d.get('parity') and d['parity'].get('ok'),
              d['roofline']['avg_launch_us'], d['roofline']['phases_ms_per_md_step'])
    except Exception as e:
        print(f, 'FAILED', e)
PY
bash scripts/gpu_call11.sh
