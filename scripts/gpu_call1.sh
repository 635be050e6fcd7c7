#!/bin/bash
# round-2 GPU check 1 (one GPU): GPU test-suite, determinism (two identical c2 runs), default bench (c5) with parity gate
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2a_gpus.txt 2>&1
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2a_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2a_tests.log
( time timeout 600 python bench.py --workload c2 --steps 5 --warmup 3 ) > gpurun_out/r2a_c2_run1.json 2> gpurun_out/r2a_c2_run1.err
( time timeout 600 python bench.py --workload c2 --steps 5 --warmup 3 --no-cpu-baseline ) > gpurun_out/r2a_c2_run2.json 2> gpurun_out/r2a_c2_run2.err
( time timeout 1200 python bench.py --steps 3 --warmup 3 ) > gpurun_out/r2a_c5.json 2> gpurun_out/r2a_c5.err
tail -3 gpurun_out/r2a_tests.log
python - <<'PY'
import json
for f in ('r2a_c2_run1', 'r2a_c2_run2', 'r2a_c5'):
    try:
        line = [l for l in open('gpurun_out/%s.json' % f) if l.startswith('{')][-1]
        d = json.loads(line)
        print(f, 'value %.4g' % d['value'], 'e2e', d['e2e'] and '%.4g' % d['e2e']['value'], 'E', d['e2e'] and d['e2e']['final_energy'],
              'parity', d.get('parity') and (d['parity']['ok'], d['parity']['force_rel_rms'], d['parity']['energy_rel'], d['parity']['pair_sets']),
              'roof', d['roofline'] and round(d['roofline']['frac'], 4))
    except Exception as e:
        print(f, 'FAILED', e)
PY
