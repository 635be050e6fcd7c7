#!/bin/bash
# shipped build (96-atom inner-loop chunks): smoke() and the default bench line without the CPU leg
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
( time timeout 300 python __graft_entry__.py smoke ) > gpurun_out/r2v_smoke.log 2>&1; tail -2 gpurun_out/r2v_smoke.log
( time timeout 900 python bench.py --no-cpu-baseline ) > gpurun_out/r2v_c5.json 2> gpurun_out/r2v_c5.err
python - <<'PY'
import json
d = json.loads([l for l in open('gpurun_out/r2v_c5.json') if l.startswith('{')][-1])
print('value %.4g' % d['value'], 'e2e %.4g' % d['e2e']['value'], 'parity', d['parity']['ok'], d['roofline']['phases_ms_per_md_step'])
PY
