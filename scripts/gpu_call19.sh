#!/bin/bash
# round-2 final: config 5 under domain decomposition on 8 GPUs
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus 8 --steps 5 --warmup 3 ) > gpurun_out/r2t_c5_n8.json 2> gpurun_out/r2t_c5_n8.err
python - <<'PY'
import json
f = 'r2t_c5_n8'
try:
    d = json.loads([l for l in open('gpurun_out/%s.json' % f) if l.startswith('{')][-1])
    print(f, 'value %.4g' % d['value'], 'e2e %.4g' % d['e2e']['value'], d['parallelism'], 'parity', d['parity']['ok'],
          d['roofline']['phases_ms_per_md_step_by_rank'], d['engine']['comm'])
except Exception as e:
    print(f, 'FAILED', e)
    import subprocess
    print(subprocess.run(['tail', '-25', 'gpurun_out/%s.err' % f], capture_output=True, text=True).stdout)
PY
