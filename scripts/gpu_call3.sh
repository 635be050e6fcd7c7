#!/bin/bash
# round-2 GPU check 3 (eight GPUs): config 5 under domain decomposition, peer-memory exchange
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
N=${1:-8}
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus $N --steps 3 --warmup 3 ) > gpurun_out/r2c_c5_n$N.json 2> gpurun_out/r2c_c5_n$N.err
echo "rc=$?"
python - <<PY
import json
f = 'r2c_c5_n$N'
try:
    line = [l for l in open('gpurun_out/%s.json' % f) if l.startswith('{')][-1]
    d = json.loads(line)
    print(f, 'value %.4g' % d['value'], 'e2e', d['e2e'] and '%.4g' % d['e2e']['value'], d['parallelism'],
          'parity', d.get('parity') and (d['parity']['ok'], d['parity']['force_rel_rms']),
          'comm', d['engine']['comm'], 'lists', d['engine']['list_stats'], 'roof', d['roofline'] and d['roofline']['avg_launch_us'])
except Exception as e:
    print(f, 'FAILED', e)
    import subprocess
    print(subprocess.run(['tail', '-25', 'gpurun_out/%s.err' % f], capture_output=True, text=True).stdout)
PY
