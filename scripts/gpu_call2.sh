#!/bin/bash
# round-2 GPU check 2 (two GPUs): domain decomposition with the peer-memory halo exchange
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2b_gpus.txt 2>&1
nvidia-smi topo -m >> gpurun_out/r2b_gpus.txt 2>&1
( time timeout 600 python -m pytest tests/test_multi_rank.py -m gpu -x -q ) > gpurun_out/r2b_dd_p2p.log 2>&1
echo "p2p rc=$?" >> gpurun_out/r2b_dd_p2p.log
( time B2_DD_EXCHANGE=nccl timeout 600 python -m pytest tests/test_multi_rank.py -m gpu -x -q ) > gpurun_out/r2b_dd_nccl.log 2>&1
echo "nccl rc=$?" >> gpurun_out/r2b_dd_nccl.log
run() { # name, extra env, args
  ( time env $2 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus 2 $3 ) > gpurun_out/$1.json 2> gpurun_out/$1.err
  echo "$1 rc=$?"
}
run r2b_c5_n2_p2p "B2_DUMMY=1" "--steps 3 --warmup 3"
run r2b_c5_n2_nccl "B2_DD_EXCHANGE=nccl" "--steps 3 --warmup 3 --no-parity --no-e2e"
tail -4 gpurun_out/r2b_dd_p2p.log; tail -4 gpurun_out/r2b_dd_nccl.log
python - <<'PY'
import json
for f in ('r2b_c5_n2_p2p', 'r2b_c5_n2_nccl'):
    try:
        line = [l for l in open('gpurun_out/%s.json' % f) if l.startswith('{')][-1]
        d = json.loads(line)
        print(f, 'value %.4g' % d['value'], 'e2e', d['e2e'] and '%.4g' % d['e2e']['value'], d['parallelism'],
              'parity', d.get('parity') and (d['parity']['ok'], d['parity']['force_rel_rms'], d['parity']['energy_rel']),
              'comm', d['engine']['comm'], 'rebuilds', d['engine']['list_stats'])
    except Exception as e:
        print(f, 'FAILED', e)
        import subprocess
        print(subprocess.run(['tail', '-15', 'gpurun_out/%s.err' % f], capture_output=True, text=True).stdout)
PY
