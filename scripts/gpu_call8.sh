#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_multi_rank.py -m gpu -x -q ) > gpurun_out/r2h_dd_p2p.log 2>&1
echo "p2p rc=$?" >> gpurun_out/r2h_dd_p2p.log
( time B2_DD_EXCHANGE=nccl timeout 600 python -m pytest tests/test_multi_rank.py -m gpu -x -q ) > gpurun_out/r2h_dd_nccl.log 2>&1
echo "nccl rc=$?" >> gpurun_out/r2h_dd_nccl.log
tail -3 gpurun_out/r2h_dd_p2p.log; tail -3 gpurun_out/r2h_dd_nccl.log
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 \
    bench.py --gpus 2 --steps 4 --warmup 3 ) > gpurun_out/r2h_c5_n2.json 2> gpurun_out/r2h_c5_n2.err
python - <<'PY'
import json
f = 'r2h_c5_n2'
try:
    d = json.loads([l for l in open('gpurun_out/%s.json' % f) if l.startswith('{')][-1])
    print(f, 'value %.4g' % d['value'], 'e2e %.4g' % d['e2e']['value'], d['parallelism'], 'parity', d['parity']['ok'], d['parity']['force_rel_rms'],
          d['roofline']['phases_ms_per_md_step_by_rank'], d['engine']['comm'])
except Exception as e:
    print(f, 'FAILED', e)
    import subprocess
    print(subprocess.run(['tail', '-25', 'gpurun_out/%s.err' % f], capture_output=True, text=True).stdout)
PY
