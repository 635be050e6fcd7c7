#!/bin/bash
# round-2 GPU check 17 (two GPUs): domain-decomposition parity test and config 5 on two ranks with the exchange clock
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
( B2_JIT_VERBOSE=1 timeout 300 python -m pytest tests/test_gpu_variants.py -m gpu -q -k "per_dof" ) > gpurun_out/r2r_jit.log 2>&1; tail -5 gpurun_out/r2r_jit.log
( time timeout 600 python -m pytest tests/test_multi_rank.py -m gpu -x -q ) > gpurun_out/r2r_dd_p2p.log 2>&1
echo "p2p rc=$?" >> gpurun_out/r2r_dd_p2p.log
tail -3 gpurun_out/r2r_dd_p2p.log
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 \
    bench.py --gpus 2 --steps 4 --warmup 3 ) > gpurun_out/r2r_c5_n2.json 2> gpurun_out/r2r_c5_n2.err
python - <<'PY'
import json
f = 'r2r_c5_n2'
try:
    d = json.loads([l for l in open('gpurun_out/%s.json' % f) if l.startswith('{')][-1])
    print(f, 'value %.4g' % d['value'], 'e2e %.4g' % d['e2e']['value'], d['parallelism'], 'parity', d['parity']['ok'], d['parity']['force_rel_rms'],
          d['roofline']['phases_ms_per_md_step_by_rank'], d['engine']['comm'])
except Exception as e:
    print(f, 'FAILED', e)
    import subprocess
    print(subprocess.run(['tail', '-25', 'gpurun_out/%s.err' % f], capture_output=True, text=True).stdout)
PY
