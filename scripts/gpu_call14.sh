#!/bin/bash
# round-2 GPU check 14: full GPU suite (device-side ordering, trimmed list emission), skin x shell-delta sweep,
# end-to-end breakdown at config 5, source-level ncu of the list build
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q ) > gpurun_out/r2n_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2n_tests.log
tail -6 gpurun_out/r2n_tests.log
for cfg in "0.15 0.02" "0.18 0.02" "0.21 0.02" "0.15 0.01" "0.15 0.03" "0.18 0.03"; do
  set -- $cfg
  B2_SKIN=$1 B2_SHELL_DELTA=$2 timeout 400 python bench.py --reps 8 --steps 4 --warmup 3 --no-cpu-baseline --no-parity --no-e2e > gpurun_out/r2n_skin$1_delta$2.json 2> gpurun_out/r2n_skin$1_delta$2.err
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2n_skin*.json')):
    try:
        d = json.loads([l for l in open(f) if l.startswith('{')][-1])
        print(f, 'value %.4g' % d['value'], [(k['kernel'], k['avg_launch_us']) for k in d['roofline']['pair_kernels']],
              d['roofline']['phases_ms_per_md_step'], d['engine']['list_stats'], d['engine']['list_rebuilds'])
    except Exception as e:
        print(f, 'FAILED', e)
PY
B2_DEBUG_TIMING=1 timeout 900 python scripts/profile_e2e.py 14 > gpurun_out/r2n_e2e_c5.log 2>&1
grep -v "^ \|^$" gpurun_out/r2n_e2e_c5.log | head -40
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name 'regex:k_build_lists' \
    --launch-skip 0 --launch-count 1 -o gpurun_out/r2n_build_786k -f python scripts/profile_step.py 8 1 > gpurun_out/r2n_ncu_build.log 2>&1
ls -la gpurun_out/r2n_build_786k.ncu-rep
