#!/bin/bash
# round-2 GPU check 13: tests of the new list build (run table + flattened walk, core/shell entry order) and inner loop,
# fused-inner-loop variants, shell-delta sweep, config-5 bench
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q ) > gpurun_out/r2m_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2m_tests.log
tail -6 gpurun_out/r2m_tests.log
: > gpurun_out/r2m_variants.txt
run() {   # tag, env assignments...
    tag="$1"; shift
    env "$@" timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name 'regex:k_inner|k_build_lists' \
        --csv --log-file gpurun_out/r2m_$tag.csv python scripts/profile_step.py 8 4 > gpurun_out/r2m_$tag.log 2>&1
    python scripts/variant_times.py $tag gpurun_out/r2m_$tag.csv >> gpurun_out/r2m_variants.txt 2>&1
}
for v in i60 i61 i51 i41; do run $v B2_LIBRARY=$PWD/atomsmm_b200/variants/lib_$v.so; done
cat gpurun_out/r2m_variants.txt
for d in 0 0.02 0.04 0.06 0.09; do
  B2_SHELL_DELTA=$d timeout 400 python bench.py --reps 8 --steps 4 --warmup 3 --no-cpu-baseline --no-parity --no-e2e > gpurun_out/r2m_delta_$d.json 2> gpurun_out/r2m_delta_$d.err
done
( time timeout 900 python bench.py --steps 4 --warmup 3 --no-cpu-baseline ) > gpurun_out/r2m_c5.json 2> gpurun_out/r2m_c5.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2m_delta_*.json')) + ['gpurun_out/r2m_c5.json']:
    try:
        d = json.loads([l for l in open(f) if l.startswith('{')][-1])
        print(f, 'value %.4g' % d['value'], 'e2e', d['e2e'] and '%.4g' % d['e2e']['value'], 'parity', d.get('parity') and d['parity'].get('ok'),
              [(k['kernel'], k['avg_launch_us']) for k in d['roofline']['pair_kernels']], d['roofline']['phases_ms_per_md_step'])
    except Exception as e:
        print(f, 'FAILED', e)
PY
