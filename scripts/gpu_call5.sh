#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python scripts/nhl_probe.py > gpurun_out/r2f_nhl.log 2>&1; cat gpurun_out/r2f_nhl.log | tail -12
( time timeout 900 python -m pytest tests/test_gpu_variants.py tests/test_gpu_forces.py -q ) > gpurun_out/r2f_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2f_tests.log
tail -8 gpurun_out/r2f_tests.log
for cfg in "1000 4.0 4 2" "2000 1.0 2 1"; do
  echo "== drift $cfg" ; timeout 600 python scripts/drift_probe.py $cfg 2>&1 | tail -9
done > gpurun_out/r2f_drift.log 2>&1
cat gpurun_out/r2f_drift.log
