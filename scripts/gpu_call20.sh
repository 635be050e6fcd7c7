#!/bin/bash
# quick check of the 96-atom chunks of the fused inner loop: integrator / constraint / variant tests
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_integrators.py tests/test_gpu_constraints.py tests/test_gpu_variants.py tests/test_gpu_scale.py -m gpu -q ) > gpurun_out/r2u_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2u_tests.log
tail -5 gpurun_out/r2u_tests.log
