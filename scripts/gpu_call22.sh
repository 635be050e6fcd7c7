#!/bin/bash
# final full GPU suite on the shipped build
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q ) > gpurun_out/r2w_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2w_tests.log
tail -6 gpurun_out/r2w_tests.log
