#!/bin/bash
# round-2 GPU check 16: run-time compiled per-DOF steps, device-side kinetic energy / pinned downloads -- tests, e2e breakdown
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q ) > gpurun_out/r2p_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2p_tests.log
tail -8 gpurun_out/r2p_tests.log
timeout 900 python scripts/profile_e2e.py 14 > gpurun_out/r2p_e2e_c5.log 2>&1
grep -v "^ \|^$" gpurun_out/r2p_e2e_c5.log | head -24
