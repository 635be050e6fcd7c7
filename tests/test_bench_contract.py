"""The reference arm of bench.py (CPU restatement on the host cores) prints ONE JSON line with the keys the
driver reads; ranks other than 0 print nothing and exit 0.  Runs without a GPU."""

import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(extra_env=None):
    env = dict(os.environ)
    env.update(extra_env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1',
                           '--warmup', '1', '--cpu-md-steps', '1', '--workload', 'c2'], capture_output=True, text=True, env=env,
                          timeout=600)


def test_reference_arm_json_line():
    result = run()
    assert result.returncode == 0, result.stderr[-2000:]
    lines = [line for line in result.stdout.splitlines() if line.strip()]
    assert len(lines) == 1
    record = json.loads(lines[0])
    assert record['impl'] == 'reference' and record['metric'] == 'atom-steps/s' and record['unit'] == 'atom-steps/s'
    assert record['higher_is_better'] is True and record['value'] > 0 and record['dtype'] == 'f64'
    assert record['config']['atoms'] == 98304 and 'workload' in record['config']
    baseline = record['cpu_baseline']
    assert baseline['kind'] == 'port' and baseline['cores'] >= 1 and baseline['value'] == record['value']
    assert record['e2e'] == dict(value=record['value'], unit='atom-steps/s', h2d_bytes_per_step=0, d2h_bytes_per_step=0)


def test_reference_arm_other_ranks_are_silent():
    result = run({'RANK': '1', 'WORLD_SIZE': '2', 'LOCAL_RANK': '1'})
    assert result.returncode == 0 and result.stdout.strip() == ''
