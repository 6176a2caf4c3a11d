"""bench.py's JSON contract (reference arm runs on CPU; the GPU arm is covered by the `gpu` marker)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {'metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling',
             'vs_baseline', 'dtype', 'data', 'config', 'e2e', 'gpu_launches'}


def test_reference_arm_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '2',
                          '--warmup', '1', '--workload', 'c2', '--shape', '512', '512'],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert BASE_KEYS <= set(line)
    assert line['impl'] == 'reference' and line['unit'] == 'Mcell-updates/s' and line['value'] > 0
    assert line['cpu_baseline']['kind'] == 'port' and line['cpu_baseline']['cores'] >= 1
    assert line['e2e'] == {'value': line['value'], 'unit': line['unit'], 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    assert 'workload' in line['config'] and line['vs_baseline'] is None


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK='1', WORLD_SIZE='2')
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--gpus', '2'],
                         capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ''


@pytest.mark.gpu
@pytest.mark.no_launch          # the launches happen in the bench.py process this test starts
def test_gpu_arm_line_small():
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--steps', '3', '--warmup', '3',
                          '--workload', 'c3', '--shape', '128', '128', '256', '--e2e-steps', '1', '--no-cpu-baseline'],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert BASE_KEYS | {'roofline', 'clocks'} <= set(line)
    assert line['gpu_launches'] == 6
    assert line['roofline']['bound'] == 'hbm' and 0 < line['roofline']['frac'] < 1.5
    assert line['e2e']['h2d_bytes_per_step'] == 2 * 128 * 128 * 256 * 4
    assert line['config']['kernel_variants'] == {'forward': 'march', 'adjoint': 'march'}
