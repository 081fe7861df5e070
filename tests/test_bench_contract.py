"""The JSON contract of bench.py: the reference arm on the CPU here, the GPU arm under -m gpu (small sizes)."""
import json
import os
import subprocess
import sys

import pytest

from tests.conftest import ROOT

COMMON = {'metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling', 'vs_baseline',
          'dtype', 'data', 'config', 'cpu_baseline', 'e2e', 'gpu_launches'}


def _run(args):
    proc = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py')] + args, capture_output=True, text=True, cwd=ROOT)
    assert proc.returncode == 0, proc.stderr[-2000:]
    lines = [l for l in proc.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, 'bench.py must print exactly one line on stdout'
    return json.loads(lines[0])


def test_reference_arm_contract():
    line = _run(['--impl', 'reference', '--queries', '3000', '--truth', '12000', '--steps', '2', '--warmup', '1', '--cpu-sample', '300'])
    assert COMMON <= set(line)
    assert line['impl'] == 'reference' and line['higher_is_better'] is True and line['unit'] == 'titles/s'
    assert line['cpu_baseline']['kind'] in ('reference', 'port') and line['cpu_baseline']['cores'] >= 1
    assert line['cpu_baseline']['value'] == line['value'] == line['e2e']['value']
    if line['cpu_baseline']['kind'] == 'reference':      # the staged reference's numba kernels, checked against the port live
        assert line['cpu_baseline']['port']['kind'] == 'port'
        assert line['parity']['reference_vs_port_mismatching_queries'] == 0
    assert abs(line['ms_per_step'] - 3000 / line['value'] * 1e3) < 1e-6 * line['ms_per_step']      # per step of ALL queries
    assert line['e2e']['h2d_bytes_per_step'] == 0 and line['e2e']['d2h_bytes_per_step'] == 0
    assert 'workload' in line['config'] and line['vs_baseline'] is None


@pytest.mark.gpu
def test_gpu_arm_contract():
    line = _run(['--queries', '4000', '--truth', '20000', '--steps', '2', '--warmup', '3', '--cpu-sample', '300', '--no-extra'])
    assert COMMON <= set(line)
    assert {'roofline', 'clocks', 'parity'} <= set(line)
    assert line['gpu_launches'] > 0 and line['value'] > 0 and line['e2e']['value'] > 0
    assert line['e2e']['h2d_bytes_per_step'] > 0 and line['e2e']['d2h_bytes_per_step'] > 0
    assert set(line['roofline']) >= {'bound', 'achieved', 'peak', 'unit', 'frac', 'traffic'}
    assert line['cpu_baseline']['kind'] in ('reference', 'port')
    assert line['parity']['mismatching_queries'] == 0 and line['parity']['e2e_equals_device_path'] is True
    assert line['parity'].get('reference_numba_mismatching_queries', 0) == 0
    assert 0.0 < line['config']['postings_hit_per_query_over_n'] < 2.0
    assert line['scaling'] in ('weak', 'strong') and line['dtype'] == 'f32' and line['data'] == 'synthetic'
