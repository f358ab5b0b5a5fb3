"""CPU: bench.py's reference arm honours the driver contract -- ONE JSON line with the metric / config of our own arm, the
exact K + W steps it was asked for, `impl: reference`, a `cpu_baseline` describing the run and an `e2e` object."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '1'],
                         capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith('{')]
    assert len(lines) == 1, out.stdout[-2000:]
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['steps'] == 1 and d['warmup'] == 1 and d['n_gpus'] == 1
    for k in ('metric', 'value', 'unit', 'ms_per_step', 'higher_is_better', 'scaling', 'dtype', 'data', 'config', 'cpu_baseline',
              'e2e'):
        assert k in d, k
    assert d['unit'] == 'clips/s' and d['higher_is_better'] is True and d['value'] > 0
    assert d['config']['network'] == 'resnet18' and d['config']['img_dim'] == 128 and d['config']['batch_per_gpu'] == 128
    cb = d['cpu_baseline']
    assert cb['kind'] in ('reference', 'port') and cb['cores'] >= 1 and abs(cb['value'] - d['value']) < 1e-9
    assert d['e2e']['h2d_bytes_per_step'] == 0 and d['e2e']['d2h_bytes_per_step'] == 0
    # value = clips of the stated sample / seconds per step
    clips = d['config']['reference_sample']['clips_per_step']
    assert abs(d['value'] - clips / (d['ms_per_step'] / 1e3)) < 1e-6 * d['value']
