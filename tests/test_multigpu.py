"""GPU (>= 2 devices): the one-process-per-GPU data-parallel step over NCCL (dpc_b200.FlatTrainer) -- parameters identical
on every rank after a step and equal to the single-process two-shard replay.  Launched through torchrun exactly like bench.py
(tests/mp_worker.py).  SURVEY.md 8(e); /root/reference/dpc/main.py:65,229-231."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize('direct', [1, 0], ids=['library-allreduce', 'torch-allreduce'])
@pytest.mark.parametrize('world', [2])
def test_nccl_data_parallel_step_matches_single_process(world, direct):
    """direct = 1: the all-reduce is dpc_flat_allreduce (csrc/comm.cu) on the library's own communicator; 0: the
    torch.distributed fallback"""
    if torch.cuda.device_count() < world:
        pytest.skip('needs %d GPUs' % world)
    env = dict(os.environ, DPC_DIRECT_NCCL=str(direct))
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(world),
           '--master-addr', '127.0.0.1', '--master-port', '29631', os.path.join(ROOT, 'tests', 'mp_worker.py')]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    line = [l for l in out.stdout.splitlines() if l.startswith('MPRESULT ')]
    assert line, out.stdout[-2000:]
    res = json.loads(line[-1][len('MPRESULT '):])
    print(res)
    assert res['direct_nccl'] == bool(direct)
    assert res['same_init'], 'FlatTrainer did not broadcast rank 0\'s parameters'
    assert res['same_after'], 'parameters differ across ranks after the step'
    # the all-reduced gradient equals the single-process sum of the two shards' gradients up to the summation order of the
    # atomics (amplified by the tiny-batch conditioning, profiles/r2_grad_table.md); the FIRST Adam step is lr * sign(g), so
    # a near-zero gradient entry that changes sign moves the update by 2 lr: measured 3.9e-2 rel-L2 = 0.04 % of the entries
    assert res['grad_rel_l2'] < 1e-2, res
    assert res['update_rel_l2'] < 1e-1, res
