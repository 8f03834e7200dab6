"""world_size-2 NCCL check of the batch-sharded step on real GPUs (skipped on a single-GPU box; bench.py --gpus N
repeats the replica-identity part inside every multi-GPU benchmark run: `ddp_weights_identical`)."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize('workload,batch_dice', [('cfg2', 0), ('cfg2', 1), ('cfg4', 0)])
def test_split_step_equals_union_step_and_replicas_stay_identical(workload, batch_dice):
    if torch.cuda.device_count() < 2:
        pytest.skip('needs two GPUs')
    port = 29600 + (os.getpid() % 300)
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr',
           '127.0.0.1', '--master-port', str(port), os.path.join(ROOT, 'scripts', 'ddp_parity_check.py'),
           '--workload', workload, '--batch-dice', str(batch_dice)]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    lines = [l for l in p.stdout.splitlines() if l.startswith('{')]
    assert p.returncode == 0 and lines, (p.returncode, p.stdout[-2000:], p.stderr[-2000:])
    res = json.loads(lines[-1])
    assert res['ok'] and res['identical_after_initialize'] and res['identical_after_steps'], res
    assert res['split_vs_emulated_update_rel_err'] < 5e-3 and res['split_vs_union_update_rel_err'] < 0.15, res
