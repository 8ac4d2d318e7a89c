"""Multi-GPU paths of optimize() on real devices (needs >= 2 GPUs; skipped on a 1-GPU box):
component sharding + all-gather and sample sharding + all-reduce, both against the single-GPU fit."""

import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_component_and_sample_sharding_match_single_gpu():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip('needs >= 2 GPUs')
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2',
           '--master-addr', '127.0.0.1', '--master-port', '29533', os.path.join(ROOT, 'tools', 'check_multi_gpu.py')]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert 'MULTI_GPU_CHECK' in r.stdout
