"""Multi-GPU parity (needs >= 2 GPUs on the box, skipped otherwise): the population-sharded plan with
the per-iteration NCCL all-gather of (return, cost) must equal the one-GPU plan bit for bit on every
rank, in both precisions (tools/multi_gpu_check.py, launched with torchrun on 127.0.0.1)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_population_sharded_plan_equals_single_gpu_plan():
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs (have %d)" % n)
    world = 2
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(world),
           '--master-addr', '127.0.0.1', '--master-port', '29631',
           os.path.join(ROOT, 'tools', 'multi_gpu_check.py')]
    out = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    sys.stdout.write(out.stdout[-3000:])
    assert out.returncode == 0, out.stderr[-3000:]
    assert out.stdout.count("bit for bit on every rank: True") == 2
    assert "replicas identical: True" in out.stdout
    assert "equal to the local plans: True" in out.stdout
