"""GPU, >= 2 devices: slab-decomposed run == single-GPU run, bitwise."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("overlap,temporal", [(1, 0), (0, 0), (1, 1), (0, 1), (1, 4), (0, 4), (1, 3)])
def test_slab_equals_single_gpu(overlap, temporal):
    n = _ngpu()
    if n < 2:
        pytest.skip("needs 2 GPUs")
    world = 4 if n >= 4 else 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(29611 + overlap + 2 * temporal),
           os.path.join(ROOT, "tests", "slab_worker.py"), "515", "300", "41", str(overlap), str(temporal)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
    out = json.loads(line)
    assert out["ok"] and out["max_abs_diff"] == 0.0 and out["world"] == world


def test_slab_with_obstacle_across_the_interface():
    """Multi-GPU obstacles: IBB cylinder straddling a slab interface."""
    n = _ngpu()
    if n < 2:
        pytest.skip("needs 2 GPUs")
    world = 4 if n >= 4 else 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", "29655",
           os.path.join(ROOT, "tests", "slab_obs_worker.py"), "60"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    out = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert out["ok"] and out["pop_equal"] and out["straddles"] and out["max_force_diff"] < 1e-12
