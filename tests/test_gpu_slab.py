"""GPU, >= 2 devices: slab-decomposed run == single-GPU run, bitwise, for both
halo-exchange mechanisms: peer stores over NVLink (CUDA IPC, the default) and NCCL send/recv."""
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


def _torchrun(world, port, script, *args):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", script)] + [str(a) for a in args]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    return json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])


def _world():
    n = _ngpu()
    if n < 2:
        pytest.skip("needs 2 GPUs")
    return 4 if n >= 4 else 2


@pytest.mark.parametrize("exchange", ["peer", "nccl"])
@pytest.mark.parametrize("overlap,temporal", [(1, 0), (0, 0), (1, 1), (0, 1), (1, 4), (0, 4), (1, 3)])
def test_slab_equals_single_gpu(overlap, temporal, exchange):
    if exchange == "peer" and overlap == 0:
        pytest.skip("peer exchange has no overlap switch")
    world = _world()
    out = _torchrun(world, 29611 + overlap + 2 * temporal + (10 if exchange == "peer" else 0), "slab_worker.py",
                    515, 300, 41, overlap, temporal, exchange)
    assert out["ok"] and out["max_abs_diff"] == 0.0 and out["world"] == world and out["exchange"] == exchange


@pytest.mark.parametrize("exchange", ["peer", "nccl"])
def test_slab_with_ramp_table(exchange):
    """Base wall profiles + per-update ramp scalars (lbm_set_ramp) on every rank == single GPU, bitwise."""
    world = _world()
    out = _torchrun(world, 29671, "slab_worker.py", 515, 300, 41, 1, 4, exchange, 1)
    assert out["ok"] and out["max_abs_diff"] == 0.0 and out["ramp"]


@pytest.mark.parametrize("exchange,depth", [("peer", 4), ("nccl", 4), ("nccl", 0), ("peer", 1)])
def test_large_slabs_equal_single_gpu(exchange, depth):
    """Slabs large enough for a launch to run for milliseconds (4100 x 3000 cells): a halo exchange that reads
    columns before the launch that writes them has finished, or a neighbour that starts its next launch before the
    halos have landed, shows up here (it cannot on 515 x 300, where a launch takes microseconds).  Populations
    compared through their bit-pattern checksums."""
    world = _world()
    out = _torchrun(world, 29681 + depth, "slab_worker.py", 4100, 3000, 13, 1, depth, exchange)
    assert out["ok"] and out["compared"] == "checksum"


@pytest.mark.parametrize("exchange", ["peer", "nccl"])
def test_slab_with_obstacle_across_the_interface(exchange):
    """Multi-GPU obstacles: IBB cylinder straddling a slab interface."""
    world = _world()
    out = _torchrun(world, 29655, "slab_obs_worker.py", 60, exchange)
    assert out["ok"] and out["pop_equal"] and out["straddles"] and out["max_force_diff"] < 1e-12


@pytest.mark.parametrize("exchange", ["peer", "nccl"])
@pytest.mark.parametrize("place", ["interior", "interface"])
def test_slab_obstacle_band_with_multi_update_groups(exchange, place):
    """Bodies + four updates per group on slabs: a cylinder inside one slab gets an obstacle band there while the
    other ranks run plain wavefront launches (multi_ok on every rank); a cylinder ON an interface makes every rank
    fall back to single updates.  Either way: populations bitwise equal to one GPU, force series to rounding."""
    world = _world()
    out = _torchrun(world, 29691, "slab_obs_worker.py", 61, exchange, place, 4)
    assert out["ok"] and out["pop_equal"] and out["max_force_diff"] < 1e-12
    assert out["multi_ok"] == (place == "interior")
