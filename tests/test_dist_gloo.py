"""Multi-rank host logic on CPU: world_size-2 gloo processes partition the shots round-robin, pack
[grad | illum | fval], all-reduce once and finalise -- the result must equal the single-rank one."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as tdist
import torch.multiprocessing as mp

SHAPE = (30, 11)
NSHOTS = 7


def _shot_contribution(i):
    rng = np.random.default_rng(100 + i)
    return rng.standard_normal(SHAPE), rng.random(SHAPE) + 0.1, float(rng.random())


def _partial(shots):
    n = SHAPE[0] * SHAPE[1]
    buf = np.zeros(2 * n + 1)
    for i in shots:
        g, il, f = _shot_contribution(i)
        buf[:n] += g.ravel()
        buf[n:2 * n] += il.ravel()
        buf[2 * n] += f
    return buf


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    from devito_fwi_b200 import dist, fwi
    dist.init_from_env("gloo")
    assert dist.rank() == rank and dist.world_size() == world
    shots = dist.local_shots(NSHOTS)
    assert shots == list(range(rank, NSHOTS, world))
    buf = torch.from_numpy(_partial(shots))
    dist.all_reduce_sum(buf)
    mask = np.ones(SHAPE)
    mask[:, :2] = 0
    f, g = fwi._finalize_objective(buf.numpy(), SHAPE, mask, True, True)
    np.save(os.path.join(out_dir, "g%d.npy" % rank), np.concatenate([[f], g]))
    tdist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


@pytest.mark.timeout(180)
def test_two_rank_objective_equals_single_rank(tmp_path):
    from devito_fwi_b200 import fwi
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    mask = np.ones(SHAPE)
    mask[:, :2] = 0
    f1, g1 = fwi._finalize_objective(_partial(range(NSHOTS)), SHAPE, mask, True, True)
    want = np.concatenate([[f1], g1])
    for r in range(world):
        got = np.load(os.path.join(str(tmp_path), "g%d.npy" % r))
        assert np.allclose(got, want, rtol=1e-13, atol=0)       # summation order differs across ranks only
    # every rank holds the same (f, g): the replicated host optimiser stays in lock-step
    assert np.array_equal(np.load(os.path.join(str(tmp_path), "g0.npy")),
                          np.load(os.path.join(str(tmp_path), "g1.npy")))


def test_weak_scaling_survey_partition():
    """bench.py's job: 29*N shots, each rank gets the 29 distinct source positions of the reference's survey."""
    import bench
    from devito_fwi_b200 import dist
    for world in (1, 2, 4, 8):
        g_true, g_init, _, _ = bench.make_survey(world)
        assert g_init.nsrc == 29 * world
        base = np.unique(g_init.src_positions, axis=0)
        assert base.shape[0] == 29
        for r in range(world):
            mine = g_init.src_positions[dist.local_shots(g_init.nsrc, r, world)]
            assert np.array_equal(np.unique(mine, axis=0), base) and mine.shape[0] == 29
