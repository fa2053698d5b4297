"""N>1 path on CPU: world_size-2 gloo.  Instances shard by contiguous global id with no data-path
collective; the one collective is the observation all-gather used when a shared-policy agent (MPLight)
evaluates on one rank.  Results must not depend on the number of ranks."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import util

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n_local, steps, out):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pyoracle import OracleSim
    from resco_b200.parallel import allgather_obs, shard_range
    sc, m = util.marshal_map("cologne8")
    first, n = shard_range(world * n_local, world, rank)
    sim = OracleSim(m, n, seed=9)
    sim.reset(9, first)
    sim.observe()
    gathered = None
    for step in range(steps):
        obs_local = torch.from_numpy(sim.obs()["mplight"])
        gathered = allgather_obs(obs_local)                       # [world*n_local, S, 13] on every rank
        act_all = util.maxpressure_actions(sc, m, gathered.numpy())  # shared policy on the full batch
        sim.env_step(act_all[first:first + n])
    if rank == 0:
        np.save(out, gathered.numpy())
    dist.destroy_process_group()


def test_two_ranks_match_single_process(tmp_path):
    from pyoracle import OracleSim
    n_local, steps = 3, 25
    out = str(tmp_path / "g.npy")
    mp.spawn(_worker, args=(2, 29533, n_local, steps, out), nprocs=2, join=True)
    got = np.load(out)
    sc, m = util.marshal_map("cologne8")
    sim = OracleSim(m, 2 * n_local, seed=9)
    sim.reset(9, 0)
    sim.observe()
    last = None
    for step in range(steps):
        last = sim.obs()["mplight"]
        sim.env_step(util.maxpressure_actions(sc, m, last))
    assert np.array_equal(got, last)


def test_shard_range():
    from resco_b200.parallel import shard_range
    assert [shard_range(10, 4, r) for r in range(4)] == [(0, 3), (3, 3), (6, 2), (8, 2)]
    assert sum(shard_range(65536, 8, r)[1] for r in range(8)) == 65536
