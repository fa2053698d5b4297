"""SURVEY 8(e) on real devices: instances shard by contiguous global id, one process per GPU over NCCL, and the results
do not depend on the number of ranks (the per-instance random streams are keyed by the GLOBAL instance id).  Needs two
GPUs on the box (`gpurun --gpus 2 -- python -m pytest tests/test_multi_gpu.py -m gpu`); skipped on a one-GPU box.  The
host-side sharding logic is covered without a GPU by tests/test_dist_gloo.py."""
import os
import sys

import numpy as np
import pytest
import torch

import util

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N_LOCAL, STEPS, SEED = 37, 40, 21          # 37: not a multiple of the eight instances a CTA steps


def _worker(rank, world, port, out):
    import torch.distributed as dist
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{rank}"))
    from resco_b200.parallel import allgather_obs, shard_range
    from resco_b200.sim import VecSim
    sc, m = util.marshal_map("cologne8", vcap=1024, tile_vcap=128)
    first, n = shard_range(world * N_LOCAL, world, rank)
    sim = VecSim(m, n, seed=SEED, device=rank)
    sim.reset(SEED, first)
    sim.observe()
    sim.policy_maxpressure(sc.meta["phase_pairs"], sc.meta["valid_acts"], m.info["signal_ids"])   # uploads the tables
    for _ in range(STEPS):
        sim.env_step_policy("maxpressure")
        gathered = allgather_obs(sim.obs_view()["mplight"])       # the learner's view of the whole batch, every step
    torch.cuda.synchronize()
    if rank == 0:
        np.save(out, gathered.cpu().numpy())
    dist.destroy_process_group()


@pytest.mark.gpu
def test_two_gpus_match_one(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs on the box")
    import torch.multiprocessing as mp
    from resco_b200.sim import VecSim
    out = str(tmp_path / "g.npy")
    mp.spawn(_worker, args=(2, 29547, out), nprocs=2, join=True)
    got = np.load(out)
    sc, m = util.marshal_map("cologne8", vcap=1024, tile_vcap=128)
    sim = VecSim(m, 2 * N_LOCAL, seed=SEED, device=0)
    sim.reset(SEED, 0)
    sim.observe()
    sim.policy_maxpressure(sc.meta["phase_pairs"], sc.meta["valid_acts"], m.info["signal_ids"])
    for _ in range(STEPS):
        sim.env_step_policy("maxpressure")
    want = sim.obs()["mplight"]
    assert got.shape == want.shape == (2 * N_LOCAL, 8, 13)
    assert np.array_equal(got, want)
    assert want[:, :, 1:].any()
