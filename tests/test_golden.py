"""Golden vectors produced by the UNMODIFIED reference Python (tools/make_golden.py) vs our MultiSignal.

CPU (`not gpu`): our MultiSignal drives the oracle's fused env-step path -> pins the oracle's RESCO-layer
restatement (yellows, phase machine, observe latch, states, rewards, metrics, step schedule) against the
reference itself.  GPU: the same goldens through the CUDA path (C-ABI)."""
import glob
import json
import os

import numpy as np
import pytest

import resco_b200.rewards as rewards
import resco_b200.states as states
from resco_b200.multi_signal import MultiSignal

GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))


def _oracle_backend(m):
    from pyoracle import OracleSim
    return OracleSim(m, 1, seed=0)


def _run_case(path, backend, tile_vcap=0):
    z = np.load(path)
    meta = json.loads(bytes(z["meta"]).decode())
    env = MultiSignal("golden", meta["map"], None, getattr(states, meta["state"]), getattr(rewards, meta["reward"]),
                      step_length=meta["step_length"], yellow_length=meta["yellow_length"],
                      max_distance=meta["max_distance"], log_dir=None, backend=backend, tile_vcap=tile_vcap)
    order = meta["ts_order"]
    assert env.ts_order == order
    assert {ts: list(env.obs_shape[ts]) for ts in order} == meta["obs_shapes"]
    real = [ts for ts in order if ts in env.signals]      # manager pseudo-agents (FMA2C) are not signals
    for ts in real:           # create_yellows + installed program (traffic_signal.py:7-24,93-100)
        assert env.signals[ts].yellow_dict == meta["yellow_dicts"][ts]
        assert [s for _, s in env.signals[ts].phases] == [s for _, s in meta["programs"][ts]]
        assert env.signals[ts].lanes == meta["lanes"][ts]
        assert len(env.phases[ts]) == meta["n_actions"][ts]
    obs = env.reset()
    got = np.concatenate([np.asarray(obs[ts], np.float64).ravel() for ts in order])
    np.testing.assert_allclose(got, z["reset_obs"], rtol=1e-9, atol=1e-12)
    for step in range(z["act"].shape[0]):
        act = {ts: int(z["act"][step, i]) for i, ts in enumerate(order) if ts in env.signals}
        obs, rew, done, info = env.step(act)
        got = np.concatenate([np.asarray(obs[ts], np.float64).ravel() for ts in order])
        np.testing.assert_allclose(got, z["obs"][step], rtol=1e-6, atol=1e-9, err_msg=f"obs step {step}")
        np.testing.assert_allclose([float(rew[ts]) for ts in order], z["rew"][step], rtol=1e-6, atol=1e-9,
                                   err_msg=f"reward step {step}")
        assert [env.signals[ts].phase if ts in env.signals else -1 for ts in order] == z["phase"][step].tolist(), f"phase step {step}"
        mt = env.metrics[-1]
        assert [mt["queue_lengths"].get(ts, -1) for ts in order] == z["queue_lengths"][step].tolist()
        assert [mt["max_queues"].get(ts, -1) for ts in order] == z["max_queues"][step].tolist()
        assert mt["step"] == z["step_time"][step]
    assert int(env.sim.stats()["n_cap_refused"][0]) == 0      # the vehicle store never truncated the episode
    env.close()


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[:-4] for p in GOLD])
def test_reference_golden_cpu(path):
    _run_case(path, _oracle_backend)


@pytest.mark.gpu
@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[:-4] for p in GOLD])
def test_reference_golden_gpu(path):
    _run_case(path, None)      # default backend = CUDA VecSim through the C-ABI


GOLD_C8 = [p for p in GOLD if os.path.basename(p).startswith("cologne8_")]


@pytest.mark.gpu
@pytest.mark.parametrize("path", GOLD_C8, ids=[os.path.basename(p)[:-4] for p in GOLD_C8])
def test_reference_golden_gpu_bench_tile(path):
    """the cologne8 goldens again on the 128-vehicle tile bench.py times (launch shape k_run<64, 8, 1>; the episodes of
    random actions jam the map beyond the tile: those steps go through the overflow pass)"""
    _run_case(path, None, tile_vcap=128)


def test_goldens_present():
    assert len(GOLD) >= 8


@pytest.mark.parametrize("case", ["cologne1_mplight_wait_maxpressure", "cologne8_mplight_wait_maxpressure"])
def test_host_wave_agent_reproduces_reference_actions(case):
    """rs_host_agent_wave (host-side agent front-end of the C library, no device work) over the observations the
    reference recorded picks the actions the reference's own MAXPRESSURE agent picked (agents/maxpressure.py)."""
    from resco_b200.sim import HostWaveAgent
    from util import load
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", case + ".npz"))
    meta = json.loads(bytes(z["meta"]).decode())
    sc = load(meta["map"])
    order = meta["ts_order"]
    agent = HostWaveAgent(sc.meta["phase_pairs"], sc.meta["valid_acts"], order)
    obs = np.concatenate([z["reset_obs"][None], z["obs"][:-1]]).reshape(-1, len(order), 13).astype(np.float32)
    np.testing.assert_array_equal(agent.act(obs), z["act"])
