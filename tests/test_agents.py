"""Batched agent front ends against vectors recorded from the reference's own classes (tools/make_golden_frap.py)."""
import glob
import json
import os

import numpy as np
import pytest
import torch

import util
from resco_b200.agents import BatchedFRAP, competition_mask

FIX = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "agents", "frap_*.npz")))


def _load(path):
    z = np.load(path)
    meta = json.loads(bytes(z["meta"]).decode())
    sc = util.load(meta["map"])
    model = BatchedFRAP(sc.meta["phase_pairs"], demand_shape=1, chunk_rows=100)    # odd chunk size: exercises chunking
    sd = {k[len("param."):]: torch.from_numpy(z[k]) for k in z.files if k.startswith("param.")}
    model.load_state_dict(sd)            # strict: the reference's parameter names and shapes
    return z, meta, sc, model


@pytest.mark.parametrize("path", FIX, ids=[os.path.basename(p)[:-4] for p in FIX])
def test_batched_frap_reproduces_the_reference_forward(path):
    z, meta, sc, model = _load(path)
    with torch.no_grad():
        q = model(torch.from_numpy(z["states"]))                                  # [B, S, 13] in one call
    assert q.shape == z["q"].shape
    np.testing.assert_allclose(q.numpy(), z["q"], rtol=2e-5, atol=2e-5)


@pytest.mark.parametrize("path", FIX, ids=[os.path.basename(p)[:-4] for p in FIX])
def test_batched_frap_greedy_actions_follow_valid_acts(path):
    z, meta, sc, model = _load(path)
    va = {sid: sc.meta["valid_acts"][sid] for sid in meta["signal_ids"]}
    acts = model.act(torch.from_numpy(z["states"]), va, meta["signal_ids"])
    np.testing.assert_array_equal(acts.numpy(), z["acts"])


def test_competition_mask_and_fixture_presence():
    assert len(FIX) == 2
    m = competition_mask([[0, 1], [1, 2], [3, 4]])
    assert m.tolist() == [[1, 0], [1, 0], [0, 0]]


# ---- the same front ends as CUDA kernels of the library (resco_b200/csrc/agents.cuh) ---------------------------------
def _frap_sim(path, n_env):
    from resco_b200.sim import VecSim
    z = np.load(path)
    meta = json.loads(bytes(z["meta"]).decode())
    sc, m = util.marshal_map(meta["map"])
    sim = VecSim(m, n_env, seed=3)
    sd = {k[len("param."):]: z[k] for k in z.files if k.startswith("param.")}
    sim.load_frap(sd, sc.meta["phase_pairs"], sc.meta["valid_acts"], m.info["signal_ids"])
    return z, meta, sc, m, sim


@pytest.mark.gpu
@pytest.mark.parametrize("path", FIX, ids=[os.path.basename(p)[:-4] for p in FIX])
def test_frap_kernel_reproduces_the_reference_forward(path):
    """rs_policy_frap over the rows the reference's FRAP module was recorded on: Q-values within 2e-5 (fp32, different
    summation order), and -- where the fixture's signals are the sim's -- the reference's greedy valid actions."""
    z, meta, sc, m, sim = _frap_sim(path, 4)
    S = sim.S
    rows = z["states"].reshape(-1, 13)
    n = (rows.shape[0] // S) * S
    obs = torch.from_numpy(rows[:n].reshape(-1, S, 13)).cuda()
    acts, q = sim.policy_frap(obs, want_q=True)
    torch.cuda.synchronize()
    np.testing.assert_allclose(q.cpu().numpy().reshape(n, -1), z["q"].reshape(-1, z["q"].shape[-1])[:n], rtol=2e-5, atol=2e-5)
    if meta["signal_ids"] == list(m.info["signal_ids"]):
        np.testing.assert_array_equal(acts.cpu().numpy(), z["acts"])
    sim.close()


@pytest.mark.gpu
@pytest.mark.parametrize("policy", ["maxpressure", "maxwave", "frap", "random"])
def test_graph_step_equals_policy_then_step(policy):
    """rs_env_step_policy (policy kernel + fused env step replayed as one CUDA graph) against the two ordinary calls, and
    the device policies against their definitions: 40 agent steps on 37 cologne8 instances (128-vehicle tile)."""
    from resco_b200.sim import VecSim
    path = [p for p in FIX if "cologne8" in p][0]
    z = np.load(path)
    sc, m = util.marshal_map("cologne8", tile_vcap=128)
    pairs, va, sig = sc.meta["phase_pairs"], sc.meta["valid_acts"], m.info["signal_ids"]
    sd = {k[len("param."):]: z[k] for k in z.files if k.startswith("param.")}
    a, b = VecSim(m, 37, seed=9), VecSim(m, 37, seed=9)
    for s in (a, b):
        s.reset(9, 100); s.observe(); s.load_frap(sd, pairs, va, sig)
    ng = util.n_green(m)
    for step in range(40):
        if policy in ("maxpressure", "maxwave"):
            act = a.policy_maxpressure(pairs, va, sig, use_wave=policy == "maxwave")
            if step == 0:
                b.policy_maxpressure(pairs, va, sig, use_wave=policy == "maxwave")      # uploads the tables
            ob = a.obs()
            x = ob["mplight"] if policy == "maxpressure" else np.concatenate([ob["mplight"][:, :, :1], ob["wave"]], 2)
            assert np.array_equal(act.cpu().numpy(), util.maxpressure_actions(sc, m, x))
        elif policy == "frap":
            act = a.policy_frap()
            own = a.policy_frap(a.obs_view()["mplight"].clone())
            assert torch.equal(act, own)
        else:
            act = a.policy_random(seed=5)
            an = act.cpu().numpy()
            assert (an >= 0).all() and (an < ng[None, :]).all() and len(np.unique(an, axis=0)) > 30
        a.env_step(act)
        b.env_step_policy(policy, seed=5)
        oa, ob_ = a.obs(), b.obs()
        for k in util.OBS_EXACT:
            assert np.array_equal(oa[k], ob_[k]), (policy, step, k)
    util.assert_same_stats(a.stats(), b.stats(), policy)
    a.close(); b.close()


@pytest.mark.gpu
def test_random_policy_is_keyed_by_global_instance_id():
    from resco_b200.sim import VecSim
    sc, m = util.marshal_map("cologne8")
    a, b = VecSim(m, 16, seed=1), VecSim(m, 6, seed=1)
    a.reset(1, 0); b.reset(1, 10)
    ra, rb = a.policy_random(7).cpu().numpy(), b.policy_random(7).cpu().numpy()
    assert np.array_equal(ra[10:16], rb)
    a.tick(1)
    assert not np.array_equal(a.policy_random(7).cpu().numpy(), ra)      # keyed by the instance's tick too
    a.close(); b.close()
