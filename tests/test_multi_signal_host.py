"""Host-side contracts of MultiSignal that the golden episodes do not exercise (CPU, oracle backend): the `gymma` list
forms EPyMARL uses (multi_signal.py:148-155,164-168,190-196), the metrics CSV (:218-226), the arguments we refuse."""
import os

import numpy as np
import pytest

import resco_b200.rewards as rewards
import resco_b200.states as states
from resco_b200.multi_signal import MultiSignal


def _env(gymma, log_dir=None, **kw):
    from pyoracle import OracleSim
    return MultiSignal("host", "cologne3", None, states.drq_norm, rewards.wait_norm, step_length=10, yellow_length=3,
                       log_dir=log_dir, gymma=gymma, backend=lambda m: OracleSim(m, 1, seed=0), seed=11, **kw)


def test_gymma_lists_are_the_dict_view_in_ts_order():
    a, b = _env(False), _env(True)
    oa, ob = a.reset(), b.reset()
    assert isinstance(ob, list) and len(ob) == len(a.ts_order) == b.n_agents
    for i, ts in enumerate(a.ts_order):
        np.testing.assert_array_equal(oa[ts], ob[i])
    ng = [len(a.phases[ts]) for ts in a.ts_order]
    for step in range(25):
        acts = [(step // 2 + i) % n for i, n in enumerate(ng)]
        ra = a.step({ts: acts[i] for i, ts in enumerate(a.ts_order)})
        rb = b.step(acts)                                   # EPyMARL passes a list in ts_order
        assert isinstance(rb[0], list) and isinstance(rb[1], list) and rb[2] == [ra[2]] and rb[3] == ra[3]
        for i, ts in enumerate(a.ts_order):
            np.testing.assert_array_equal(ra[0][ts], rb[0][i])
            assert float(ra[1][ts]) == float(rb[1][i])
    assert [sp.n for sp in b.action_space] == ng
    assert [tuple(sp.shape) for sp in b.observation_space] == [tuple(a.obs_shape[ts]) for ts in a.ts_order]
    a.close(); b.close()


def test_metrics_csv_lines_parse_like_the_reference_reader(tmp_path):
    env = _env(False, log_dir=str(tmp_path))
    env.reset()
    ng = [len(env.phases[ts]) for ts in env.ts_order]
    for step in range(6):
        env.step({ts: step % ng[i] for i, ts in enumerate(env.ts_order)})
    env.metrics_saved = list(env.metrics)
    env.reset()                                             # the reference saves metrics_<run>.csv on the next reset
    path = os.path.join(str(tmp_path), env.connection_name, "metrics_1.csv")
    lines = open(path).read().splitlines()
    assert len(lines) == 6
    for n, line in enumerate(lines):
        assert float(line.split(",")[0]) == env._begin + 10 * (n + 1)
        # the reference's reader (utils/readCSV.py:31-40): third '}'-separated field = queue_lengths, ':'-split per signal
        queues = line.split('}')[2]
        signals = queues.split(':')
        got = [int(sig.split(',')[0]) for sig in signals[1:]]
        assert got == [env.metrics_saved[n]['queue_lengths'][ts] for ts in env.ts_order]
    env.close()


def test_unsupported_arguments_are_refused_loudly():
    with pytest.raises(NotImplementedError):
        _env(False, warmup=5)
    with pytest.raises(NotImplementedError):
        _env(False, step_ratio=4)
    with pytest.raises(ValueError):
        _env(False, lights=["not-a-signal"])
    with pytest.raises(FileNotFoundError):
        MultiSignal("host", "no_such_map", None, states.mplight, rewards.wait, log_dir=None)


@pytest.mark.parametrize("map_name,key,rkey", [("cologne8", "fma2c", "fma2c"), ("ingolstadt7", "fma2c_full", "fma2c_full"),
                                               ("cologne3", "drq_norm", "wait_norm"), ("cologne8", "mplight", "pressure"),
                                               ("cologne1", "wave", "wait"), ("ingolstadt21", "drq", "pressure"),
                                               ("cologne8", "mplight_full", "pressure")])
def test_batched_callables_match_the_dict_view_on_the_oracle(map_name, key, rkey):
    """`.batched(env)` device tensors (kernel outputs; static gather plans for FMA2C; arrivals / departures from
    lane_arrivals and presence_counts) against the per-instance dict callables the reference goldens pin, N = 3."""
    from pyoracle import OracleSim
    batched_vs_dict(map_name, key, rkey, lambda m: OracleSim(m, 3, seed=0))


def batched_vs_dict(map_name, key, rkey, backend):
    import util
    sc = util.load(map_name)
    mc = sc.meta["map_config"]
    sfn = getattr(states, key)
    rfn = getattr(rewards, rkey)
    n_env = 3
    env = MultiSignal("host", map_name, None, sfn, rfn, step_length=mc["step_length"], yellow_length=mc["yellow_length"],
                      log_dir=None, n_env=n_env, seed=5, backend=backend)
    ng = np.array([len(env.phases[ts]) for ts in env.signal_ids])
    for inst in (0, 2):
        env.run = 0
        obs_b = env.reset()
        env._refresh_views(inst)
        _cmp_obs(env, key, obs_b, sfn(env.signals), inst, "reset")
        for step in range(25):
            act = ((step // 2 + np.arange(n_env)[:, None] + np.arange(len(ng))[None, :]) % ng[None, :]).astype(np.int32)
            obs_b, rew_b, done, _ = env.step(act)
            env._refresh_views(inst)
            _cmp_obs(env, key, obs_b, sfn(env.signals), inst, f"step {step}")
            ref_r = rfn(env.signals)
            for i, k in enumerate(ref_r):
                got = rew_b[k][inst] if isinstance(rew_b, dict) else rew_b[inst, i]
                np.testing.assert_allclose(float(got), float(ref_r[k]), rtol=1e-5, atol=1e-4, err_msg=f"step {step} reward {k}")
    env.close()


def _cmp_obs(env, key, got, ref, inst, ctx):
    if isinstance(got, dict):
        assert list(got.keys()) == list(ref.keys())
        for k in ref:
            np.testing.assert_allclose(got[k][inst].cpu().numpy(), ref[k], rtol=1e-5, atol=1e-6, err_msg=f"{ctx} obs {k}")
    elif key.startswith("drq"):      # [N, n_sig_lanes, 5] rows, signal-major; the dict view is [1, lanes, 5] per signal
        for s, ts in enumerate(env.signal_ids):
            np.testing.assert_allclose(got[inst, env.sig_lane_slices[s]].cpu().numpy(), np.asarray(ref[ts])[0], rtol=1e-5,
                                       atol=1e-6, err_msg=f"{ctx} obs {ts}")
    else:                            # mplight / wave: [N, S, 13 | 12]
        for s, ts in enumerate(env.signal_ids):
            np.testing.assert_allclose(got[inst, s].cpu().numpy(), np.asarray(ref[ts]), rtol=1e-5, atol=1e-5, err_msg=f"{ctx} obs {ts}")


def test_epymarl_registration_mirrors_the_reference(tmp_path):
    """resco_benchmark/__init__.py:16-61: 18 algorithms x 8 maps x 29 trials, gymma list forms, and the registered
    kwargs construct an environment (here on the oracle) that steps."""
    import resco_b200
    from pyoracle import OracleSim
    got = {}
    ids = resco_b200.register_epymarl(register=lambda id, entry_point, kwargs: got.__setitem__(id, (entry_point, kwargs)),
                                      log_dir=str(tmp_path))
    assert len(ids) == len(got) == 18 * 8 * 29
    ep, kw = got["cologne3-qmix_ns-v7"]
    assert ep == "resco_b200.multi_signal:MultiSignal"
    assert kw["run_name"] == "qmix_ns-tr7" and kw["gymma"] is True and kw["yellow_length"] == 4 and kw["step_length"] == 10
    assert kw["state_fn"] is states.drq_norm and kw["reward_fn"] is rewards.wait_norm
    env = MultiSignal(**kw, backend=lambda m: OracleSim(m, 1, seed=0), seed=1)
    obs = env.reset()
    assert isinstance(obs, list) and len(obs) == env.n_agents
    obs, rew, done, info = env.step([0] * env.n_agents)
    assert isinstance(rew, list) and done == [False] and info == {'eps': 1}
    env.close()
