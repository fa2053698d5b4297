"""tripinfo_<run>.xml writer + the avg-delay restatement of utils/readXML.py:27-77."""
import os
import xml.etree.ElementTree as ET

import numpy as np
import pytest

import resco_b200.rewards as rewards
import resco_b200.states as states
import util
from resco_b200.metrics import avg_delay_from_tripinfo
from resco_b200.multi_signal import MultiSignal


def _episode(backend, tmp_path, steps=120):
    env = MultiSignal("t", "cologne1", None, states.mplight, rewards.wait, step_length=10, yellow_length=3,
                      max_distance=200, log_dir=str(tmp_path), backend=backend)
    env.reset()
    rng = np.random.default_rng(0)
    for _ in range(steps):
        env.step({ts: int(rng.integers(len(env.phases[ts]))) for ts in env.ts_order})
    path = env.save_tripinfo()
    st = env.sim.stats()[0]
    return env, path, st


def _check(env, path, st):
    root = ET.parse(path).getroot()
    entries = list(root)
    assert len(entries) == st["n_arrived"] + st["n_active"]
    assert sum(1 for e in entries if float(e.get("arrival")) < 0) == st["n_active"]
    for e in entries[:50]:
        assert float(e.get("duration")) >= 0 and float(e.get("timeLoss")) >= 0 and float(e.get("departDelay")) >= 0
    # readXML.py average == the device accumulators (trip-type demand: no never-departed term)
    want = (float(st["sum_delay_arrived"]) + float(st["sum_delay_running"])) / (st["n_arrived"] + st["n_active"])
    got = avg_delay_from_tripinfo(path)
    assert abs(got - want) < 0.02, (got, want)          # XML carries 2 decimals
    # waitingTime (utils/readXML.py metric 'waitingTime' -> utils/avg_waitingTime.py): seconds with speed < 0.1 m/s;
    # whole seconds, never longer than the trip, and the finished trips add up to RsStats.sum_wait_arrived
    waits = np.array([float(e.get("waitingTime")) for e in entries])
    durs = np.array([float(e.get("duration")) for e in entries])
    assert (waits == np.round(waits)).all() and (waits <= durs + 1e-6).all() and waits.max() > 0
    done = np.array([float(e.get("arrival")) >= 0 for e in entries])
    assert abs(waits[done].sum() - float(st["sum_wait_arrived"])) < 1e-3
    assert abs(avg_delay_from_tripinfo(path, metric="waitingTime") - waits.mean()) < 1e-9


def test_tripinfo_cpu(tmp_path):
    from pyoracle import OracleSim
    env, path, st = _episode(lambda m: OracleSim(m, 1, seed=0), tmp_path)
    _check(env, path, st)
    assert os.path.exists(os.path.join(env._log_path(), "tripinfo_1.xml"))


@pytest.mark.gpu
def test_tripinfo_gpu_matches_oracle(tmp_path):
    from pyoracle import OracleSim
    env_g, path_g, st_g = _episode(None, tmp_path / "g")
    env_o, path_o, st_o = _episode(lambda m: OracleSim(m, 1, seed=0), tmp_path / "o")
    _check(env_g, path_g, st_g)
    rg, ro = env_g.sim.trip_records(0), env_o.sim.trip_records(0)
    for k in rg:
        assert np.array_equal(rg[k][ro["arrival"] >= 0], ro[k][ro["arrival"] >= 0]), k
    assert open(path_g).read() == open(path_o).read()


def test_never_departed_rule(tmp_path):
    """<vehicle>-type demand: vehicles scheduled after the last departed one are charged end_time - depart."""
    sc = util.load("grid4x4")
    from resco_b200.metrics import episode_trip_range
    t0, t1 = episode_trip_range(sc, 0)          # the scenario holds several route files back to back: episode 0
    ids = sc.meta["trip_ids"][t0:t1]
    sched = sc.meta["begin"] + sc.arrays["trip_depart"][t0:t1].astype(float)
    order = np.argsort(sched, kind="stable")
    a, b = int(order[0]), int(order[len(order) // 2])
    p = tmp_path / "t.xml"
    p.write_text('<tripinfos>\n'
                 f'<tripinfo id="{ids[a]}" depart="{sched[a] + 2:.2f}" departDelay="2.00" arrival="90.00" duration="80.00" timeLoss="10.00" waitingTime="0.00"/>\n'
                 f'<tripinfo id="{ids[b]}" depart="{sched[b]:.2f}" departDelay="0.00" arrival="-1.00" duration="5.00" timeLoss="1.00" waitingTime="0.00"/>\n'
                 '</tripinfos>\n')
    never = sched[sched > sched[b]]
    want = (10 + 2 + 1 + 0 + np.sum(3600.0 - never)) / (2 + len(never))
    got = avg_delay_from_tripinfo(str(p), sc, end_time=3600.0, vehicle_demand=True)
    assert abs(got - want) < 1e-9
    assert abs(avg_delay_from_tripinfo(str(p)) - (13.0 / 2)) < 1e-12


def test_waiting_time_is_the_count_of_halted_seconds():
    """the per-vehicle accumulator behind tripinfo waitingTime: +1 for every tick the vehicle ends below 0.1 m/s"""
    from pyoracle import OracleSim
    sc, m = util.marshal_map("cologne1", controlled=False)
    o = OracleSim(m, 1, seed=4)
    o.reset(4, 0)
    halted, before = {}, set()
    for _ in range(600):
        o.tick(1)
        v = o.vehicles(0)
        for vid, sp in zip(v["vid"], v["speed"]):
            if sp < 0.1 and int(vid) in before:      # a vehicle makes its first move in the tick AFTER its insertion
                halted[int(vid)] = halted.get(int(vid), 0) + 1
        before = set(int(x) for x in v["vid"])
    v = o.vehicles(0)
    assert len(v["vid"]) > 5
    for vid, aw in zip(v["vid"], v["acc_wait"]):
        assert aw == halted.get(int(vid), 0)


def test_capacity_refusals_are_counted():
    """RsStats.n_cap_refused: zero with room to spare, positive and equal to the per-tick refusals when the vehicle
    store is far too small; the refused trips stay in the backlog (insertion is put off, never dropped)."""
    from pyoracle import OracleSim
    sc, m = util.marshal_map("cologne8", controlled=False)
    o = OracleSim(m, 1, seed=1); o.reset(1, 0); o.tick(900)
    assert o.stats()["n_cap_refused"][0] == 0
    sc, m = util.marshal_map("cologne8", controlled=False, vcap=16)
    o = OracleSim(m, 1, seed=1); o.reset(1, 0); o.tick(900)
    st = o.stats()
    assert st["n_cap_refused"][0] > 0 and st["n_active"][0] <= 16 and st["n_backlog"][0] > 0
    assert st["n_inserted"][0] == st["n_arrived"][0] + st["n_active"][0]
