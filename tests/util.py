"""Shared helpers for the test-suite (scenario loading, policies, comparisons)."""
import os

import numpy as np

from resco_b200.abi import marshal
from resco_b200.scenario import Scenario

DATA = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "resco_b200", "data")


def load(map_name):
    return Scenario.load(os.path.join(DATA, map_name + ".npz"))


def marshal_map(map_name, **kw):
    sc = load(map_name)
    mc = sc.meta["map_config"]
    kw.setdefault("step_length", mc["step_length"])
    kw.setdefault("yellow_length", mc["yellow_length"])
    return sc, marshal(sc, **kw)


def n_green(m):
    return np.array([len(m.info["green_states"][s]) for s in m.info["signal_ids"]], np.int32)


def cyclic_actions(m, n_env, step, period=3):
    ng = n_green(m)
    base = (step // period) % ng
    off = np.arange(n_env)[:, None]
    return ((base[None, :] + off) % ng[None, :]).astype(np.int32)


def maxpressure_actions(sc, m, mplight):
    """agents/maxwave.py:18-38 over states.mplight[1:] (agents/maxpressure.py:13-18), batched numpy."""
    pairs = sc.meta["phase_pairs"]
    va = sc.meta["valid_acts"]
    n_env = mplight.shape[0]
    sig = m.info["signal_ids"]
    act = np.zeros((n_env, len(sig)), np.int32)
    ob = mplight[:, :, 1:]
    for i, s in enumerate(sig):
        idxs = list(range(len(pairs))) if va is None else [int(k) for k in va[s].keys()]
        press = np.stack([ob[:, i, pairs[k][0]] + ob[:, i, pairs[k][1]] for k in idxs], 1)
        best = np.argmax(press, 1)       # first maximum, like the reference's strict '>' scan
        if va is None:
            act[:, i] = np.asarray(idxs)[best]
        else:
            act[:, i] = np.asarray([va[s][str(k)] for k in idxs])[best]
    return act


VEH_EXACT = ["lane", "pos", "speed", "wait", "rwait", "tloss", "vid", "vtype", "route", "cursor", "sf", "depart", "acc_wait"]
OBS_CLOSE = ["lane_speed_sum", "drq", "drq_norm", "mplight_full"]
OBS_EXACT = ["lane_queue", "lane_approach", "lane_total_wait", "lane_max_wait", "lane_arrivals", "phase", "mplight", "wave",
             "reward_wait", "reward_wait_norm", "reward_pressure", "sig_queue_len", "sig_max_queue"]


def assert_same_state(a, b, env=0, ctx=""):
    va, vb = a.vehicles(env), b.vehicles(env)
    assert len(va["pos"]) == len(vb["pos"]), f"{ctx}: vehicle count {len(va['pos'])} vs {len(vb['pos'])}"
    for k in VEH_EXACT:
        x, y = va[k], vb[k]
        if x.dtype.kind == "f":
            assert np.array_equal(x.view(np.uint32), y.view(np.uint32)), \
                f"{ctx}: vehicle field {k} differs at {np.nonzero(x != y)[0][:5]}: {x[x != y][:5]} vs {y[x != y][:5]}"
        else:
            assert np.array_equal(x, y), f"{ctx}: vehicle field {k} differs at {np.nonzero(x != y)[0][:5]}"
    assert np.array_equal(a.phases(env), b.phases(env)), f"{ctx}: tls phases differ"


def assert_same_obs(oa, ob, ctx=""):
    for k in OBS_EXACT:
        assert np.array_equal(oa[k], ob[k]), f"{ctx}: obs {k} differs: {np.argwhere(oa[k] != ob[k])[:4]}"
    # the speed sums (warp-shuffle tree vs sequential order) and what is derived from them: rtol 1e-5
    for k in OBS_CLOSE:
        if k in oa and k in ob:
            np.testing.assert_allclose(oa[k], ob[k], rtol=1e-5, atol=1e-5, err_msg=f"{ctx}: {k}")


STATS_INT = ["tick", "n_active", "n_inserted", "n_arrived", "n_backlog", "anomalies", "sum_active_ticks", "n_cap_refused"]
STATS_FLOAT = ["sum_delay_arrived", "sum_delay_running", "sum_delay_pending", "sum_duration_arrived", "sum_wait_arrived"]


def assert_same_stats(sa, sb, ctx=""):
    """episode statistics: bit-exact (float sums are accumulated in the same order on both sides)"""
    for k in STATS_INT + STATS_FLOAT:
        assert np.array_equal(sa[k], sb[k]), f"{ctx}: stats {k} differ: {sa[k][:4]} vs {sb[k][:4]}"
