"""Statistical anchors: episode avg delay (utils/readXML.py:38-76 definition) of the rule set, run on the CPU
oracle, next to the reference's published per-episode means (utils/avg_timeLoss.py).  TEST INFRASTRUCTURE
(uses oracle/); bands, not parity -- the reference's runs are --random seeded SUMO runs.

usage: python tools/anchors.py [map ...] [--seeds N]
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import util  # noqa: E402
from oracle.pyoracle import OracleSim  # noqa: E402

# reference means: utils/avg_timeLoss.py
REF = {("grid4x4", "MAXWAVE"): 34.3, ("grid4x4", "MAXPRESSURE"): 52.6,
       ("arterial4x4", "MAXWAVE"): 820.8, ("arterial4x4", "MAXPRESSURE"): 952.7,
       ("ingolstadt1", "FIXED"): 39.4, ("ingolstadt1", "MAXWAVE"): 28.3, ("ingolstadt1", "MAXPRESSURE"): 23.6,
       ("ingolstadt7", "FIXED"): 91.3, ("ingolstadt7", "MAXWAVE"): 80.6, ("ingolstadt7", "MAXPRESSURE"): 46.8,
       ("ingolstadt21", "FIXED"): 133.1, ("ingolstadt21", "MAXWAVE"): 76.3, ("ingolstadt21", "MAXPRESSURE"): 136.7,
       ("cologne1", "FIXED"): 56.6, ("cologne1", "MAXWAVE"): 27.8, ("cologne1", "MAXPRESSURE"): 31.9,   # MP: first episode
       ("cologne3", "FIXED"): 46.4, ("cologne3", "MAXWAVE"): 21.1, ("cologne3", "MAXPRESSURE"): 24.7,   # first episodes
       ("cologne8", "FIXED"): 63.8, ("cologne8", "MAXWAVE"): 21.9, ("cologne8", "MAXPRESSURE"): 28.8}   # MP: first episode


def delay_all(st):
    """every trip whose departure time has passed, inserted or not (NOT the reference's metric: see episode_delay)"""
    n = st["n_arrived"] + st["n_active"] + st["n_backlog"]
    return (st["sum_delay_arrived"] + st["sum_delay_running"] + st["sum_delay_pending"]) / np.maximum(n, 1)


def episode_delay(sim, sc, m, env, tmpdir):
    """The reference's per-episode number (utils/readXML.py:27-77): write the tripinfo file of instance `env` the way
    SUMO would (--tripinfo-output.write-unfinished) and average timeLoss + departDelay over its entries.  Trips that
    were never inserted are NOT in a tripinfo file; readXML charges them (end_time - depart) only for <vehicle>-type
    route files (grid4x4 / arterial4x4) and only those scheduled after the last vehicle that did depart."""
    from resco_b200.metrics import avg_delay_from_tripinfo, write_tripinfo
    path = os.path.join(tmpdir, f"tripinfo_{env}.xml")
    tick = int(sim.stats()["tick"][env])
    write_tripinfo(path, sc, sim.trip_records(env), sim.vehicles(env), tick)
    vehicle_demand = sc.meta["map_name"] in ("grid4x4", "arterial4x4")
    return avg_delay_from_tripinfo(path, sc, end_time=float(sc.meta["map_config"]["end_time"]), vehicle_demand=vehicle_demand)


def run(map_name, policy, seeds, vcap=4096):
    sc = util.load(map_name)
    mc = sc.meta["map_config"]
    T = int(mc["end_time"] - mc["start_time"])
    if policy == "FIXED":
        m = util.marshal_map(map_name, controlled=False, vcap=vcap, record_trips=True)[1]
    else:
        m = util.marshal_map(map_name, vcap=vcap, max_distance=50.0 if policy == "MAXWAVE" else 200.0, record_trips=True)[1]
    o = OracleSim(m, seeds, seed=1)
    o.reset(1, 0)
    if policy == "FIXED":
        o.tick(T)
    else:
        act = np.zeros((seeds, m.struct.n_signals), np.int32)
        o.observe()
        for _ in range(T // m.struct.step_length):
            o.env_step(act)
            ob = o.obs()
            x = ob["mplight"] if policy == "MAXPRESSURE" else np.concatenate([ob["mplight"][:, :, :1], ob["wave"]], 2)
            act = util.maxpressure_actions(sc, m, x)
    st = o.stats()
    import tempfile
    with tempfile.TemporaryDirectory() as td:
        d = np.asarray([episode_delay(o, sc, m, e, td) for e in range(seeds)])
    return d, st


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("maps", nargs="*", default=["cologne1", "cologne3", "cologne8", "ingolstadt1", "ingolstadt7", "ingolstadt21", "grid4x4", "arterial4x4"])
    ap.add_argument("--seeds", type=int, default=4)
    ap.add_argument("--policies", default="FIXED,MAXPRESSURE,MAXWAVE")
    a = ap.parse_args()
    for mp in a.maps:
        for pol in a.policies.split(","):
            if pol == "FIXED" and mp in ("grid4x4", "arterial4x4"):
                continue  # not applicable in the reference either (utils/graph.py:91)
            t0 = time.time()
            d, st = run(mp, pol, a.seeds)
            print(f"{mp:13s} {pol:12s} delay mean {d.mean():7.1f} [{d.min():6.1f}..{d.max():6.1f}]  ref {REF.get((mp, pol))}  "
                  f"arrived {st['n_arrived'].mean():.0f} active {st['n_active'].mean():.0f} backlog {st['n_backlog'].mean():.0f} "
                  f"anom {st['anomalies'].sum()}  ({time.time() - t0:.1f}s)", flush=True)
