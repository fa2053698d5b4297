"""departPos="random_free" (arterial4x4's route files: `<vehicle ... departPos="random_free">`): SUMO tries ten uniformly
drawn positions on the departure lane for one where the vehicle fits, then a free insertion (MSLane::insertVehicle,
RANDOM_FREE; restated in oracle/microsim.c step 6, DESIGN.md §4.2 rule 7).  The other seven maps use departPos="base".
The CUDA side of the rule is covered bit for bit by tests/test_gpu_parity.py (arterial4x4 cases)."""
import ctypes as C

import numpy as np

import util
from pyoracle import OracleSim

STREAM_DEPARTPOS = 5


def _uniform(lib, seed, env_id, vid, tick, k):
    """the k-th draw of trip `vid` at `tick` (Philox keyed like every other stream of the model)"""
    c = (C.c_uint32 * 4)(env_id & 0xFFFFFFFF, env_id >> 32, vid, (tick * 4 + (k >> 2)) & 0xFFFFFFFF)
    lib.orc_philox(c, (seed & 0xFFFFFFFF) ^ ((STREAM_DEPARTPOS * 0x632BE5AB) & 0xFFFFFFFF), seed >> 32)
    return np.float32(c[k & 3] >> 8) * np.float32(1.0 / 16777216.0)


MAPS = ["cologne1", "cologne8", "ingolstadt7", "grid4x4", "arterial4x4"]


def test_compiled_flag():
    for name in MAPS:
        a = util.load(name).arrays
        assert len(a["trip_depart_pos"]) == len(a["trip_depart"])
        assert bool(a["trip_depart_pos"].all()) == (name == "arterial4x4") and bool(a["trip_depart_pos"].any()) == (name == "arterial4x4")


def test_first_vehicles_stand_at_their_first_draw(oracle_lib):
    """Known answer: on an empty lane the first drawn position always fits, so the vehicles that depart at tick 0 stand at
    len + u * (lane length - len) with u the first Philox draw of (seed, instance id, trip, tick)."""
    sc, m = util.marshal_map("arterial4x4")
    a = sc.arrays
    seed, first = 11, 40
    o = OracleSim(m, 3, seed=seed)
    o.reset(seed, first)
    o.tick(1)
    seen = 0
    for e in range(3):
        v = o.vehicles(e)
        assert len(v["vid"]) > 0 and (v["depart"] == 0).all() and (v["speed"] == 0).all()
        for lane, pos, vid, vt in zip(v["lane"], v["pos"], v["vid"], v["vtype"]):
            ln = np.float32(a["vtype"][vt, 0])      # VT_LEN is the first column of the vType table
            L = np.float32(a["lane_len"][lane])
            u = _uniform(oracle_lib, seed, first + e, int(vid), 0, 0)
            assert pos == np.float32(ln + np.float32(u * np.float32(L - ln))), (e, vid)
            assert ln <= pos <= L
            seen += 1
    assert seen >= 6
    # another instance id / another seed: another position for the same trip
    p0 = {int(k): float(p) for k, p in zip(o.vehicles(0)["vid"], o.vehicles(0)["pos"])}
    p1 = {int(k): float(p) for k, p in zip(o.vehicles(1)["vid"], o.vehicles(1)["pos"])}
    assert p0.keys() == p1.keys() and any(p0[k] != p1[k] for k in p0)


def test_base_maps_insert_at_the_lane_start():
    sc, m = util.marshal_map("cologne8")
    o = OracleSim(m, 1, seed=2)
    o.reset(2, 0)
    for _ in range(40):
        o.tick(1)
        v = o.vehicles(0)
        new = v["depart"] == o.stats()["tick"][0] - 1
        assert (v["pos"][new] == sc.arrays["vtype"][v["vtype"][new], 0]).all()


def test_insertion_between_vehicles_keeps_the_lane_order():
    """Under load vehicles are inserted between others: every lane stays sorted front to back, nobody overlaps, and some
    newcomer is not the last vehicle of its lane (which departPos="base" can never produce)."""
    sc, m = util.marshal_map("arterial4x4")
    vt = sc.arrays["vtype"]
    o = OracleSim(m, 2, seed=5)
    o.reset(5, 0)
    o.observe()
    mid = 0
    for step in range(150):
        o.env_step(util.cyclic_actions(m, 2, step))
        for e in range(2):
            v = o.vehicles(e)
            tick = o.stats()["tick"][e]
            for lane in np.unique(v["lane"][v["depart"] >= tick - 5]):
                sel = np.nonzero(v["lane"] == lane)[0]
                pos, ln = v["pos"][sel], vt[v["vtype"][sel], 0]
                assert (np.diff(pos) <= 0).all()
                assert (pos[:-1] - ln[:-1] - pos[1:] >= 0).all(), "a newcomer overlaps its leader"
                fresh = v["depart"][sel] >= tick - 5
                mid += int(fresh[:-1].any())
    st = o.stats()
    assert (st["anomalies"] == 0).all() and mid > 0
