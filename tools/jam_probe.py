"""Diagnostic (test infrastructure, uses oracle/): where does a map jam?  Runs one oracle instance and prints, at
intervals, the lane heads that have been standing longest together with the reason the model gives (orc_explain).

usage: python tools/jam_probe.py MAP [FIXED|MAXPRESSURE|MAXWAVE] [--every 600] [--top 12] [--seed 1]
"""
import argparse
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import util  # noqa: E402
from oracle.pyoracle import OracleSim, lib  # noqa: E402

REASON = {0: "free", 1: "red", 2: "yellow", 3: "stop-sign", 4: "slot taken", 7: "keep-clear", 8: "wrong lane", 9: "leader ahead", -1: "route end"}


def reason(code):
    if code >= 400000: return f"major: crossing foe link {code - 400000}"
    if code >= 300000: return f"yield to foe in junction slot, link {code - 300000}"
    if code >= 200000: return f"yield to approaching foe, link {code - 200000}"
    if code >= 100000: return f"foe crossing, link {code - 100000}"
    return REASON.get(code, str(code))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("map")
    ap.add_argument("policy", nargs="?", default="FIXED")
    ap.add_argument("--every", type=int, default=600)
    ap.add_argument("--top", type=int, default=12)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--until", type=int, default=0)
    a = ap.parse_args()
    sc = util.load(a.map)
    mc = sc.meta["map_config"]
    T = a.until or int(mc["end_time"] - mc["start_time"])
    if a.policy == "FIXED":
        m = util.marshal_map(a.map, controlled=False, vcap=8192)[1]
    else:
        m = util.marshal_map(a.map, vcap=8192, max_distance=50.0 if a.policy == "MAXWAVE" else 200.0)[1]
    o = OracleSim(m, 1, seed=a.seed)
    o.reset(a.seed, 0)
    lane_ids = sc.meta["lane_ids"]
    L = lib()
    L.orc_explain.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]
    act = np.zeros((1, m.struct.n_signals), np.int32)
    if a.policy != "FIXED":
        o.observe()
    t = 0
    step = m.struct.step_length
    while t < T:
        if a.policy == "FIXED":
            o.tick(a.every); t += a.every
        else:
            for _ in range(a.every // step):
                o.env_step(act)
                ob = o.obs()
                x = ob["mplight"] if a.policy == "MAXPRESSURE" else np.concatenate([ob["mplight"][:, :, :1], ob["wave"]], 2)
                act = util.maxpressure_actions(sc, m, x)
                t += step
        v = o.vehicles(0)
        st = o.stats()[0]
        print(f"== t={t} active {st['n_active']} arrived {st['n_arrived']} backlog {st['n_backlog']}")
        # lane heads: first vehicle of each lane (vehicles are lane-major, front first)
        lanes = v["lane"]
        first = np.r_[True, lanes[1:] != lanes[:-1]]
        heads = np.nonzero(first)[0]
        cnt = np.bincount(lanes, minlength=len(lane_ids))
        hw = v["wait"][heads]
        order = np.argsort(-hw)[:a.top]
        out = np.zeros(4, np.int32)
        for j in order:
            i = heads[j]
            if v["wait"][i] < 20:
                continue
            ln = int(lanes[i])
            L.orc_explain(o._h, 0, ln, out.ctypes.data)
            link = int(out[0])
            extra = ""
            if out[1] == 9:
                extra = f" (lane ahead {lane_ids[link]} n={cnt[link]})"
            elif link >= 0:
                extra = f" link {link}: {lane_ids[sc.arrays['link_from'][link]]} -> {lane_ids[sc.arrays['link_to'][link]]} state {chr(sc.arrays['link_state'][link])} tls {sc.arrays['link_tls'][link]} dir {chr(sc.arrays['link_dir'][link]) if sc.arrays['link_dir'][link] > 0 else '?'}"
            print(f"  lane {lane_ids[ln]:28s} n={cnt[ln]:3d} head wait {v['wait'][i]:5.0f} pos {v['pos'][i]:7.1f}/{sc.arrays['lane_len'][ln]:7.1f} "
                  f"vid {v['vid'][i]} hop {out[2]} seen {out[3] / 100:.1f}: {reason(int(out[1]))}{extra}")


if __name__ == "__main__":
    main()
