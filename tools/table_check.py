"""Consistency check of the reference's valid_acts tables (config/signal_config.py) against the green phases of the
shipped tlLogics: for every (signal, phase pair -> local action) row, does the green phase the action selects show
green to at least one lane of the pair's movements?  Rows where the pair's lanes sit AT the signal and all of them
are shown red are inconsistent: under MAXWAVE / MAXPRESSURE the approach behind such a row is served only when some
other pair wins (agents/maxwave.py:18-38), i.e. it starves once its own pressure dominates.

usage: python tools/table_check.py [map ...]      (reads the compiled scenarios; no oracle, no GPU)
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from resco_b200.abi import marshal  # noqa: E402
from resco_b200.scenario import Scenario  # noqa: E402


def check(map_name):
    """-> list of (signal, pair index, movements, action, green state, {lane: chars shown}) for the inconsistent rows,
    and the total number of rows."""
    sc = Scenario.load(os.path.join(ROOT, "resco_b200", "data", map_name + ".npz"))
    a, ids = sc.arrays, sc.meta["lane_ids"]
    mc = sc.meta["map_config"]
    m = marshal(sc, step_length=mc["step_length"], yellow_length=mc["yellow_length"])
    pairs, va = sc.meta["phase_pairs"], sc.meta["valid_acts"]
    lane_index = {n: i for i, n in enumerate(ids)}
    bad, total = [], 0
    for s in m.info["signal_ids"]:
        greens = m.info["green_states"][s]
        ls = sc.meta["signals"][s]["lane_sets"]
        order = list(ls.keys())
        t = sc.meta["tls_ids"].index(s)
        acts = va[s] if va is not None else {str(i): i for i in range(len(pairs))}
        for pk, act in acts.items():
            total += 1
            if act >= len(greens):
                bad.append((s, int(pk), [], act, None, {}))
                continue
            shown = {}
            for mv in pairs[int(pk)]:
                for lane in (ls[order[mv]] if mv < len(order) else []):
                    li = lane_index[lane]
                    ch = "".join(greens[act][a["link_tlidx"][k]] for k in range(a["lane_link_off"][li], a["lane_link_off"][li + 1])
                                 if a["link_tls"][k] == t)
                    if ch:
                        shown[lane] = ch
            if shown and not any(c in "Gg" for ch in shown.values() for c in ch):
                bad.append((s, int(pk), [order[x] for x in pairs[int(pk)] if x < len(order)], act, greens[act], shown))
    return bad, total


if __name__ == "__main__":
    maps = sys.argv[1:] or ["cologne1", "cologne3", "cologne8", "ingolstadt1", "ingolstadt7", "ingolstadt21", "grid4x4", "arterial4x4"]
    for mp in maps:
        bad, total = check(mp)
        print(f"{mp}: {len(bad)} inconsistent rows of {total}")
        for row in bad:
            print("   ", row)
