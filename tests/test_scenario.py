"""Scenario compiler: table consistency on every packaged map."""
import numpy as np
import pytest

import util
from resco_b200.scenario.compiler import green_phase_indices

MAPS = ["cologne1", "cologne3", "cologne8", "ingolstadt1", "ingolstadt7", "ingolstadt21", "grid4x4", "arterial4x4"]


@pytest.mark.parametrize("name", MAPS)
def test_tables_consistent(name):
    sc = util.load(name)
    a = sc.arrays
    L, K = sc.n_lanes, sc.n_links
    assert a["lane_link_off"][0] == 0 and a["lane_link_off"][-1] == K
    assert (np.diff(a["lane_link_off"]) >= 0).all()
    for k in range(K):
        fl = a["link_from"][k]
        assert a["lane_link_off"][fl] <= k < a["lane_link_off"][fl + 1]
    # every internal lane has exactly one way on; via chains end on the link's target lane
    internal = np.nonzero(a["lane_internal"])[0]
    assert (np.diff(a["lane_link_off"])[internal] <= 1).all()
    for k in range(K):
        v = a["link_via"][k]
        hops = 0
        while v >= 0 and hops < 8:
            kk = a["lane_link_off"][v]
            assert a["link_to"][kk] == a["link_to"][k]
            v = a["link_via"][kk]
            hops += 1
    # foes reference entry links of the same junction (they share no from-lane with an internal lane)
    assert (a["lane_internal"][a["link_from"][a["foe_link"]]] == 0).all()
    # routes: consecutive edges are connected, masks are non-empty
    ro = a["route_off"]
    conn = {(int(a["lane_edge"][a["link_from"][k]]), int(a["link_to_edge"][k])) for k in range(K)
            if not a["lane_internal"][a["link_from"][k]]}
    for r in range(len(ro) - 1):
        ed = a["route_edge"][ro[r]:ro[r + 1]]
        for x, y in zip(ed[:-1], ed[1:]):
            assert (int(x), int(y)) in conn
        assert ((a["route_mask"][ro[r]:ro[r + 1]] & 0xFF) != 0).all()
        assert (((a["route_mask"][ro[r]:ro[r + 1]] >> 8) & 0xFF) != 0).all()
    # trips sorted by departure inside each origin
    for o in range(len(a["origin_lane"])):
        d = a["trip_depart"][a["origin_off"][o]:a["origin_off"][o + 1]]
        assert (np.diff(d) >= 0).all()
    assert sc.meta["n_unroutable"] == 0
    # signals: lanes are incoming lanes of TLS-controlled links or lead to one; unique across signals
    sl = a["sig_lane"]
    assert len(set(sl.tolist())) == len(sl)
    for t in sc.meta["tls_ids"]:
        assert len(green_phase_indices(sc.meta["programs"][t])) >= 1


def test_marshal_programs():
    sc, m = util.marshal_map("cologne8")
    for s in m.info["signal_ids"]:
        prog = m.info["programs_installed"][s]
        ng = len(m.info["green_states"][s])
        assert all('y' not in st for _, st in prog[:ng])
        assert all('y' in st for _, st in prog[ng:])
        assert all(d == 3 for d, _ in prog[ng:])
    assert m.struct.n_signals == 8 and m.struct.n_sig_lanes == 33


def test_generate_config_matches_reference():
    """Signal.generate_config fallback (traffic_signal.py:106-164): golden produced by the reference class
    itself on grid4x4 under a map name without a signal_configs entry (tools/make_golden_generate_config.py)."""
    import json, os
    from resco_b200.scenario.compiler import generate_signal_config
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "generate_config_grid4x4.json")))
    sc = util.load("grid4x4")
    assert set(gold) == set(sc.meta["tls_ids"])
    for t, want in gold.items():
        got = generate_signal_config(t, sc.meta["controlled_links"][t])
        assert got["lanes"] == want["lanes"], t
        assert got["lane_sets"] == want["lane_sets"], t
        assert got["downstream"] == want["downstream"], t
    # and on this regular grid the generated topology equals the hand-written signal_configs entry
    for t in sc.meta["tls_ids"]:
        got = generate_signal_config(t, sc.meta["controlled_links"][t])
        assert got["lane_sets"] == sc.meta["signals"][t]["lane_sets"], t
        assert got["downstream"] == sc.meta["signals"][t]["downstream"], t


def test_lane_change_horizon_masks():
    """route_mask per route step: 'best' lanes are a subset of the 'ok' lanes, every ok lane has a connection to the next
    route edge, and a connected lane is only left out of 'ok' if its lane-change-free continuation beyond this edge is
    shorter than the planning horizon (scenario/compiler.py:LC_HORIZON) -- ingolstadt7 has such lanes."""
    import numpy as np
    from resco_b200.scenario.compiler import LC_HORIZON
    import util
    assert LC_HORIZON == 100.0
    sc = util.load("ingolstadt7")
    a = sc.arrays
    conn = set()
    for k in range(len(a["link_from"])):
        fl = int(a["link_from"][k])
        if not a["lane_internal"][fl]:
            conn.add((fl, int(a["link_to_edge"][k])))
    dropped = 0
    for r in range(len(a["route_off"]) - 1):
        r0, r1 = int(a["route_off"][r]), int(a["route_off"][r + 1])
        for c in range(r0, r1):
            mask = int(a["route_mask"][c])
            ok, best = mask & 0xFF, (mask >> 8) & 0xFF
            assert best & ~ok == 0
            if c + 1 == r1:
                continue
            e, ne = int(a["route_edge"][c]), int(a["route_edge"][c + 1])
            l0, n = int(a["edge_lane0"][e]), min(int(a["edge_nlanes"][e]), 8)
            connected = sum(1 << j for j in range(n) if (l0 + j, ne) in conn)
            assert ok & ~connected == 0, "an ok lane without a connection to the next route edge"
            assert ok != 0 or connected == 0
            dropped += bin(connected & ~ok).count("1")
    assert dropped > 0
