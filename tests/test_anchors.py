"""Statistical anchors of the simulator core against the reference's published episode averages.

SUMO is not available here or on the GPU box, so per-vehicle parity with libsumo cannot be tested.  What the
reference does publish is utils/avg_timeLoss.py: per-episode averages of timeLoss + departDelay (utils/readXML.py:27-77)
of `sumo --random` runs for every (agent, map).  tests/golden/anchors_ref.json holds their distribution (made by
tools/make_anchor_fixture.py); this test runs the same controllers on the CPU oracle -- which the CUDA path reproduces
bit for bit (tests/test_gpu_parity.py) -- measures the same quantity through the same tripinfo pipeline
(tools/anchors.py -> metrics.write_tripinfo -> metrics.avg_delay_from_tripinfo) and asserts the ratio to the
reference's median stays inside the band written next to each row.  A model change that moves an anchor fails here.

Bands: the default is [0.70, 1.25].  Rows with their own band say why.
"""
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import anchors  # noqa: E402
import table_check  # noqa: E402
import util  # noqa: E402

REF = json.load(open(os.path.join(ROOT, "tests", "golden", "anchors_ref.json")))
SEEDS = 3
DEFAULT = (0.70, 1.25)
# (map, policy) -> (lo, hi, statistic over seeds, why)
BANDS = {
    # SUMO itself gridlocks on cologne3 under MAXPRESSURE in more than one episode in ten (reference: median 28 s,
    # p90 488 s, max 692 s); so does this model on some seeds.  The anchor is the best seed against the median.
    ("cologne3", "MAXPRESSURE"): (0.70, 1.25, "min", "reference distribution is bimodal (gridlock episodes)"),
    # OPEN DEVIATION.  The two signals 243641585 / gneJ257 are 12 m apart and run uncoordinated programs (86 s / 90 s
    # cycles); 364 trips/h turn left through both.  The model discharges ~130 veh/h there and the queue spills back
    # over the -201201945#* chain.  Whether SUMO does better at this spot cannot be checked here.  The band is a
    # regression guard around today's value, NOT an agreement claim.
    ("ingolstadt21", "FIXED"): (1.6, 2.2, "median", "open deviation: coupled signals 243641585/gneJ257"),
    # The reference's 30 MAXWAVE episodes on arterial4x4 are skewed (median 711 s, mean 821 s, p90 1219 s, max 1258 s): the
    # map is oversaturated (under half of the 2484 trips ever depart) and some episodes lock up.  Five seeds of this model
    # with departPos="random_free" spread over 825 .. 1011 s; the median of three has to stay below the reference's p90.
    ("arterial4x4", "MAXWAVE"): (0.70, 1.45, "median", "oversaturated map, skewed reference distribution"),
}
MAPS_FIXED = ["cologne1", "cologne3", "cologne8", "ingolstadt1", "ingolstadt7", "ingolstadt21"]
MAPS_CTRL = ["cologne1", "cologne3", "cologne8", "ingolstadt1", "ingolstadt7", "grid4x4", "arterial4x4"]
ROWS = [(m, "FIXED") for m in MAPS_FIXED] + [(m, p) for m in MAPS_CTRL for p in ("MAXPRESSURE", "MAXWAVE")]


@pytest.mark.parametrize("map_name,policy", ROWS, ids=[f"{m}-{p}" for m, p in ROWS])
def test_anchor(map_name, policy):
    ref = REF[f"{policy} {map_name}"]
    d, st = anchors.run(map_name, policy, SEEDS)
    assert int(st["anomalies"].sum()) == 0
    lo, hi, stat, _why = BANDS.get((map_name, policy), DEFAULT + ("median", ""))
    ours = float(np.min(d) if stat == "min" else np.median(d))
    ratio = ours / ref["median"]
    assert lo <= ratio <= hi, f"{map_name}/{policy}: {ours:.1f} s vs reference median {ref['median']:.1f} s (ratio {ratio:.2f}, band [{lo}, {hi}])"


def test_valid_acts_tables_are_consistent_except_ingolstadt21():
    """tools/table_check.py: on seven maps every (phase pair -> action) row of config/signal_config.py selects a green
    phase that serves the pair; on ingolstadt21 three rows select a phase that shows the pair's own approach red (the
    shipped table does not match the shipped tlLogic order at 243641585 and cluster_1427494838_273472399)."""
    for mp in MAPS_CTRL:
        assert table_check.check(mp)[0] == []
    bad, total = table_check.check("ingolstadt21")
    assert sorted((s, p) for s, p, *_ in bad) == [("243641585", 2), ("243641585", 7), ("cluster_1427494838_273472399", 12)]


def _with_consistent_table(fn):
    orig = util.load

    def patched(name):
        sc = orig(name)
        if name == "ingolstadt21":     # the assignment under which every row serves its own pair (see the test above)
            sc.meta["valid_acts"]["243641585"] = {"2": 2, "4": 0, "7": 1}
            sc.meta["valid_acts"]["cluster_1427494838_273472399"] = {"4": 0, "7": 1, "12": 3}
        return sc
    util.load = patched
    try:
        return fn()
    finally:
        util.load = orig


def test_ingolstadt21_maxwave():
    """C3 / C4 run on ingolstadt21.  With the table AS SHIPPED the approach 23166741#5 of signal 243641585 (705 trips/h)
    is never served under MAXWAVE -- its own pair selects the phase that shows it red and no other pair can beat it
    (S pair up to 21 vehicles in range, W pair at most 4 on its 12 m lanes) -- and the network gridlocks; no simulator
    can produce the published 70 s from these files.  With the three inconsistent rows re-assigned so that every pair
    selects a phase serving it, the model lands on the reference's median: that is the anchor of the simulator core on
    this map."""
    ref = REF["MAXWAVE ingolstadt21"]
    shipped, _ = anchors.run("ingolstadt21", "MAXWAVE", 2)
    assert np.median(shipped) > 3.0 * ref["median"]           # starvation signature of the shipped table
    d, st = _with_consistent_table(lambda: anchors.run("ingolstadt21", "MAXWAVE", SEEDS))
    ratio = float(np.median(d)) / ref["median"]
    assert 0.80 <= ratio <= 1.25, f"consistent table: {np.median(d):.1f} s vs {ref['median']:.1f} s"


def test_ingolstadt21_maxpressure_open_deviation():
    """OPEN DEVIATION (regression guard, not an agreement claim): MAXPRESSURE on ingolstadt21 with the consistent table
    sits at about twice the reference's median (reference: median 116 s, p90 192 s, max 373 s); max-pressure without a
    fairness term starves low-pressure approaches for tens of minutes, and which ones tip over depends on details of
    the queue dynamics that cannot be compared with SUMO here."""
    ref = REF["MAXPRESSURE ingolstadt21"]
    d, st = _with_consistent_table(lambda: anchors.run("ingolstadt21", "MAXPRESSURE", SEEDS))
    ratio = float(np.median(d)) / ref["median"]
    assert 1.2 <= ratio <= 2.6, f"{np.median(d):.1f} s vs {ref['median']:.1f} s"
