"""Extract the reference's published per-episode average delays (utils/avg_timeLoss.py, the numbers behind the
paper's delay plots) into a small fixture: tests/golden/anchors_ref.json.  Run in the dev container, where
/root/reference is mounted; the fixture travels with the repo.

For every (agent, map) row: number of episodes, mean, median, 10th / 90th percentile, min, max and the first episode.
Learning agents improve over episodes; the static controllers (FIXED / MAXWAVE / MAXPRESSURE / STOCHASTIC) are
stationary, so their spread is the run-to-run variation of `sumo --random` -- including the episodes in which SUMO
itself gridlocks (e.g. MAXPRESSURE cologne3: median 26 s, mean 162 s, max 692 s).
"""
import ast
import json
import os
import re
import sys

import numpy as np

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
src = open(os.path.join(REF, "resco_benchmark", "utils", "avg_timeLoss.py")).read()
m = re.search(r"=\s*\{", src)
data = ast.literal_eval(src[m.end() - 1:])
out = {}
for key, vals in data.items():
    if key.endswith("_yerr") or len(vals) == 0:
        continue
    agent, map_name = key.split()[:2]
    v = np.asarray(vals, np.float64)
    out[f"{agent} {map_name}"] = dict(n=int(len(v)), mean=float(v.mean()), median=float(np.median(v)),
                                      p10=float(np.percentile(v, 10)), p90=float(np.percentile(v, 90)),
                                      min=float(v.min()), max=float(v.max()), first=float(v[0]))
dst = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "anchors_ref.json")
json.dump(out, open(dst, "w"), indent=1, sort_keys=True)
print("wrote", dst, len(out), "rows")
