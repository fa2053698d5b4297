"""Synthetic demand (BASELINE.json configs[4] / SURVEY §8(d) C5): independent arrivals per ENTRY LANE.

Every tick each entry lane draws a Bernoulli(rate/3600) insertion request from Philox keyed
(seed, instance, lane, tick) -- the discrete-time form of a Poisson stream, capped at one request per
lane and second, which is also the physical insertion limit -- and the j-th vehicle of a lane picks its
route uniformly (Philox keyed (seed, instance, lane, j)) among the shortest routes from that lane's
edge to every exit edge that the lane's connections allow (dedicated turn lanes on grid4x4:
lane 0 -> right, 1 -> straight, 2 -> left).  Requests that find no room wait in a per-lane backlog
(departDelay grows; expected above signalised capacity).
"""
from __future__ import annotations

from typing import Dict

import numpy as np

from .compiler import Router, Scenario, compile_demand, compile_watch
from .netxml import Demand, Trip, make_vtype


def synth_demand(sc: Scenario, rate_veh_h_lane: float) -> Dict[str, np.ndarray]:
    a = sc.arrays
    E = len(a["edge_lane0"])
    edge_ids = sc.meta["edge_ids"]
    normal = [e for e in range(E) if not a["lane_internal"][a["edge_lane0"][e]]]
    has_in, has_out = set(), set()
    for k in range(len(a["link_from"])):
        fl = int(a["link_from"][k])
        if a["lane_internal"][fl]:
            continue
        has_out.add(int(a["lane_edge"][fl]))
        has_in.add(int(a["link_to_edge"][k]))
    entries = [e for e in normal if e not in has_in and e in has_out]
    exits = [e for e in normal if e not in has_out and e in has_in]
    router = Router(a, E)
    trips = []
    for en in entries:
        for ex in exits:
            r = router.route(en, ex, 1)
            if r is None or len(r) <= 2:       # unreachable, or the immediate U-turn at the first junction
                continue
            trips.append(Trip(f"{edge_ids[en]}>{edge_ids[ex]}", 0.0, "synthetic", edge_ids[en], edge_ids[ex],
                              [edge_ids[x] for x in r]))
    vt = make_vtype({"id": "synthetic", "vClass": "passenger"})
    idx = dict(edge_idx={e: i for i, e in enumerate(edge_ids)})
    darr, _ = compile_demand(a, sc.meta, idx, Demand({"synthetic": vt}, trips), 0.0)
    ro = darr["route_off"]
    n_routes = len(ro) - 1
    lanes = []
    for en in entries:
        l0, n = int(a["edge_lane0"][en]), int(a["edge_nlanes"][en])
        lanes += [l0 + i for i in range(n)]
    lanes.sort()
    per_lane = {l: [] for l in lanes}
    for r in range(n_routes):
        e0 = int(darr["route_edge"][ro[r]])
        m0 = int(darr["route_mask"][ro[r]])
        best = (m0 >> 8) & 0xFF or (m0 & 0xFF)
        for i in range(int(a["edge_nlanes"][e0])):
            if (best >> i) & 1:
                per_lane[int(a["edge_lane0"][e0]) + i].append(r)
    lanes = [l for l in lanes if per_lane[l]]
    off = [0]
    flat = []
    for l in lanes:
        flat += per_lane[l]
        off.append(len(flat))
    p = min(max(rate_veh_h_lane / 3600.0, 0.0), 1.0)
    w_off, w_lane, w_dist = compile_watch(a, lanes)
    return dict(origin_lane=np.array(lanes, np.int32), origin_rate=np.full(len(lanes), int(round(p * (1 << 24))), np.int32),
                origin_route_off=np.array(off, np.int32), origin_route=np.array(flat, np.int32),
                origin_watch_off=w_off, origin_watch_lane=w_lane, origin_watch_dist=w_dist,
                route_off=darr["route_off"], route_edge=darr["route_edge"], route_mask=darr["route_mask"],
                vtype_table=darr["vtype"], vtype_bit=darr["vtype_bit"], vtype=0,
                rate_veh_h_lane=float(rate_veh_h_lane), n_entry_lanes=len(lanes))
