"""Scenario compiler: SUMO net/route files + RESCO config dicts -> flat index tables.

Everything the device kernels read is a flat int32/float32 array (see ``Scenario``);
string ids live only in ``meta`` for the per-instance dict view that the unmodified
reference agents consume.  Cites: lane-set / downstream topology follows
``traffic_signal.py:46-87``; green-phase discovery follows ``multi_signal.py:52-59``.

The compiled form is saved as ``.npz`` (arrays) with one JSON string (``meta``).  It is
built in the dev container by ``tools/compile_scenarios.py`` from the files under
``/root/reference/resco_benchmark/environments`` and committed under ``resco_b200/data``
so that the GPU box (which has no reference tree) can load it.
"""
from __future__ import annotations

import heapq
import json
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from .netxml import Net, Demand, VCLASS_BITS, VType

MOVEMENTS = ['S-W', 'S-S', 'S-E', 'W-N', 'W-W', 'W-S', 'N-E', 'N-N', 'N-W', 'E-S', 'E-E', 'E-N']

# link state codes (connection ``state`` attribute / tlLogic state chars) are kept as ASCII
DIR_CODES = {c: i for i, c in enumerate("slrtLRi")}

LC_HORIZON = 100.0       # metres a connected lane must at least continue beyond its edge to count as "ok"
MAX_ROUTE_LANES = 8      # ok_mask is a u8 over lane indices of an edge


@dataclass
class Scenario:
    """Flat tables (numpy).  ``arrays`` keys are documented in DESIGN.md §3."""
    arrays: Dict[str, np.ndarray]
    meta: Dict[str, object]

    def save(self, path: str) -> None:
        np.savez_compressed(path, meta=np.frombuffer(json.dumps(self.meta).encode(), dtype=np.uint8),
                            **self.arrays)

    @staticmethod
    def load(path: str) -> "Scenario":
        z = np.load(path)
        meta = json.loads(bytes(z["meta"]).decode())
        arrays = {k: z[k] for k in z.files if k != "meta"}
        return Scenario(arrays, meta)

    # convenience ------------------------------------------------------------------------
    @property
    def n_lanes(self) -> int:
        return int(self.arrays["lane_len"].shape[0])

    @property
    def n_links(self) -> int:
        return int(self.arrays["link_from"].shape[0])

    @property
    def n_tls(self) -> int:
        return len(self.meta["tls_ids"])


# ------------------------------------------------------------------------------------------------
def _bit(s: str, j: int) -> bool:
    """request bitstrings are written MSB-first: char for link j is s[len-1-j]."""
    k = len(s) - 1 - j
    return 0 <= k < len(s) and s[k] == '1'


def compile_net(net: Net) -> Tuple[Dict[str, np.ndarray], Dict[str, object], Dict[str, object]]:
    """Lanes, links, junction foes, tlLogic programs."""
    lane_ids: List[str] = []
    lane_idx: Dict[str, int] = {}
    edge_ids: List[str] = []
    edge_idx: Dict[str, int] = {}
    for e in net.edges.values():
        edge_idx[e.id] = len(edge_ids)
        edge_ids.append(e.id)
        for ln in e.lanes:
            lane_idx[ln.id] = len(lane_ids)
            lane_ids.append(ln.id)
    L, E = len(lane_ids), len(edge_ids)
    lane_len = np.zeros(L, np.float32)
    lane_vmax = np.zeros(L, np.float32)
    lane_edge = np.zeros(L, np.int32)
    lane_index = np.zeros(L, np.int32)
    lane_perm = np.zeros(L, np.int32)
    lane_internal = np.zeros(L, np.int32)
    lane_left = np.full(L, -1, np.int32)
    lane_right = np.full(L, -1, np.int32)
    edge_lane0 = np.zeros(E, np.int32)
    edge_nlanes = np.zeros(E, np.int32)
    edge_internal = np.zeros(E, np.int32)
    for e in net.edges.values():
        ei = edge_idx[e.id]
        edge_lane0[ei] = lane_idx[e.lanes[0].id]
        edge_nlanes[ei] = len(e.lanes)
        edge_internal[ei] = int(e.internal)
        for k, ln in enumerate(e.lanes):
            li = lane_idx[ln.id]
            lane_len[li] = ln.length
            lane_vmax[li] = ln.speed
            lane_edge[li] = ei
            lane_index[li] = ln.index
            lane_perm[li] = ln.perm
            lane_internal[li] = int(e.internal)
            if not e.internal:
                if k > 0:
                    lane_right[li] = lane_idx[e.lanes[k - 1].id]
                if k + 1 < len(e.lanes):
                    lane_left[li] = lane_idx[e.lanes[k + 1].id]

    tls_ids = list(net.tls_order)
    tls_idx = {t: i for i, t in enumerate(tls_ids)}

    # ---- links, grouped by from-lane ------------------------------------------------------
    raw = []
    for c in net.connections:
        if c.frm not in net.edges or c.to not in net.edges:
            continue
        fe, te = net.edges[c.frm], net.edges[c.to]
        if c.from_lane >= len(fe.lanes) or c.to_lane >= len(te.lanes):
            continue
        fl = lane_idx[fe.lanes[c.from_lane].id]
        tl = lane_idx[te.lanes[c.to_lane].id]
        via = lane_idx[c.via] if c.via else -1
        raw.append((fl, tl, via, c))
    raw.sort(key=lambda r: (r[0], r[1], r[2]))
    K = len(raw)
    link_from = np.array([r[0] for r in raw], np.int32)
    link_to = np.array([r[1] for r in raw], np.int32)
    link_via = np.array([r[2] for r in raw], np.int32)
    link_tls = np.array([tls_idx.get(r[3].tl, -1) if r[3].tl else -1 for r in raw], np.int32)
    link_tlidx = np.array([r[3].link_index if r[3].tl else -1 for r in raw], np.int32)
    link_dir = np.array([DIR_CODES.get(r[3].dir, 0) for r in raw], np.int32)
    link_state = np.array([ord(r[3].state) for r in raw], np.int32)
    lane_link_off = np.zeros(L + 1, np.int32)
    for r in raw:
        lane_link_off[r[0] + 1] += 1
    lane_link_off = np.cumsum(lane_link_off).astype(np.int32)

    def links_of(lane: int) -> range:
        return range(lane_link_off[lane], lane_link_off[lane + 1])

    # final normal edge reached by each link (follows the via chain) and total via length
    link_to_edge = np.zeros(K, np.int32)
    link_via_len = np.zeros(K, np.float32)
    link_last_int = np.full(K, -1, np.int32)   # last internal lane before the normal target lane
    link_first_int = np.full(K, -1, np.int32)
    link_cont = np.zeros(K, np.int32)          # 1: has an internal junction (two-part via)
    for k in range(K):
        link_to_edge[k] = lane_edge[link_to[k]]
        v = link_via[k]
        tot, last, hops = 0.0, -1, 0
        while v >= 0 and hops < 8:
            tot += float(lane_len[v])
            last = v
            hops += 1
            nxt = list(links_of(v))
            v = link_via[nxt[0]] if nxt else -1
        link_via_len[k] = tot
        link_last_int[k] = last
        link_first_int[k] = link_via[k]
        if not lane_internal[link_from[k]] and hops >= 2:
            link_cont[k] = 1

    # ---- junction request rows -> foe lists (entry links only) --------------------------------
    # intLanes[i] is the (last-part) internal lane of the junction's link i.
    int_lane_to_jlink: Dict[int, Tuple[str, int]] = {}
    for j in net.junctions.values():
        if j.type == "internal":
            continue
        for i, lid in enumerate(j.int_lanes):
            if lid in lane_idx:
                int_lane_to_jlink[lane_idx[lid]] = (j.id, i)
    link_junc = [None] * K
    link_jidx = np.full(K, -1, np.int32)
    jlink_to_link: Dict[Tuple[str, int], int] = {}
    for k in range(K):
        if lane_internal[link_from[k]]:
            continue
        last = link_last_int[k]
        if last >= 0 and last in int_lane_to_jlink:
            jid, i = int_lane_to_jlink[last]
            link_junc[k] = jid
            link_jidx[k] = i
            jlink_to_link[(jid, i)] = k
    foe_link: List[int] = []
    foe_flags: List[int] = []
    link_foe_off = np.zeros(K + 1, np.int32)
    for k in range(K):
        jid = link_junc[k]
        if jid is not None:
            j = net.junctions[jid]
            i = int(link_jidx[k])
            if i < len(j.requests):
                resp, foes, _cont = j.requests[i]
                for jj in range(len(j.requests)):
                    f, r = _bit(foes, jj), _bit(resp, jj)
                    if (f or r) and (jid, jj) in jlink_to_link and jj != i:
                        fk = jlink_to_link[(jid, jj)]
                        mutual = _bit(j.requests[jj][0], i)
                        foe_link.append(fk)
                        foe_flags.append((1 if r else 0) | (2 if f else 0) | (4 if mutual else 0))
        link_foe_off[k + 1] = len(foe_link)
    # internal links inherit the entry link's foes (second part of a 'cont' link yields there)
    link_parent = np.full(K, -1, np.int32)     # for links leaving an internal lane: the entry link
    for k in range(K):
        if lane_internal[link_from[k]]:
            continue
        v = link_via[k]
        hops = 0
        while v >= 0 and hops < 8:
            for kk in links_of(v):
                link_parent[kk] = k
                v = link_via[kk]
                break
            else:
                v = -1
            hops += 1

    # ---- tlLogic programs ----------------------------------------------------------------------
    tls_nlinks = np.zeros(len(tls_ids), np.int32)
    for k in range(K):
        if link_tls[k] >= 0:
            tls_nlinks[link_tls[k]] = max(tls_nlinks[link_tls[k]], link_tlidx[k] + 1)
    programs = {}
    controlled = {}
    for t in tls_ids:
        lg = net.tls[t]
        programs[t] = [[p.duration, p.state] for p in lg.phases]
        n = max(int(tls_nlinks[tls_idx[t]]), max((len(p.state) for p in lg.phases), default=0))
        slots: List[List[List[str]]] = [[] for _ in range(n)]
        for k in range(K):
            if link_tls[k] == tls_idx[t]:
                via = lane_ids[link_via[k]] if link_via[k] >= 0 else ''
                slots[link_tlidx[k]].append([lane_ids[link_from[k]], lane_ids[link_to[k]], via])
        controlled[t] = slots

    arrays = dict(
        lane_len=lane_len, lane_vmax=lane_vmax, lane_edge=lane_edge, lane_index=lane_index,
        lane_perm=lane_perm, lane_internal=lane_internal, lane_left=lane_left, lane_right=lane_right,
        lane_link_off=lane_link_off, edge_lane0=edge_lane0, edge_nlanes=edge_nlanes,
        edge_internal=edge_internal,
        link_from=link_from, link_to=link_to, link_via=link_via, link_tls=link_tls,
        link_tlidx=link_tlidx, link_dir=link_dir, link_state=link_state, link_to_edge=link_to_edge,
        link_via_len=link_via_len, link_last_int=link_last_int, link_cont=link_cont,
        link_parent=link_parent, link_foe_off=link_foe_off,
        foe_link=np.array(foe_link, np.int32).reshape(-1), foe_flags=np.array(foe_flags, np.int32).reshape(-1),
    )
    meta = dict(lane_ids=lane_ids, edge_ids=edge_ids, tls_ids=tls_ids, programs=programs,
                controlled_links=controlled)
    idx = dict(lane_idx=lane_idx, edge_idx=edge_idx, tls_idx=tls_idx)
    return arrays, meta, idx


# ------------------------------------------------------------------------------------------------
class Router:
    """Shortest travel-time routing for ``<trip from= to=>`` demand.

    SUMO routes trips at load time with Dijkstra on travel time: ``length / speed`` of the normal and of the
    internal edges plus ``weights.minor-penalty`` on unsignalised minor links; tie-breaking in SUMO is not
    documented, here ties resolve to the lower edge index (deterministic).
    """

    def __init__(self, arrays: Dict[str, np.ndarray], n_edges: int):
        self.a = arrays
        self.E = n_edges
        a = arrays
        # successor normal edges per normal edge, per vclass mask.  The transition cost is what SUMO's router adds
        # between two normal edges: the travel time over the internal (via) lanes plus weights.minor-penalty
        # (1.5 s) for every unsignalised minor link on the way (MSEdge::recalcCache).
        self.succ: List[List[Tuple[int, int, float]]] = [[] for _ in range(n_edges)]  # (to_edge, perm, via cost)
        best: Dict[Tuple[int, int, int], float] = {}
        lo = a["lane_link_off"]

        def minor(k: int) -> bool:
            return int(a["link_tls"][k]) < 0 and chr(int(a["link_state"][k])) in "m=Zws"

        for k in range(len(a["link_from"])):
            fl, tl = int(a["link_from"][k]), int(a["link_to"][k])
            if a["lane_internal"][fl]:
                continue
            fe, te = int(a["lane_edge"][fl]), int(a["link_to_edge"][k])
            perm = int(a["lane_perm"][fl]) & int(a["lane_perm"][tl])
            c = 1.5 if minor(k) else 0.0
            v, hops = int(a["link_via"][k]), 0
            while v >= 0 and hops < 8:
                c += float(a["lane_len"][v]) / max(float(a["lane_vmax"][v]), 0.1)
                k2 = int(lo[v])
                if k2 >= int(lo[v + 1]):
                    break
                if minor(k2):
                    c += 1.5
                v = int(a["link_via"][k2])
                hops += 1
            key = (fe, te, perm)
            if key not in best or c < best[key]:
                best[key] = c
        for (fe, te, perm), c in best.items():
            self.succ[fe].append((te, perm, c))
        for lst in self.succ:
            lst.sort()
        self.cost = np.zeros(n_edges, np.float64)
        for e in range(n_edges):
            l0 = int(a["edge_lane0"][e])
            self.cost[e] = float(a["lane_len"][l0]) / max(float(a["lane_vmax"][l0]), 0.1)
        self._cache: Dict[Tuple[int, int, int], Optional[List[int]]] = {}

    def route(self, src: int, dst: int, vbit: int) -> Optional[List[int]]:
        key = (src, dst, vbit)
        if key in self._cache:
            return self._cache[key]
        dist = {src: self.cost[src]}
        prev: Dict[int, int] = {}
        pq = [(self.cost[src], src)]
        done = set()
        found = False
        while pq:
            d, u = heapq.heappop(pq)
            if u in done:
                continue
            done.add(u)
            if u == dst:
                found = True
                break
            for v, perm, via in self.succ[u]:
                if not (perm & vbit):
                    continue
                nd = d + via + self.cost[v]
                if v not in dist or nd < dist[v] - 1e-12:
                    dist[v] = nd
                    prev[v] = u
                    heapq.heappush(pq, (nd, v))
        out = None
        if found:
            out = [dst]
            while out[-1] != src:
                out.append(prev[out[-1]])
            out.reverse()
        self._cache[key] = out
        return out


def compile_demand(arrays: Dict[str, np.ndarray], meta: Dict[str, object], idx: Dict[str, object],
                   demand, begin: float) -> Tuple[Dict[str, np.ndarray], Dict[str, object]]:
    """vTypes, deduplicated route table (edge + ok-lane mask per step) and the trip table.

    `demand` is one Demand or a list of them: the reference loads a different route file every episode on grid4x4 and
    arterial4x4 (``route + '_' + str(self.run) + '.rou.xml'``, multi_signal.py:124).  A list compiles into ONE trip
    table -- the episodes back to back, each grouped by origin lane -- over a shared route / origin / vType table, plus
    ``bank_origin_off[R, O + 1]``: per episode, the range of the trip table that belongs to each origin lane.
    ``origin_off`` is the first episode's row; MultiSignal.reset() installs the row of its run (rs_set_demand_window)."""
    demands = list(demand) if isinstance(demand, (list, tuple)) else [demand]
    demand = Demand({k: v for d in demands for k, v in d.vtypes.items()}, [t for d in demands for t in d.trips])
    episode_of: List[int] = [e for e, d in enumerate(demands) for _ in d.trips]
    file_index: List[int] = [i for d in demands for i in range(len(d.trips))]
    edge_idx: Dict[str, int] = idx["edge_idx"]
    a = arrays
    E = len(meta["edge_ids"])
    router = Router(a, E)

    vt_ids = list(demand.vtypes.keys())
    vt_index = {v: i for i, v in enumerate(vt_ids)}
    VT = len(vt_ids)
    vt = np.zeros((VT, 8), np.float32)   # length, minGap, accel, decel, tau, sigma, maxSpeed, speedDev
    vt_bit = np.zeros(VT, np.int32)
    for i, v in enumerate(vt_ids):
        t: VType = demand.vtypes[v]
        vt[i] = [t.length, t.min_gap, t.accel, t.decel, t.tau, t.sigma, t.max_speed, t.speed_dev]
        vt_bit[i] = VCLASS_BITS.get(t.vclass, 1)

    # connection lookup: (from_edge, to_edge) -> [(from-lane index, to-lane index, permission)]
    conn_lanes: Dict[Tuple[int, int], List[Tuple[int, int, int]]] = {}
    for k in range(len(a["link_from"])):
        fl = int(a["link_from"][k])
        if a["lane_internal"][fl]:
            continue
        tl = int(a["link_to"][k])
        fe, te = int(a["lane_edge"][fl]), int(a["link_to_edge"][k])
        perm = int(a["lane_perm"][fl]) & int(a["lane_perm"][tl])
        conn_lanes.setdefault((fe, te), []).append((int(a["lane_index"][fl]), int(a["lane_index"][tl]), perm))

    route_key_to_id: Dict[Tuple[Tuple[int, ...], int], int] = {}
    route_edges: List[int] = []
    route_mask: List[int] = []
    route_off: List[int] = [0]

    def intern_route(edges: Sequence[int], vbit: int) -> int:
        """Per step: bits 0-7 = lanes with a connection to the next route edge ("ok"),
        bits 8-15 = the ok lanes from which the route can be followed furthest without a lane
        change ("best"; backward DP over the connection graph, cf. SUMO's bestLanes)."""
        key = (tuple(edges), vbit)
        if key in route_key_to_id:
            return route_key_to_id[key]
        rid = len(route_off) - 1
        n_steps = len(edges)
        cont: List[List[float]] = [[] for _ in range(n_steps)]
        oks = [0] * n_steps
        bests = [0] * n_steps
        for s in range(n_steps - 1, -1, -1):
            e = edges[s]
            n = min(int(a["edge_nlanes"][e]), MAX_ROUTE_LANES)
            l0 = int(a["edge_lane0"][e])
            ln = float(a["lane_len"][l0])
            c = [0.0] * n
            ok = 0
            for li in range(n):
                if not (int(a["lane_perm"][l0 + li]) & vbit):
                    c[li] = -1.0
                    continue
                if s + 1 < n_steps:
                    best_next = -1.0
                    for fi, ti, perm in conn_lanes.get((e, edges[s + 1]), []):
                        if fi == li and (perm & vbit) and ti < len(cont[s + 1]) and cont[s + 1][ti] >= 0:
                            best_next = max(best_next, cont[s + 1][ti])
                    if best_next >= 0:
                        ok |= 1 << li
                        c[li] = ln + best_next
                    else:
                        c[li] = ln
                else:
                    ok |= 1 << li
                    c[li] = ln
            cont[s] = c
            cand = [li for li in range(n) if (ok >> li) & 1]
            if cand:
                mx = max(c[li] for li in cand)
                for li in cand:
                    if c[li] >= mx - 0.5:
                        bests[s] |= 1 << li
                # planning horizon (cf. LC2013's look-ahead distance): a lane that does connect to the next route
                # edge but runs out within LC_HORIZON metres beyond this edge is not a lane to stay on -- the change
                # has to happen on THIS edge, while there is still room for it
                ok = 0
                for li in cand:
                    if c[li] >= mx - 0.5 or c[li] - ln >= LC_HORIZON:
                        ok |= 1 << li
            oks[s] = ok
        for s, e in enumerate(edges):
            route_edges.append(e)
            route_mask.append(oks[s] | (bests[s] << 8))
        route_off.append(len(route_edges))
        route_key_to_id[key] = rid
        return rid

    n_unroutable = 0
    rows = []
    n_before_begin = 0
    for fi, tr in enumerate(demand.trips):
        if tr.depart < begin:     # SUMO discards vehicles whose departure lies before --begin (cologne3: 1638)
            n_before_begin += 1
            continue
        vti = vt_index.get(tr.vtype, vt_index.get("DEFAULT_VEHTYPE", 0))
        vbit = int(vt_bit[vti])
        if tr.edges is not None:
            if any(e not in edge_idx for e in tr.edges):
                n_unroutable += 1
                continue
            edges = [edge_idx[e] for e in tr.edges]
        else:
            if tr.frm not in edge_idx or tr.to not in edge_idx:
                n_unroutable += 1
                continue
            edges = router.route(edge_idx[tr.frm], edge_idx[tr.to], vbit)
            if edges is None:
                n_unroutable += 1
                continue
        rid = intern_route(edges, vbit)
        m0 = route_mask[route_off[rid]]
        if m0 & 0xFF == 0:
            n_unroutable += 1
            continue
        m0 = (m0 >> 8) & 0xFF or (m0 & 0xFF)   # insert on the lowest-index best lane
        if m0 == 0:
            n_unroutable += 1
            continue
        lane_i = (m0 & -m0).bit_length() - 1      # lowest-index usable lane ("best"-lane insertion)
        origin_lane = int(a["edge_lane0"][edges[0]]) + lane_i
        rows.append((origin_lane, tr.depart - begin, file_index[fi], rid, vti, episode_of[fi], fi,
                     1 if tr.depart_pos == "random_free" else 0))
    rows.sort(key=lambda r: (r[5], r[0], r[1], r[2]))
    origins = sorted({r[0] for r in rows})
    o_index = {o: i for i, o in enumerate(origins)}
    origin_lane = np.array(origins, np.int32)
    R, O = len(demands), len(origins)
    counts = np.zeros((R, O), np.int64)
    for r in rows:
        counts[r[5], o_index[r[0]]] += 1
    bank = np.zeros((R, O + 1), np.int32)
    base = 0
    for e in range(R):
        bank[e, 0] = base
        bank[e, 1:] = base + np.cumsum(counts[e])
        base = int(bank[e, O])
    origin_off = bank[0].copy()
    out = dict(
        vtype=vt, vtype_bit=vt_bit,
        route_off=np.array(route_off, np.int32), route_edge=np.array(route_edges, np.int32).reshape(-1),
        route_mask=np.array(route_mask, np.int32).reshape(-1),
        origin_lane=origin_lane, origin_off=origin_off,
        trip_depart=np.array([r[1] for r in rows], np.float32).reshape(-1),
        trip_file=np.array([r[2] for r in rows], np.int32).reshape(-1),
        trip_route=np.array([r[3] for r in rows], np.int32).reshape(-1),
        trip_vtype=np.array([r[4] for r in rows], np.int32).reshape(-1),
        # <vehicle departPos="random_free"> (arterial4x4 route files): 1, else 0 (departPos="base")
        trip_depart_pos=np.array([r[7] for r in rows], np.int32).reshape(-1),
    )
    if R > 1:
        out["bank_origin_off"] = bank
    dmeta = dict(vtype_ids=vt_ids, n_trips_file=len(demands[0].trips), n_unroutable=n_unroutable,
                 n_before_begin=n_before_begin, n_demand_episodes=R,
                 trip_ids=[demand.trips[r[6]].id for r in rows])
    return out, dmeta


# ------------------------------------------------------------------------------------------------
def green_phase_indices(program: List[List[object]]) -> List[int]:
    """``multi_signal.py:52-59``: keep phases with no 'y' and at least one 'g'/'G'."""
    return [i for i, (_d, st) in enumerate(program) if 'y' not in st and 'g' in st.lower()]


def generate_signal_config(sig_id: str, controlled_links: List[List[List[str]]]) -> Dict[str, object]:
    """Fallback topology for a signal without a ``signal_configs`` entry, restating
    ``Signal.generate_config`` (traffic_signal.py:106-164): every third controlled link opens one of the
    12 movements in the fixed order S-W, S-S, S-E, W-N, ... ; the inbound lane of that link is the
    movement's lane set; the downstream signal of a direction is the first ``letters+digits`` token of
    the through-lane's id unless it names a fringe (top/right/left/bottom).  ``lanes`` keeps the order in
    which inbound lanes first appear in the link list.  (Like the reference, no outbound lane sets are
    derived: states that subtract downstream queues need a real ``signal_configs`` entry.)"""
    import re
    order = ['S-W', 'S-S', 'S-E', 'W-N', 'W-W', 'W-S', 'N-E', 'N-N', 'N-W', 'E-S', 'E-E', 'E-N']
    lane_sets: Dict[str, List[str]] = {mv: [] for mv in order}
    lanes: List[str] = []
    for i, slot in enumerate(controlled_links):
        if not slot:
            continue
        inbound = slot[0][0]
        if inbound not in lanes:
            lanes.append(inbound)
        if i % 3 == 0 and i // 3 < len(order):
            lane_sets[order[i // 3]].append(inbound)
    downstream: Dict[str, Optional[str]] = {'N': None, 'E': None, 'S': None, 'W': None}
    for through, direction in (('S-S', 'N'), ('N-N', 'S'), ('W-W', 'E'), ('E-E', 'W')):
        if not lane_sets[through]:
            continue
        tokens = re.findall('[a-zA-Z]+[0-9]+', lane_sets[through][0])
        if tokens and not any(f in tokens[0] for f in ('top', 'right', 'left', 'bottom')):
            downstream[direction] = tokens[0]
    return dict(lane_sets=lane_sets, downstream=downstream, lanes=lanes)


def compile_signals(meta: Dict[str, object], idx: Dict[str, object], signal_config: Dict[str, object],
                    lights: Sequence[str]) -> Tuple[Dict[str, np.ndarray], Dict[str, object]]:
    """Per-signal lane topology, in the iteration order of ``traffic_signal.py:46-87``."""
    lane_idx: Dict[str, int] = idx["lane_idx"]
    tls_ids: List[str] = meta["tls_ids"]
    sig_ids = list(lights) if len(lights) > 0 else list(tls_ids)
    sig_index = {s: i for i, s in enumerate(sig_ids)}
    reversed_directions = {'N': 'S', 'E': 'W', 'S': 'N', 'W': 'E'}

    sig_lane_off = [0]
    sig_lanes: List[int] = []
    mv_off = [0]              # [S*12+1] movement -> inbound lanes (positions into the signal's lane list)
    mv_lane: List[int] = []
    mvo_off = [0]             # movement -> outbound (downstream) lanes: (down signal, slot in its lane list)
    mvo_sig: List[int] = []
    mvo_slot: List[int] = []
    out_off = [0]             # signal -> outbound_lanes (for rewards.pressure)
    out_sig: List[int] = []
    out_slot: List[int] = []
    sig_meta = {}
    # pass 1: lane lists
    lanes_of: Dict[str, List[str]] = {}
    cfg_of: Dict[str, Dict[str, object]] = {}
    for s in sig_ids:
        if s in signal_config:
            cfg = signal_config[s]
            lanes: List[str] = []
            for direction in cfg['lane_sets']:
                for lane in cfg['lane_sets'][direction]:
                    if lane not in lanes:
                        lanes.append(lane)
        else:       # Signal.generate_config fallback (traffic_signal.py:88-89,106-164)
            cfg = generate_signal_config(s, meta["controlled_links"][s])
            lanes = list(cfg['lanes'])
        cfg_of[s] = cfg
        lanes_of[s] = lanes
    for s in sig_ids:
        cfg = cfg_of[s]
        lane_sets = cfg['lane_sets']
        downstream = cfg['downstream']
        lanes = lanes_of[s]
        inbounds_fr_direction: Dict[str, List[str]] = {}
        for direction in lane_sets:
            for lane in lane_sets[direction]:
                fr = reversed_directions[direction.split('-')[0]]
                inbounds_fr_direction.setdefault(fr, [])
                if lane not in inbounds_fr_direction[fr]:
                    inbounds_fr_direction[fr].append(lane)
        outbound_lanes: List[str] = []
        out_lane_to_signalid: Dict[str, str] = {}
        lane_sets_outbound: Dict[str, List[str]] = {k: [] for k in lane_sets}
        for direction in downstream:
            dwn = downstream[direction]
            if dwn is None:
                continue
            if dwn not in cfg_of or 'lanes' in cfg:     # generated configs derive no outbound lane sets
                continue
            dwn_lane_sets = cfg_of[dwn]['lane_sets']
            for key in dwn_lane_sets:
                if key.split('-')[0] == direction:
                    dwn_lane_set = dwn_lane_sets[key]
                    for lane in dwn_lane_set:
                        if lane not in outbound_lanes:
                            outbound_lanes.append(lane)
                        out_lane_to_signalid[lane] = dwn
                        for selfkey in lane_sets:
                            if selfkey.split('-')[1] == key.split('-')[0]:
                                lane_sets_outbound[selfkey] += dwn_lane_set
        for key in lane_sets_outbound:   # reference dedups through set(); order is irrelevant (sums)
            lane_sets_outbound[key] = sorted(set(lane_sets_outbound[key]))

        for lane in lanes:
            sig_lanes.append(lane_idx[lane])
        sig_lane_off.append(len(sig_lanes))
        for mvname in lane_sets:           # dict order == the 12 movement keys in config order
            for lane in lane_sets[mvname]:
                mv_lane.append(lanes.index(lane))
            mv_off.append(len(mv_lane))
            for lane in lane_sets_outbound[mvname]:
                dwn = out_lane_to_signalid[lane]
                if dwn in sig_index:   # ``if dwn_signal in signal.signals`` (states.py:75)
                    mvo_sig.append(sig_index[dwn])
                    mvo_slot.append(lanes_of[dwn].index(lane))
            mvo_off.append(len(mvo_sig))
        for lane in outbound_lanes:
            dwn = out_lane_to_signalid[lane]
            if dwn in sig_index:
                out_sig.append(sig_index[dwn])
                out_slot.append(lanes_of[dwn].index(lane))
        out_off.append(len(out_sig))
        sig_meta[s] = dict(lanes=lanes, lane_sets=lane_sets, downstream=downstream,
                           lane_sets_outbound=lane_sets_outbound, outbound_lanes=outbound_lanes,
                           out_lane_to_signalid=out_lane_to_signalid,
                           inbounds_fr_direction=inbounds_fr_direction,
                           movement_keys=list(lane_sets.keys()))
    tls_index = {t: i for i, t in enumerate(tls_ids)}
    arrays = dict(
        sig_tls=np.array([tls_index[s] for s in sig_ids], np.int32),
        sig_lane_off=np.array(sig_lane_off, np.int32), sig_lane=np.array(sig_lanes, np.int32).reshape(-1),
        mv_off=np.array(mv_off, np.int32), mv_lane=np.array(mv_lane, np.int32).reshape(-1),
        mvo_off=np.array(mvo_off, np.int32), mvo_sig=np.array(mvo_sig, np.int32).reshape(-1),
        mvo_slot=np.array(mvo_slot, np.int32).reshape(-1),
        out_off=np.array(out_off, np.int32), out_sig=np.array(out_sig, np.int32).reshape(-1),
        out_slot=np.array(out_slot, np.int32).reshape(-1),
    )
    return arrays, dict(signal_ids=sig_ids, signals=sig_meta)


def compile_tls_dist(arrays: Dict[str, np.ndarray]) -> np.ndarray:
    """Distance from the END of each lane to the next TLS stop line along the (unique) way ahead.

    ``vehicle.getNextTLS(v)[0][2]`` (traffic_signal.py:241-244) is "distance to the next
    traffic light on the route".  0 for lanes whose links are TLS controlled; for lanes with a
    single successor the successor's value plus its length; -1 (no TLS ahead / ambiguous) else.
    """
    a = arrays
    L = len(a["lane_len"])
    out = np.full(L, -1.0, np.float32)
    off = a["lane_link_off"]
    for l in range(L):
        ks = range(off[l], off[l + 1])
        if any(a["link_tls"][k] >= 0 for k in ks):
            out[l] = 0.0
    changed = True
    it = 0
    while changed and it < 64:
        changed = False
        it += 1
        for l in range(L):
            if out[l] >= 0:
                continue
            nxt = {int(a["link_via"][k]) if a["link_via"][k] >= 0 else int(a["link_to"][k])
                   for k in range(off[l], off[l + 1])}
            if len(nxt) >= 1 and all(out[n] >= 0 for n in nxt):
                vals = {round(float(out[n] + a["lane_len"][n]), 3) for n in nxt}
                if len(vals) == 1:
                    out[l] = np.float32(next(iter(vals)))
                    changed = True
    return out


def compile_watch(arrays: Dict[str, np.ndarray], origin_lanes: Sequence[int], reach: float = 60.0
                  ) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Per origin lane: the lanes within `reach` metres upstream (reverse BFS over links) and the
    distance from the END of each to the START of the origin lane.  Used by the insertion safety
    check against vehicles that are about to drive onto the origin lane."""
    a = arrays
    L = len(a["lane_len"])
    preds: List[List[int]] = [[] for _ in range(L)]
    for k in range(len(a["link_from"])):
        nxt = int(a["link_via"][k]) if a["link_via"][k] >= 0 else int(a["link_to"][k])
        fl = int(a["link_from"][k])
        if fl not in preds[nxt]:
            preds[nxt].append(fl)
    off = [0]
    lanes: List[int] = []
    dists: List[float] = []
    for o in origin_lanes:
        seen = {}
        frontier = [(int(o), 0.0)]
        while frontier:
            nxt_frontier = []
            for lane, d in frontier:
                for p in sorted(preds[lane]):
                    if p in seen and seen[p] <= d:
                        continue
                    seen[p] = d
                    dp = d + float(a["lane_len"][p])
                    if dp < reach:
                        nxt_frontier.append((p, dp))
            frontier = nxt_frontier
        for p in sorted(seen):
            lanes.append(p)
            dists.append(seen[p])
        off.append(len(lanes))
    return (np.array(off, np.int32), np.array(lanes, np.int32).reshape(-1), np.array(dists, np.float32).reshape(-1))


def compile_scenario(net: Net, demand: Optional[Demand], map_name: str, map_config: Dict[str, object],
                     signal_config: Dict[str, object], begin: float) -> Scenario:
    arrays, meta, idx = compile_net(net)
    if demand is not None:
        darr, dmeta = compile_demand(arrays, meta, idx, demand, begin)
        arrays.update(darr)
        meta.update(dmeta)
    sarr, smeta = compile_signals(meta, idx, signal_config, map_config.get('lights', []))
    arrays.update(sarr)
    meta.update(smeta)
    arrays["lane_tls_dist"] = compile_tls_dist(arrays)
    if "origin_lane" in arrays:
        w_off, w_lane, w_dist = compile_watch(arrays, arrays["origin_lane"].tolist())
        arrays.update(origin_watch_off=w_off, origin_watch_lane=w_lane, origin_watch_dist=w_dist)
    # lanes a vehicle can change into: the same upstream table, so that a lane changer sees who is about to come
    # out of the junction behind it (empty for single-lane edges and internal lanes)
    L = len(arrays["lane_len"])
    multi = [l for l in range(L) if not arrays["lane_internal"][l]
             and (arrays["lane_left"][l] >= 0 or arrays["lane_right"][l] >= 0)]
    m_off, m_lane, m_dist = compile_watch(arrays, multi)
    lw_off = np.zeros(L + 1, np.int32)
    for i, l in enumerate(multi):
        lw_off[l + 1] = m_off[i + 1] - m_off[i]
    arrays.update(lane_watch_off=np.cumsum(lw_off).astype(np.int32), lane_watch_lane=m_lane, lane_watch_dist=m_dist)
    meta.update(dict(map_name=map_name, begin=begin,
                     map_config={k: v for k, v in map_config.items() if k not in ('net', 'route')},
                     phase_pairs=signal_config.get('phase_pairs'),
                     valid_acts=({k: {str(a): b for a, b in v.items()} for k, v in signal_config['valid_acts'].items()}
                                 if signal_config.get('valid_acts') else None)))
    return Scenario(arrays, meta)
