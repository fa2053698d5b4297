"""SUMO ``.net.xml`` / ``.rou.xml`` / ``.sumocfg`` reader (host side, pure Python).

The reference never parses these files itself -- it hands them to SUMO
(``multi_signal.py:117-137``).  The B200 backend replaces SUMO, so the scenario
compiler has to read the same inputs.  Only the elements the microsimulation
needs are kept: edges/lanes, junction right-of-way ``request`` rows,
connections (incl. internal ``via`` lanes, ``tl``/``linkIndex``) and ``tlLogic``
programs.
"""
from __future__ import annotations

import os
import xml.etree.ElementTree as ET
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

# vClass bits used for lane permissions (only the classes the shipped maps use)
VCLASS_BITS = {"passenger": 1, "bus": 2, "truck": 4, "delivery": 8, "taxi": 16,
               "motorcycle": 32, "trailer": 64, "coach": 128}
ALL_ROAD = 0xFF


@dataclass
class Lane:
    id: str
    edge: str
    index: int
    speed: float
    length: float
    perm: int            # vClass bitmask
    internal: bool
    shape: str = ""


@dataclass
class Edge:
    id: str
    internal: bool
    frm: Optional[str]
    to: Optional[str]
    priority: int
    lanes: List[Lane] = field(default_factory=list)


@dataclass
class Connection:
    frm: str             # from edge id
    to: str              # to edge id
    from_lane: int
    to_lane: int
    via: Optional[str]   # via lane id
    tl: Optional[str]
    link_index: int
    dir: str
    state: str


@dataclass
class Junction:
    id: str
    type: str
    inc_lanes: List[str]
    int_lanes: List[str]
    requests: List[Tuple[str, str, int]]  # (response, foes, cont) by index


@dataclass
class TLPhase:
    duration: float
    state: str


@dataclass
class TLLogic:
    id: str
    type: str
    program: str
    offset: float
    phases: List[TLPhase]


@dataclass
class Net:
    edges: Dict[str, Edge]
    lanes: Dict[str, Lane]
    junctions: Dict[str, Junction]
    connections: List[Connection]
    tls: Dict[str, TLLogic]          # insertion order == file order
    tls_order: List[str]


def _perm(allow: Optional[str], disallow: Optional[str]) -> int:
    if allow is not None:
        m = 0
        for c in allow.split():
            m |= VCLASS_BITS.get(c, 0)
        return m
    if disallow is not None:
        m = ALL_ROAD
        for c in disallow.split():
            m &= ~VCLASS_BITS.get(c, 0)
        return m
    return ALL_ROAD


def read_net(path: str) -> Net:
    root = ET.parse(path).getroot()
    edges: Dict[str, Edge] = {}
    lanes: Dict[str, Lane] = {}
    junctions: Dict[str, Junction] = {}
    conns: List[Connection] = []
    tls: Dict[str, TLLogic] = {}
    for el in root:
        tag = el.tag
        if tag == "edge":
            internal = el.get("function") == "internal"
            if el.get("function") in ("crossing", "walkingarea"):
                continue
            e = Edge(el.get("id"), internal, el.get("from"), el.get("to"),
                     int(el.get("priority", "-1")))
            for le in el.findall("lane"):
                ln = Lane(le.get("id"), e.id, int(le.get("index")), float(le.get("speed")),
                          float(le.get("length")), _perm(le.get("allow"), le.get("disallow")),
                          internal, le.get("shape", ""))
                e.lanes.append(ln)
                lanes[ln.id] = ln
            e.lanes.sort(key=lambda x: x.index)
            edges[e.id] = e
        elif tag == "junction":
            reqs = []
            for r in el.findall("request"):
                reqs.append((int(r.get("index")), r.get("response"), r.get("foes"),
                             int(r.get("cont", "0"))))
            reqs.sort()
            junctions[el.get("id")] = Junction(
                el.get("id"), el.get("type"), (el.get("incLanes") or "").split(),
                (el.get("intLanes") or "").split(), [(a, b, c) for _, a, b, c in reqs])
        elif tag == "connection":
            conns.append(Connection(el.get("from"), el.get("to"), int(el.get("fromLane")),
                                    int(el.get("toLane")), el.get("via"), el.get("tl"),
                                    int(el.get("linkIndex", "-1")), el.get("dir", "s"),
                                    el.get("state", "M")))
        elif tag == "tlLogic":
            ph = [TLPhase(float(p.get("duration")), p.get("state")) for p in el.findall("phase")]
            tid = el.get("id")
            if tid not in tls:  # SUMO uses the first program ("programs[0]", traffic_signal.py:97)
                tls[tid] = TLLogic(tid, el.get("type"), el.get("programID"),
                                   float(el.get("offset", "0")), ph)
    return Net(edges, lanes, junctions, conns, tls, list(tls.keys()))


# ----------------------------------------------------------------------------------------------
@dataclass
class VType:
    id: str
    vclass: str = "passenger"
    length: float = 5.0
    min_gap: float = 2.5
    accel: float = 2.6
    decel: float = 4.5
    tau: float = 1.0
    sigma: float = 0.5
    max_speed: float = 55.55
    speed_dev: float = 0.1
    speed_factor: float = 1.0


# vClass defaults restated from SUMO's published "Vehicle Type Parameter Defaults" table
# (https://sumo.dlr.de/docs/Vehicle_Type_Parameter_Defaults.html) -- not present in /root/reference.
_VCLASS_DEFAULTS = {
    "passenger": dict(length=5.0, min_gap=2.5, accel=2.6, decel=4.5, max_speed=55.55, speed_dev=0.1),
    "bus": dict(length=12.0, min_gap=2.5, accel=1.2, decel=4.0, max_speed=23.61, speed_dev=0.1),
    "truck": dict(length=7.1, min_gap=2.5, accel=1.3, decel=4.0, max_speed=36.11, speed_dev=0.05),
    "delivery": dict(length=6.5, min_gap=2.5, accel=2.6, decel=4.5, max_speed=55.55, speed_dev=0.05),
    "trailer": dict(length=16.5, min_gap=2.5, accel=1.1, decel=4.0, max_speed=36.11, speed_dev=0.05),
    "motorcycle": dict(length=2.2, min_gap=2.5, accel=6.0, decel=10.0, max_speed=55.55, speed_dev=0.1),
    "taxi": dict(length=5.0, min_gap=2.5, accel=2.6, decel=4.5, max_speed=55.55, speed_dev=0.1),
    "coach": dict(length=14.0, min_gap=2.5, accel=2.0, decel=4.0, max_speed=27.78, speed_dev=0.05),
}


def make_vtype(attrs: Dict[str, str]) -> VType:
    vclass = attrs.get("vClass", "passenger")
    d = dict(_VCLASS_DEFAULTS.get(vclass, _VCLASS_DEFAULTS["passenger"]))
    vt = VType(attrs.get("id", "DEFAULT_VEHTYPE"), vclass, **d)
    for k_xml, k in (("length", "length"), ("minGap", "min_gap"), ("accel", "accel"), ("decel", "decel"),
                     ("tau", "tau"), ("sigma", "sigma"), ("maxSpeed", "max_speed"),
                     ("speedDev", "speed_dev"), ("speedFactor", "speed_factor")):
        if k_xml in attrs:
            try:
                setattr(vt, k, float(attrs[k_xml]))
            except ValueError:
                pass  # e.g. speedFactor="normc(...)": keep default
    return vt


@dataclass
class Trip:
    id: str
    depart: float
    vtype: str
    frm: Optional[str] = None
    to: Optional[str] = None
    edges: Optional[List[str]] = None    # explicit route
    depart_pos: str = "base"


@dataclass
class Demand:
    vtypes: Dict[str, VType]
    trips: List[Trip]


def read_routes_xml(text_or_path, is_text: bool = False) -> Demand:
    root = ET.fromstring(text_or_path) if is_text else ET.parse(text_or_path).getroot()
    vtypes: Dict[str, VType] = {}
    trips: List[Trip] = []
    named_routes: Dict[str, List[str]] = {}
    for el in root:
        if el.tag == "vType":
            vt = make_vtype(dict(el.attrib))
            vtypes[vt.id] = vt
        elif el.tag == "route":
            named_routes[el.get("id")] = el.get("edges").split()
        elif el.tag == "trip":
            trips.append(Trip(el.get("id"), float(el.get("depart")), el.get("type", "DEFAULT_VEHTYPE"),
                              el.get("from"), el.get("to"), None, el.get("departPos", "base")))
        elif el.tag == "vehicle":
            r = el.find("route")
            edges = r.get("edges").split() if r is not None else named_routes[el.get("route")]
            trips.append(Trip(el.get("id"), float(el.get("depart")), el.get("type", "DEFAULT_VEHTYPE"),
                              edges[0], edges[-1], edges, el.get("departPos", "base")))
    if "DEFAULT_VEHTYPE" not in vtypes:
        vtypes["DEFAULT_VEHTYPE"] = make_vtype({"id": "DEFAULT_VEHTYPE"})
    return Demand(vtypes, trips)


def read_sumocfg(path: str) -> Dict[str, object]:
    """``-c <cfg>`` as used at ``multi_signal.py:40,125``: net-file, route-files, begin, end."""
    root = ET.parse(path).getroot()
    base = os.path.dirname(path)
    out: Dict[str, object] = {"begin": 0.0, "end": -1.0}
    for el in root.iter():
        if el.tag == "net-file":
            out["net"] = os.path.join(base, el.get("value"))
        elif el.tag == "route-files":
            out["routes"] = [os.path.join(base, v.strip()) for v in el.get("value").split(",")]
        elif el.tag == "begin":
            out["begin"] = float(el.get("value"))
        elif el.tag == "end":
            out["end"] = float(el.get("value"))
    return out
