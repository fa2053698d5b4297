"""ctypes mirror of ``include/resco_b200.h`` and the Scenario -> RsScenario marshalling.

The struct layouts here must match the header field for field (checked by
``tests/test_abi.py`` against ``rs_abi_version`` and a size probe).
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from .scenario.compiler import Scenario, green_phase_indices

RS_ABI_VERSION = 5

_I32P = C.POINTER(C.c_int32)
_F32P = C.POINTER(C.c_float)
_U8P = C.POINTER(C.c_uint8)

_SIZES = ["n_lanes", "n_edges", "n_links", "n_foes", "n_tls", "n_phases", "n_state_chars", "n_signals",
          "n_sig_lanes", "n_mv_lanes", "n_mvo", "n_out", "n_yellow",
          "n_vtypes", "n_routes", "n_route_steps", "n_origins", "n_trips", "n_origin_routes", "n_watch", "n_lane_watch"]

_PTRS: List[Tuple[str, object]] = [
    ("lane_len", _F32P), ("lane_vmax", _F32P), ("lane_edge", _I32P), ("lane_index", _I32P),
    ("lane_perm", _I32P), ("lane_internal", _I32P), ("lane_left", _I32P), ("lane_right", _I32P),
    ("lane_link_off", _I32P), ("lane_tls_dist", _F32P), ("lane_sig", _I32P), ("lane_sig_slot", _I32P),
    ("edge_lane0", _I32P), ("edge_nlanes", _I32P),
    ("link_from", _I32P), ("link_to", _I32P), ("link_via", _I32P), ("link_tls", _I32P),
    ("link_tlidx", _I32P), ("link_state", _I32P), ("link_to_edge", _I32P), ("link_via_len", _F32P),
    ("link_last_int", _I32P), ("link_cont", _I32P), ("link_parent", _I32P), ("link_foe_off", _I32P),
    ("foe_link", _I32P), ("foe_flags", _I32P),
    ("tls_phase_off", _I32P), ("tls_nlinks", _I32P), ("tls_init_phase", _I32P), ("tls_init_left", _I32P),
    ("phase_dur", _I32P), ("phase_state_off", _I32P), ("state_chars", _U8P),
    ("sig_tls", _I32P), ("sig_n_green", _I32P), ("sig_yellow_off", _I32P), ("yellow_idx", _I32P),
    ("sig_lane_off", _I32P), ("sig_lane", _I32P), ("mv_off", _I32P), ("mv_lane", _I32P),
    ("mvo_off", _I32P), ("mvo_sig", _I32P), ("mvo_slot", _I32P), ("out_off", _I32P), ("out_sig", _I32P),
    ("out_slot", _I32P),
    ("vtype", _F32P), ("vtype_bit", _I32P), ("route_off", _I32P), ("route_edge", _I32P),
    ("route_mask", _I32P), ("origin_lane", _I32P), ("origin_off", _I32P), ("trip_depart", _F32P),
    ("trip_route", _I32P), ("trip_vtype", _I32P), ("trip_file", _I32P), ("trip_depart_pos", _I32P),
    ("origin_rate", _I32P), ("origin_route_off", _I32P), ("origin_route", _I32P),
    ("origin_watch_off", _I32P), ("origin_watch_lane", _I32P), ("origin_watch_dist", _F32P),
    ("origin_watch_owner", _I32P),
    ("lane_watch_off", _I32P), ("lane_watch_lane", _I32P), ("lane_watch_dist", _F32P),
]

_PARAMS = [("synthetic", C.c_int32), ("synthetic_vtype", C.c_int32), ("step_length", C.c_int32),
           ("yellow_length", C.c_int32), ("end_tick", C.c_int32), ("max_distance", C.c_float),
           ("sigma_override", C.c_float), ("speed_dev_override", C.c_float), ("vcap", C.c_int32),
           ("lane_change", C.c_int32), ("record_trips", C.c_int32), ("tile_vcap", C.c_int32)]


class RsScenario(C.Structure):
    _fields_ = ([("abi_version", C.c_int32)] + [(n, C.c_int32) for n in _SIZES] + _PTRS + _PARAMS)


class RsObsView(C.Structure):
    _fields_ = [("n_env", C.c_int32), ("n_signals", C.c_int32), ("n_sig_lanes", C.c_int32),
                ("lane_queue", C.c_void_p), ("lane_approach", C.c_void_p), ("lane_total_wait", C.c_void_p),
                ("lane_max_wait", C.c_void_p), ("lane_speed_sum", C.c_void_p), ("phase", C.c_void_p),
                ("mplight", C.c_void_p), ("wave", C.c_void_p), ("reward_wait", C.c_void_p),
                ("reward_wait_norm", C.c_void_p), ("reward_pressure", C.c_void_p),
                ("sig_queue_len", C.c_void_p), ("sig_max_queue", C.c_void_p), ("lane_arrivals", C.c_void_p),
                ("drq", C.c_void_p), ("drq_norm", C.c_void_p), ("mplight_full", C.c_void_p)]

RS_OUT_DRQ, RS_OUT_DRQ_NORM, RS_OUT_MPLIGHT_FULL = 1, 2, 4
HOSTOBS = {"mplight": 0, "wave": 1, "drq_norm": 2, "drq": 3, "mplight_full": 4}


class RsStats(C.Structure):
    _fields_ = [("tick", C.c_int32), ("n_active", C.c_int32), ("n_inserted", C.c_int32),
                ("n_arrived", C.c_int32), ("n_backlog", C.c_int32), ("anomalies", C.c_int32),
                ("sum_delay_arrived", C.c_float), ("sum_delay_running", C.c_float),
                ("sum_delay_pending", C.c_float), ("sum_duration_arrived", C.c_float),
                ("sum_wait_arrived", C.c_float), ("sum_active_ticks", C.c_int32), ("n_cap_refused", C.c_int32)]


STATS_DTYPE = np.dtype([(n, np.int32 if t is C.c_int32 else np.float32) for n, t in RsStats._fields_])


# ------------------------------------------------------------------------------------------------
def create_yellows(green_states: Sequence[str]) -> Tuple[List[str], Dict[str, int]]:
    """Yellow-phase synthesis, restating ``traffic_signal.py:7-24``.

    For every ordered green pair (i != j) the yellow string is phase i with 'G'/'g' -> 'y' wherever
    phase j shows 'r'/'s'; it is appended (after the greens, in (i, j) order) only if at least one
    link changes.  Returns (yellow state strings, {"i_j": index in greens+yellows}).
    """
    n = len(green_states)
    yellows: List[str] = []
    ydict: Dict[str, int] = {}
    for i in range(n):
        for j in range(n):
            if i == j:
                continue
            need, ys = False, []
            for a, b in zip(green_states[i], green_states[j]):
                if a in 'Gg' and b in 'rs':
                    need = True
                    ys.append('y')
                else:
                    ys.append(a)
            if need:
                yellows.append(''.join(ys))
                ydict[f"{i}_{j}"] = n + len(yellows) - 1
    return yellows, ydict


class Marshalled:
    """RsScenario + the numpy arrays that back its pointers (kept alive together)."""

    def __init__(self, struct: RsScenario, keep: Dict[str, np.ndarray], info: Dict[str, object]):
        self.struct = struct
        self.keep = keep
        self.info = info


def _ptr(arr: np.ndarray, ctype):
    return arr.ctypes.data_as(ctype)


def marshal(sc: Scenario, *, step_length: int = 10, yellow_length: int = 3, max_distance: float = 200.0,
            end_time: Optional[float] = None, controlled: bool = True, sigma: float = -1.0,
            speed_dev: float = -1.0, vcap: int = 0, lane_change: bool = True, record_trips: bool = False,
            synthetic: Optional[Dict[str, np.ndarray]] = None, tile_vcap: int = 0) -> Marshalled:
    """Build the C struct.  ``controlled=False`` keeps every tlLogic on its original program
    (the reference's FIXED rows: SUMO default programs, no Signal objects)."""
    a = sc.arrays
    meta = sc.meta
    keep: Dict[str, np.ndarray] = {}

    def put(name: str, arr, dtype) -> np.ndarray:
        x = np.ascontiguousarray(np.asarray(arr, dtype=dtype).reshape(-1))
        if x.size == 0:
            x = np.zeros(1, dtype)      # never hand out NULL
        keep[name] = x
        return x

    L = sc.n_lanes
    tls_ids: List[str] = meta["tls_ids"]
    sig_ids: List[str] = meta["signal_ids"] if controlled else []
    sig_tls = a["sig_tls"] if controlled else np.zeros(0, np.int32)
    S = len(sig_ids)
    begin = float(meta["begin"])

    # ---- installed programs ------------------------------------------------------------------
    tls_phase_off = [0]
    phase_dur: List[int] = []
    phase_state_off: List[int] = []
    chars = bytearray()
    tls_nlinks, tls_init_phase, tls_init_left = [], [], []
    sig_n_green, sig_yellow_off, yellow_idx = [], [0], []
    programs_installed: Dict[str, List[Tuple[int, str]]] = {}
    yellow_dicts: Dict[str, Dict[str, int]] = {}
    green_states_of: Dict[str, List[str]] = {}
    ctrl = set(sig_ids)
    for t in tls_ids:
        prog = meta["programs"][t]
        if t in ctrl:
            gidx = green_phase_indices(prog)
            greens = [(int(round(prog[i][0])), prog[i][1]) for i in gidx]
            ystates, ydict = create_yellows([g[1] for g in greens])
            inst = greens + [(int(yellow_length), y) for y in ystates]
            init_phase, init_left = 0, inst[0][0]
            yellow_dicts[t] = ydict
            green_states_of[t] = [g[1] for g in greens]
        else:
            inst = [(int(round(d)), st) for d, st in prog]
            cycle = sum(d for d, _ in inst)
            # SUMO starts a static program where it would be had it run since t=0 (offset 0 here)
            pos = int(begin) % cycle if cycle > 0 else 0
            init_phase, init_left = 0, inst[0][0]
            acc = 0
            for i, (d, _) in enumerate(inst):
                if pos < acc + d:
                    init_phase, init_left = i, acc + d - pos
                    break
                acc += d
        programs_installed[t] = inst
        n = len(inst[0][1])
        for d, st in inst:
            phase_dur.append(max(int(d), 1))
            phase_state_off.append(len(chars))
            chars.extend(st.encode())
        tls_phase_off.append(len(phase_dur))
        tls_nlinks.append(n)
        tls_init_phase.append(init_phase)
        tls_init_left.append(init_left)
    for s in sig_ids:
        ng = len(green_states_of[s])
        sig_n_green.append(ng)
        tbl = np.full((ng, ng), -1, np.int32)
        for key, v in yellow_dicts[s].items():
            i, j = key.split('_')
            tbl[int(i), int(j)] = v
        yellow_idx.extend(tbl.reshape(-1).tolist())
        sig_yellow_off.append(len(yellow_idx))

    lane_sig = np.full(L, -1, np.int32)
    lane_sig_slot = np.full(L, -1, np.int32)
    if controlled:
        for s in range(S):
            for slot, q in enumerate(range(a["sig_lane_off"][s], a["sig_lane_off"][s + 1])):
                lane_sig[a["sig_lane"][q]] = s
                lane_sig_slot[a["sig_lane"][q]] = slot

    st = RsScenario()
    st.abi_version = RS_ABI_VERSION
    n_trips = int(a["trip_depart"].shape[0]) if "trip_depart" in a else 0
    sizes = dict(
        n_lanes=L, n_edges=len(a["edge_lane0"]), n_links=len(a["link_from"]), n_foes=len(a["foe_link"]),
        n_tls=len(tls_ids), n_phases=len(phase_dur), n_state_chars=len(chars), n_signals=S,
        n_sig_lanes=int(a["sig_lane_off"][S]) if controlled else 0,
        n_mv_lanes=len(a["mv_lane"]) if controlled else 0, n_mvo=len(a["mvo_sig"]) if controlled else 0,
        n_out=len(a["out_sig"]) if controlled else 0, n_yellow=len(yellow_idx),
        n_vtypes=len(a["vtype_bit"]), n_routes=len(a["route_off"]) - 1, n_route_steps=len(a["route_edge"]),
        n_origins=len(a["origin_lane"]), n_trips=n_trips, n_origin_routes=0,
        n_watch=len(a["origin_watch_lane"]), n_lane_watch=len(a["lane_watch_lane"]))
    arrays: Dict[str, Tuple[object, object]] = {}
    for name, ct in _PTRS:
        if name in a:
            arrays[name] = (a[name], np.float32 if ct is _F32P else np.int32)
    arrays.update(
        lane_sig=(lane_sig, np.int32), lane_sig_slot=(lane_sig_slot, np.int32),
        tls_phase_off=(tls_phase_off, np.int32), tls_nlinks=(tls_nlinks, np.int32),
        tls_init_phase=(tls_init_phase, np.int32), tls_init_left=(tls_init_left, np.int32),
        phase_dur=(phase_dur, np.int32), phase_state_off=(phase_state_off, np.int32),
        state_chars=(np.frombuffer(bytes(chars), np.uint8), np.uint8),
        sig_tls=(sig_tls, np.int32), sig_n_green=(sig_n_green, np.int32),
        sig_yellow_off=(sig_yellow_off, np.int32), yellow_idx=(yellow_idx, np.int32),
        origin_watch_owner=(np.repeat(np.arange(len(a["origin_lane"])), np.diff(a["origin_watch_off"])), np.int32),
        origin_rate=(np.zeros(1, np.int32), np.int32), origin_route_off=(np.zeros(sizes["n_origins"] + 1, np.int32), np.int32),
        origin_route=(np.zeros(1, np.int32), np.int32))
    if not controlled:
        for nm in ("sig_lane_off", "mv_off", "mvo_off", "out_off"):
            arrays[nm] = (np.zeros(1, np.int32), np.int32)
    if synthetic is not None:
        for nm in ("origin_lane", "origin_rate", "origin_route_off", "origin_route", "origin_watch_off",
                   "origin_watch_lane"):
            arrays[nm] = (synthetic[nm], np.int32)
        arrays["origin_watch_dist"] = (synthetic["origin_watch_dist"], np.float32)
        sizes["n_watch"] = len(synthetic["origin_watch_lane"])
        arrays["origin_watch_owner"] = (np.repeat(np.arange(len(synthetic["origin_lane"])),
                                                  np.diff(synthetic["origin_watch_off"])), np.int32)
        sizes["n_origins"] = len(synthetic["origin_lane"])
        sizes["n_origin_routes"] = len(synthetic["origin_route"])
        sizes["n_trips"] = 0
        arrays["origin_off"] = (np.zeros(sizes["n_origins"] + 1, np.int32), np.int32)
        if "route_off" in synthetic:     # the synthetic demand brings its own route / vType tables
            for nm in ("route_off", "route_edge", "route_mask", "vtype_bit"):
                arrays[nm] = (synthetic[nm], np.int32)
            arrays["vtype"] = (synthetic["vtype_table"], np.float32)
            sizes["n_routes"] = len(synthetic["route_off"]) - 1
            sizes["n_route_steps"] = len(synthetic["route_edge"])
            sizes["n_vtypes"] = len(synthetic["vtype_bit"])
        for nm in ("trip_depart",):
            arrays[nm] = (np.zeros(1, np.float32), np.float32)
        for nm in ("trip_route", "trip_vtype", "trip_file", "trip_depart_pos"):
            arrays[nm] = (np.zeros(1, np.int32), np.int32)
    if "trip_depart_pos" not in arrays:        # scenarios compiled before the field existed: departPos="base" everywhere
        arrays["trip_depart_pos"] = (np.zeros(max(n_trips, 1), np.int32), np.int32)
    for k, v in sizes.items():
        setattr(st, k, int(v))
    for name, ct in _PTRS:
        arr, dt = arrays[name]
        x = put(name, arr, dt)
        setattr(st, name, _ptr(x, ct))
    end_t = float(end_time) if end_time is not None else float(meta["map_config"]["end_time"])
    st.synthetic = 1 if synthetic is not None else 0
    st.synthetic_vtype = int(synthetic.get("vtype", 0)) if synthetic is not None else 0
    st.step_length = int(step_length)
    st.yellow_length = int(yellow_length)
    st.end_tick = int(round(end_t - begin))
    st.max_distance = float(max_distance)
    st.sigma_override = float(sigma)
    st.speed_dev_override = float(speed_dev)
    if vcap <= 0:
        vcap = default_vcap(sc)
    st.vcap = int(vcap)
    st.lane_change = 1 if lane_change else 0
    st.record_trips = 1 if (record_trips and synthetic is None) else 0
    st.tile_vcap = int(tile_vcap)
    info = dict(programs_installed=programs_installed, yellow_dicts=yellow_dicts, signal_ids=sig_ids,
                tls_ids=tls_ids, green_states=green_states_of, vcap=int(vcap), sizes=sizes)
    return Marshalled(st, keep, info)


def smem_bytes(vcap: int, n_lanes: int, n_tls: int, n_signals: int, n_origins: int, n_sig_lanes: int,
               n_vtypes: int) -> int:
    """Mirror of rs::make_layout (resco_b200/csrc/sim.cu): shared memory of one instance with a ping-pong tile of
    `vcap` vehicles in shared memory."""
    def al(x):
        return (x + 15) & ~15
    o = 0
    o = al(o + 11 * vcap * 4)
    o = al(o + 11 * vcap * 4)
    o = al(o + vcap * 4)
    for _ in range(4):
        o = al(o + vcap * 2)
    o = al(o + (2 * vcap + max(n_origins, 1)) * 2)
    o = al(o + (n_lanes + 1) * 2)
    o = al(o + (n_lanes + 1) * 2)
    o = al(o + n_lanes * 4)
    o = al(o + n_lanes * 4)
    o = al(o + n_tls * 4)
    o = al(o + n_tls * 4)
    o = al(o + n_tls * 4)
    o = al(o + max(n_signals, 1) * 4)
    o = al(o + max(n_origins, 1) * 4)
    o = al(o + max(n_origins, 1) * 4)
    o = al(o + max(n_origins, 1) * 16)
    o = al(o + n_vtypes * 32)
    o = al(o + 16 * 4)
    o = al(o + 48 * 4)
    if n_sig_lanes * 20 > vcap * 6:      # else the observe scratch shares the plan scratch
        o = al(o + max(n_sig_lanes, 1) * 20)
    o = al(o + 16)
    o = al(o + max(n_origins, 1) * 2)
    o = al(o + ((n_lanes + 31) // 32 + 2) * 4)
    return o


def default_vcap(sc: Scenario) -> int:
    """Capacity of an instance's vehicle store in HBM: a quarter of the jam capacity of the normal lanes, rounded up to
    a multiple of 64, clamped to [256, 4096] (44 B per vehicle and instance).  The shared-memory tile the kernel works
    on is smaller (RsScenario.tile_vcap); instances that outgrow it are stepped out of a global-memory workspace, so
    this number bounds memory, not speed."""
    a = sc.arrays
    normal = a["lane_internal"] == 0
    jam = float(np.sum(np.floor(a["lane_len"][normal] / 7.5) + 1))
    v = int(np.ceil(jam / 4 / 64.0)) * 64
    return int(min(max(v, 256), 4096))
