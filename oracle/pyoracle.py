"""ctypes loader for the CPU oracle (oracle/microsim.c).  TEST INFRASTRUCTURE ONLY.

May be imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs -- never by anything under resco_b200/.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(_HERE))
from resco_b200.abi import RsScenario, RsStats, STATS_DTYPE, Marshalled  # noqa: E402

_LIB = None


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "microsim.c")
    hdr = os.path.join(_HERE, "..", "include", "resco_b200.h")
    if force or not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(os.environ.get("ORACLE_LIB") or build())   # ORACLE_LIB: an experimental build of the rule set
        _LIB.orc_create.restype = C.c_void_p
        _LIB.orc_create.argtypes = [C.POINTER(RsScenario), C.c_int32, C.c_uint64]
        _LIB.orc_destroy.argtypes = [C.c_void_p]
        _LIB.orc_reset.argtypes = [C.c_void_p, C.c_uint64, C.c_int64]
        _LIB.orc_set_phase.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        _LIB.orc_set_demand_window.argtypes = [C.c_void_p, C.c_void_p]
        _LIB.orc_tick.argtypes = [C.c_void_p, C.c_int32]
        _LIB.orc_observe.argtypes = [C.c_void_p]
        _LIB.orc_env_step.argtypes = [C.c_void_p, C.c_void_p]
        _LIB.orc_get_obs.argtypes = [C.c_void_p] + [C.c_void_p] * 17
        _LIB.orc_get_stats.argtypes = [C.c_void_p, C.c_void_p]
        _LIB.orc_dump_vehicles.restype = C.c_int
        _LIB.orc_dump_vehicles.argtypes = [C.c_void_p, C.c_int32] + [C.c_void_p] * 14
        _LIB.orc_get_phases.argtypes = [C.c_void_p, C.c_int32, C.c_void_p]
        _LIB.orc_get_trip_records.restype = C.c_int
        _LIB.orc_get_trip_records.argtypes = [C.c_void_p, C.c_int32] + [C.c_void_p] * 5
        for f in ("orc_brake_gap", "orc_max_safe_stop_speed", "orc_free_speed"):
            getattr(_LIB, f).restype = C.c_float
            getattr(_LIB, f).argtypes = [C.c_float] * 3
        _LIB.orc_follow_speed.restype = C.c_float
        _LIB.orc_follow_speed.argtypes = [C.c_float] * 5
        _LIB.orc_philox.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32]
    return _LIB


OBS_FIELDS = [("lane_queue", np.float32, "L"), ("lane_approach", np.float32, "L"),
              ("lane_total_wait", np.float32, "L"), ("lane_max_wait", np.float32, "L"),
              ("lane_speed_sum", np.float32, "L"), ("phase", np.int32, "S"), ("mplight", np.float32, "S13"),
              ("wave", np.float32, "S12"), ("reward_wait", np.float32, "S"), ("reward_wait_norm", np.float32, "S"),
              ("reward_pressure", np.float32, "S"), ("sig_queue_len", np.int32, "S"), ("sig_max_queue", np.int32, "S"), ("lane_arrivals", np.float32, "L"),
              ("drq", np.float32, "L5"), ("drq_norm", np.float32, "L5"), ("mplight_full", np.float32, "S49")]

VEH_FIELDS = [("lane", np.int32), ("pos", np.float32), ("speed", np.float32), ("accel", np.float32),
              ("wait", np.float32), ("rwait", np.float32), ("tloss", np.float32), ("vid", np.int32),
              ("vtype", np.int32), ("route", np.int32), ("cursor", np.int32), ("sf", np.float32),
              ("depart", np.int32), ("acc_wait", np.float32)]


class OracleSim:
    """Same call shape as resco_b200.sim.VecSim (so parity tests read alike)."""

    def __init__(self, m: Marshalled, n_env: int, seed: int = 0):
        self.m = m
        self.n_env = n_env
        self.S = m.struct.n_signals
        self.SL = m.struct.n_sig_lanes
        self.vcap = m.struct.vcap
        self._h = lib().orc_create(C.byref(m.struct), n_env, seed)
        if not self._h:
            raise RuntimeError("orc_create failed")

    def close(self):
        if self._h:
            lib().orc_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reset(self, seed: int = 0, first_env_id: int = 0):
        lib().orc_reset(self._h, seed, first_env_id)

    def set_demand_window(self, origin_off):
        w = np.ascontiguousarray(origin_off, np.int32)
        assert w.shape == (self.m.struct.n_origins + 1,)
        if lib().orc_set_demand_window(self._h, w.ctypes.data) != 0:
            raise RuntimeError("set_demand_window: synthetic demand has no trip table")

    def set_phase(self, phase, mask=None):
        p = np.ascontiguousarray(phase, np.int32)
        mk = None if mask is None else np.ascontiguousarray(mask, np.uint8)
        lib().orc_set_phase(self._h, p.ctypes.data, None if mk is None else mk.ctypes.data)

    def tick(self, n: int = 1):
        lib().orc_tick(self._h, n)

    def observe(self):
        lib().orc_observe(self._h)

    def env_step(self, actions):
        a = np.ascontiguousarray(actions, np.int32).reshape(self.n_env, self.S)
        lib().orc_env_step(self._h, a.ctypes.data)

    def obs(self):
        out = {}
        ptrs = []
        for name, dt, shp in OBS_FIELDS:
            shape = {"L": (self.n_env, self.SL), "S": (self.n_env, self.S), "S13": (self.n_env, self.S, 13),
                     "S12": (self.n_env, self.S, 12), "L5": (self.n_env, self.SL, 5), "S49": (self.n_env, self.S, 49)}[shp]
            out[name] = np.zeros(shape, dt)
            ptrs.append(out[name].ctypes.data)
        lib().orc_get_obs(self._h, *ptrs)
        return out

    def select_outputs(self, *names):      # the oracle always computes every tensor
        pass

    def obs_view(self):
        """Same keys and shapes as VecSim.obs_view(), as CPU torch tensors (copies): lets the CPU test tier run the
        `.batched(env)` state / reward expressions of resco_b200.states / rewards on the oracle."""
        import torch
        return {k: torch.from_numpy(v) for k, v in self.obs().items()}

    def stats(self):
        st = np.zeros(self.n_env, STATS_DTYPE)
        lib().orc_get_stats(self._h, st.ctypes.data)
        return st

    def vehicles(self, env: int = 0):
        arrs = {n: np.zeros(self.vcap, dt) for n, dt in VEH_FIELDS}
        n = lib().orc_dump_vehicles(self._h, env, *[arrs[k].ctypes.data for k, _ in VEH_FIELDS])
        return {k: v[:n] for k, v in arrs.items()}

    def trip_records(self, env: int = 0):
        n = self.m.struct.n_trips
        out = dict(arrival=np.zeros(n, np.int32), depart=np.zeros(n, np.int32), time_loss=np.zeros(n, np.float32),
                   depart_delay=np.zeros(n, np.int32), waiting_time=np.zeros(n, np.float32))
        rc = lib().orc_get_trip_records(self._h, env, out["arrival"].ctypes.data, out["depart"].ctypes.data,
                                        out["time_loss"].ctypes.data, out["depart_delay"].ctypes.data,
                                        out["waiting_time"].ctypes.data)
        if rc != 0:
            raise RuntimeError("trip records were not enabled (marshal(record_trips=True))")
        return out

    def phases(self, env: int = 0):
        p = np.zeros(self.m.struct.n_tls, np.int32)
        lib().orc_get_phases(self._h, env, p.ctypes.data)
        return p
