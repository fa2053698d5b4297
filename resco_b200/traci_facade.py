"""TraCI-subset facade over one simulator instance (N = 1).

The reference reaches the simulator only through ``self.sumo`` (multi_signal.py:44,47,134,137;
traffic_signal.py:29).  This object implements exactly the calls listed in SURVEY.md §8(b) so that the
UNMODIFIED reference ``MultiSignal`` / ``Signal`` classes can run on the B200 backend (or, in the
tests, on the CPU oracle) -- it is the drop-in seam, and the parity harness.

Backend protocol (``resco_b200.sim.VecSim`` and the test-only ``pyoracle.OracleSim`` both satisfy it):
``reset(seed, first_env_id)``, ``tick(n)``, ``set_phase(phase[N,S], mask[N,S])``, ``vehicles(env)``,
``phases(env)``, ``stats()``.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Tuple

import numpy as np

from .abi import Marshalled, marshal
from .scenario.compiler import Scenario


class Phase:
    """traci.trafficlight.Phase(duration, state) -- constructor used by traffic_signal.py:22."""

    def __init__(self, duration, state, minDur=-1, maxDur=-1, next=(), name=""):
        self.duration = duration
        self.state = state
        self.minDur = minDur
        self.maxDur = maxDur
        self.next = next
        self.name = name

    def __repr__(self):
        return f"Phase(duration={self.duration}, state='{self.state}')"


class Logic:
    def __init__(self, programID, type, currentPhaseIndex, phases):
        self.programID = programID
        self.type = type
        self.currentPhaseIndex = currentPhaseIndex
        self.phases = list(phases)

    def getPhases(self):
        return self.phases


class _TrafficLightDomain:
    Phase = Phase
    Logic = Logic

    def __init__(self, f: "TraciFacade"):
        self._f = f

    def getIDList(self):
        return tuple(self._f.sc.meta["tls_ids"])

    def getAllProgramLogics(self, tls_id):
        f = self._f
        if tls_id in f.installed:
            prog = f.m.info["programs_installed"][tls_id]
        else:
            prog = [(d, s) for d, s in f.sc.meta["programs"][tls_id]]
        return [Logic("0", 0, 0, [Phase(d, s) for d, s in prog])]

    def setProgramLogic(self, tls_id, logic):
        """Signal.__init__ installs greens+yellows (traffic_signal.py:93-100).  The device program
        table was built from the same net with our own yellow synthesis; the two must agree."""
        f = self._f
        want = [(int(p.duration), p.state) for p in logic.phases]
        have = [(int(d), s) for d, s in f.m.info["programs_installed"][tls_id]]
        if [s for _, s in want] != [s for _, s in have]:
            raise ValueError(f"setProgramLogic({tls_id}): program differs from the compiled one:\n{want}\n{have}")
        f.installed.add(tls_id)

    def getPhase(self, tls_id):
        return int(self._f.backend.phases(0)[self._f.tls_index[tls_id]])

    def setPhase(self, tls_id, index):
        f = self._f
        if tls_id not in f.sig_index:
            raise KeyError(f"{tls_id} is not a controlled signal of this scenario")
        S = len(f.sig_index)
        ph = np.zeros((1, S), np.int32)
        mk = np.zeros((1, S), np.uint8)
        ph[0, f.sig_index[tls_id]] = int(index)
        mk[0, f.sig_index[tls_id]] = 1
        f.backend.set_phase(ph, mk)

    def getControlledLinks(self, tls_id):
        return [[tuple(x) for x in slot] for slot in self._f.sc.meta["controlled_links"][tls_id]]


class _SimulationDomain:
    def __init__(self, f):
        self._f = f

    def getTime(self):
        return float(self._f.sc.meta["begin"]) + float(self._f.tick)


class _LaneDomain:
    def __init__(self, f):
        self._f = f

    def getLastStepVehicleIDs(self, lane_id):
        f = self._f
        v = f.snapshot()
        li = f.lane_index[lane_id]
        return tuple(f.veh_name(int(x)) for x in v["vid"][v["lane"] == li])


class _VehicleDomain:
    def __init__(self, f):
        self._f = f

    def _row(self, veh_id) -> int:
        return self._f.snapshot_index()[veh_id]

    def getNextTLS(self, veh_id):
        f = self._f
        v = f.snapshot()
        i = self._row(veh_id)
        lane = int(v["lane"][i])
        a = f.sc.arrays
        td = float(a["lane_tls_dist"][lane])
        if td < 0:
            return []
        dist = np.float32(np.float32(a["lane_len"][lane]) - v["pos"][i]) + np.float32(td)
        return [(f.lane_next_tls.get(lane, ""), 0, float(dist), "G")]

    def getWaitingTime(self, veh_id):
        return float(self._f.snapshot()["wait"][self._row(veh_id)])

    def getSpeed(self, veh_id):
        return float(self._f.snapshot()["speed"][self._row(veh_id)])

    def getAcceleration(self, veh_id):
        f = self._f
        i = self._row(veh_id)
        v = f.snapshot()
        prev = f.prev_speed.get(int(v["vid"][i]), 0.0)
        return float(v["speed"][i]) - float(prev)

    def getLanePosition(self, veh_id):
        return float(self._f.snapshot()["pos"][self._row(veh_id)])

    def getTypeID(self, veh_id):
        f = self._f
        return f.sc.meta["vtype_ids"][int(f.snapshot()["vtype"][self._row(veh_id)])]


class TraciFacade:
    """One connection (= one simulator instance)."""

    def __init__(self, sc: Scenario, m: Marshalled, backend, seed: int = 0):
        self.sc, self.m, self.backend = sc, m, backend
        self.tls_index = {t: i for i, t in enumerate(sc.meta["tls_ids"])}
        self.sig_index = {s: i for i, s in enumerate(m.info["signal_ids"])}
        self.lane_index = {l: i for i, l in enumerate(sc.meta["lane_ids"])}
        self.installed = set()
        self.tick = 0
        self._snap = None
        self._snap_index = None
        self.prev_speed: Dict[int, float] = {}
        self.lane_next_tls: Dict[int, str] = {}
        a = sc.arrays
        for s, sid in enumerate(m.info["signal_ids"]):
            for q in range(a["sig_lane_off"][s], a["sig_lane_off"][s + 1]):
                self.lane_next_tls[int(a["sig_lane"][q])] = sid
        self.trafficlight = _TrafficLightDomain(self)
        self.simulation = _SimulationDomain(self)
        self.lane = _LaneDomain(self)
        self.vehicle = _VehicleDomain(self)
        backend.reset(seed, 0)

    def veh_name(self, vid: int) -> str:
        ids = self.sc.meta.get("trip_ids")
        return ids[vid] if ids is not None and 0 <= vid < len(ids) else f"veh{vid}"

    def snapshot(self):
        if self._snap is None:
            self._snap = self.backend.vehicles(0)
            self._snap_index = None
        return self._snap

    def snapshot_index(self):
        if self._snap_index is None:
            v = self.snapshot()
            self._snap_index = {self.veh_name(int(x)): i for i, x in enumerate(v["vid"])}
        return self._snap_index

    def simulationStep(self, step=0):
        v = self.snapshot()
        self.prev_speed = {int(i): float(s) for i, s in zip(v["vid"], v["speed"])}
        self.backend.tick(1)
        self.tick += 1
        self._snap = None
        self._snap_index = None

    def close(self):
        self._snap = None


def open_facade(sc: Scenario, backend_factory: Callable[[Marshalled], object], *, step_length=10, yellow_length=3,
                max_distance=200.0, seed=0, controlled=True, **kw) -> TraciFacade:
    m = marshal(sc, step_length=step_length, yellow_length=yellow_length, max_distance=max_distance,
                controlled=controlled, **kw)
    return TraciFacade(sc, m, backend_factory(m), seed)
