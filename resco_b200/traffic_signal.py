"""``Signal`` with the reference's class surface (traffic_signal.py:27-247), backed by the vectorised
simulator instead of per-vehicle TraCI calls.

The phase machine (prep_phase / set_phase), the observation sweep (observe) and the waiting-time
latch run inside the fused device kernel; this object is the per-instance *view* that the reference's
state / reward callables and agents read: ``lanes`` (row order of the observation), ``lane_sets``,
``lane_sets_outbound``, ``downstream``, ``outbound_lanes``, ``out_lane_to_signalid``,
``full_observation``, ``phase``, ``signals``.
"""
from __future__ import annotations

from typing import Dict, List

from .abi import create_yellows  # noqa: F401  (re-exported under the reference's module name)


class Signal:
    def __init__(self, env, sig_id: str, index: int):
        meta = env.scenario.meta["signals"][sig_id]
        self._env = env
        self._index = index
        self.id = sig_id
        self.yellow_time = env.yellow_length
        self.next_phase = 0
        self.lanes: List[str] = list(meta["lanes"])
        self.lane_sets: Dict[str, List[str]] = {k: list(v) for k, v in meta["lane_sets"].items()}
        self.lane_sets_outbound = {k: list(v) for k, v in meta["lane_sets_outbound"].items()}
        self.downstream = dict(meta["downstream"])
        self.outbound_lanes = list(meta["outbound_lanes"])
        self.out_lane_to_signalid = dict(meta["out_lane_to_signalid"])
        self.inbounds_fr_direction = {k: list(v) for k, v in meta["inbounds_fr_direction"].items()}
        info = env.marshalled.info
        self.phases = [tuple(p) for p in info["programs_installed"][sig_id]]      # greens + yellows
        self.yellow_dict = dict(info["yellow_dicts"][sig_id])
        self.green_phases = list(info["green_states"][sig_id])
        self.waiting_times: Dict[str, float] = dict()
        self.signals = None
        self.full_observation = None
        self.last_step_vehicles = None

    @property
    def phase(self) -> int:
        return int(self._env._phase_of(self._index))

    # The three methods below exist for callers that drive a Signal by hand (N = 1).
    def prep_phase(self, new_phase):
        cur = self.phase
        if cur == new_phase:
            self.next_phase = cur
        else:
            self.next_phase = new_phase
            key = str(cur) + '_' + str(new_phase)
            if key in self.yellow_dict:
                self._env._set_phase_one(self._index, self.yellow_dict[key])

    def set_phase(self):
        self._env._set_phase_one(self._index, int(self.next_phase))

    def observe(self, step_length=None, distance=None):
        self._env._observe_into_signals()
