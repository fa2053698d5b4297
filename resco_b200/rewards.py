"""Reward callables with the reference's names and signatures (``f(signals) -> dict[id -> scalar]``),
plus ``batched(env)`` device-tensor variants ([N, S]) read from the fused kernel's outputs."""
from __future__ import annotations

import numpy as np


def _sum_over_lanes(sig, key):
    return sum(sig.full_observation[lane][key] for lane in sig.lanes)


def wait(signals):
    """rewards.py:6-14."""
    return {sid: -_sum_over_lanes(sig, 'total_wait') for sid, sig in signals.items()}


def wait_norm(signals):
    """rewards.py:17-25 -- clip(-sum(total_wait)/224, -4, 4) as float32."""
    return {sid: np.clip(-_sum_over_lanes(sig, 'total_wait') / 224, -4, 4).astype(np.float32)
            for sid, sig in signals.items()}


def pressure(signals):
    """rewards.py:28-41 -- -(inbound queue - queue on the lanes feeding controlled downstream signals)."""
    out = dict()
    for sid, sig in signals.items():
        q = _sum_over_lanes(sig, 'queue')
        for lane in sig.outbound_lanes:
            dwn = sig.out_lane_to_signalid[lane]
            if dwn in sig.signals:
                q -= sig.signals[dwn].full_observation[lane]['queue']
        out[sid] = -q
    return out


wait.batched = lambda env: env.sim.obs_view()["reward_wait"]
wait_norm.batched = lambda env: env.sim.obs_view()["reward_wait_norm"]
pressure.batched = lambda env: env.sim.obs_view()["reward_pressure"]
