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


def _fma2c(signals, key):
    from .states import _mdp, _region_fringes
    cfg = _mdp(signals, key)
    supervisors, neighbors_of = cfg['supervisors'], cfg['management_neighbors']
    fringes = _region_fringes(signals, cfg)
    fringe_arrivals = {mgr: 0 for mgr in cfg['management']}
    liquidity = {mgr: 0 for mgr in cfg['management']}
    for sid, sig in signals.items():
        mgr = supervisors[sid]
        arrivals = sig.full_observation['arrivals']
        liquidity[mgr] += len(sig.full_observation['departures']) - len(arrivals)
        for lane in sig.lanes:
            if lane in fringes[mgr]:
                fringe_arrivals[mgr] += sum(1 for v in sig.full_observation[lane]['vehicles'] if v['id'] in arrivals)
    managers = dict()
    for mgr in cfg['management']:
        r = fringe_arrivals[mgr] + liquidity[mgr]
        for nb in neighbors_of[mgr]:
            r += cfg['alpha'] * (fringe_arrivals[nb] + liquidity[nb])
        managers[mgr] = r
    own = dict()
    for sid, sig in signals.items():
        r = 0
        for lane in sig.lanes:
            r += sig.full_observation[lane]['queue']
            r += sig.full_observation[lane]['max_wait'] * cfg['coef']
        own[sid] = -r
    out = dict()
    for sid, sig in signals.items():
        tot = own[sid]
        for neighbor in sig.downstream.values():
            if neighbor is not None and supervisors[neighbor] == supervisors[sid]:
                tot += cfg['alpha'] * own[neighbor]
        out[sid] = tot
    out.update(managers)
    return out


def fma2c(signals):
    """rewards.py:72-136 -- worker: -(queue + coef*max_wait) with same-region neighbours; manager:
    fringe arrivals + liquidity (departures - arrivals) with neighbouring regions."""
    return _fma2c(signals, 'FMA2C')


def fma2c_full(signals):
    """rewards.py:139-202 (same rule with the FMA2CFull hyper-parameters)."""
    return _fma2c(signals, 'FMA2CFull')


wait.batched = lambda env: env.sim.obs_view()["reward_wait"]
wait_norm.batched = lambda env: env.sim.obs_view()["reward_wait_norm"]
pressure.batched = lambda env: env.sim.obs_view()["reward_pressure"]
