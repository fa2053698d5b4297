"""Observation callables with the reference's names and signatures (``f(signals) -> dict``).

Each function accepts the per-instance dict view (``signals``: id -> Signal, exactly what
``resco_benchmark/states.py`` receives) and also carries a ``batched(env)`` attribute that returns
the same quantity for all N instances as a device tensor built from the fused kernel's outputs.
Reference quirks are reproduced on purpose: ``drq*`` compares the LANE index with the phase index
(states.py:40-45); ``mplight_full`` keeps only the last lane's speed sum (states.py:97).
"""
from __future__ import annotations

import numpy as np


def _speed_sum(signal, lane):
    tot = 0
    for veh in signal.full_observation[lane]['vehicles']:
        tot += veh['speed']
    return tot


def _speed_sum_norm(signal, lane):
    tot = 0
    for veh in signal.full_observation[lane]['vehicles']:
        tot += (veh['speed'] / 20 / 28)
    return tot


def _drq_rows(signals, norm):
    out = dict()
    for sid, sig in signals.items():
        act = sig.phase
        rows = []
        for i, lane in enumerate(sig.lanes):
            fo = sig.full_observation[lane]
            onehot = 1 if i == act else 0
            if norm:
                rows.append([onehot, fo['approach'] / 28, fo['total_wait'] / 28, fo['queue'] / 28,
                             _speed_sum_norm(sig, lane)])
            else:
                rows.append([onehot, fo['approach'], fo['total_wait'], fo['queue'], _speed_sum(sig, lane)])
        out[sid] = np.expand_dims(np.asarray(rows), axis=0)
    return out


def drq(signals):
    """states.py:6-31 -- per signal [1, n_lanes, 5]."""
    return _drq_rows(signals, norm=False)


def drq_norm(signals):
    """states.py:34-59."""
    return _drq_rows(signals, norm=True)


def _pressure_of(sig, direction):
    q = 0
    for lane in sig.lane_sets[direction]:
        q += sig.full_observation[lane]['queue']
    for lane in sig.lane_sets_outbound[direction]:
        dwn = sig.out_lane_to_signalid[lane]
        if dwn in sig.signals:
            q -= sig.signals[dwn].full_observation[lane]['queue']
    return q


def mplight(signals):
    """states.py:62-80 -- [phase, 12 x (inbound queue - downstream queue)]."""
    out = dict()
    for sid, sig in signals.items():
        out[sid] = np.asarray([sig.phase] + [_pressure_of(sig, d) for d in sig.lane_sets])
    return out


def mplight_full(signals):
    """states.py:83-113 -- [phase, 12 x (pressure, wait/28, last-lane speed sum, approach/28)]."""
    out = dict()
    for sid, sig in signals.items():
        obs = [sig.phase]
        for d in sig.lane_sets:
            total_wait, total_speed, tot_approach = 0, 0, 0
            for lane in sig.lane_sets[d]:
                fo = sig.full_observation[lane]
                total_wait += fo['total_wait'] / 28
                total_speed = _speed_sum(sig, lane)      # reset per lane: reference behaviour
                tot_approach += fo['approach'] / 28
            obs += [_pressure_of(sig, d), total_wait, total_speed, tot_approach]
        out[sid] = np.asarray(obs)
    return out


def wave(signals):
    """states.py:116-127 -- 12 x sum(queue + approach)."""
    out = dict()
    for sid, sig in signals.items():
        st = []
        for d in sig.lane_sets:
            st.append(sum(sig.full_observation[l]['queue'] + sig.full_observation[l]['approach']
                          for l in sig.lane_sets[d]))
        out[sid] = np.asarray(st)
    return out


# ---- batched device views (N instances) ----------------------------------------------------------
def _b_mplight(env):
    return env.sim.obs_view()["mplight"]


def _b_wave(env):
    return env.sim.obs_view()["wave"]


def _b_drq(env, norm):
    """[N, n_sig_lanes, 5] rows in signal-major lane order (ragged per signal: see env.sig_lane_slices)."""
    import torch
    v = env.sim.obs_view()
    SL = env.sim.SL
    phase = v["phase"]                                       # [N, S]
    lane_sig = env.lane_sig_t                                # [SL] signal of each row
    lane_slot = env.lane_slot_t                              # [SL] row index inside the signal
    onehot = (lane_slot[None, :] == phase[:, lane_sig]).to(torch.float32)
    if norm:
        cols = [onehot, v["lane_approach"] / 28, v["lane_total_wait"] / 28, v["lane_queue"] / 28,
                v["lane_speed_sum"] / 20 / 28]
    else:
        cols = [onehot, v["lane_approach"], v["lane_total_wait"], v["lane_queue"], v["lane_speed_sum"]]
    return torch.stack(cols, dim=-1).view(-1, SL, 5)


mplight.batched = _b_mplight
wave.batched = _b_wave
drq.batched = lambda env: _b_drq(env, False)
drq_norm.batched = lambda env: _b_drq(env, True)
