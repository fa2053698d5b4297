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


def _mdp(signals, key):
    """mdp_configs[key] of the map (config/mdp_config.py), with the worker->manager reverse map that
    main.py:62-70 precomputes under 'supervisors'."""
    env = next(iter(signals.values()))._env
    return env.mdp_config(key)


def _region_fringes(signals, cfg):
    """Inbound lanes through which traffic enters a manager's region (states.py:168-180)."""
    supervisors = cfg['supervisors']
    fringes = {mgr: [] for mgr in cfg['management']}
    for sid, sig in signals.items():
        for key, neighbor in sig.downstream.items():
            if neighbor is None or supervisors[neighbor] != supervisors[sid]:
                inbounds = sig.inbounds_fr_direction.get(key)
                if inbounds is not None:
                    fringes[supervisors[sid]] += inbounds
    return fringes


def _fma2c(signals, key, full):
    cfg = _mdp(signals, key)
    supervisors, neighbors_of = cfg['supervisors'], cfg['management_neighbors']
    fringes = _region_fringes(signals, cfg)
    lane_wave = {lane: sig.full_observation[lane]['queue'] + sig.full_observation[lane]['approach']
                 for sig in signals.values() for lane in sig.lanes}
    manager_obs = {mgr: np.clip(np.asarray([lane_wave[l] for l in lanes]) / cfg['norm_wave'], 0, cfg['clip_wave'])
                   for mgr, lanes in fringes.items()}
    managers = {mgr: np.concatenate([manager_obs[mgr]] + [cfg['alpha'] * manager_obs[nb] for nb in neighbors_of[mgr]])
                for mgr in manager_obs}
    signal_wave = dict()
    for sid, sig in signals.items():
        waves = []
        for lane in sig.lanes:
            waves.append(lane_wave[lane])
            if full:
                waves.append(sig.full_observation[lane]['total_wait'] / 28)
                waves.append(_speed_sum_norm(sig, lane))
        signal_wave[sid] = np.clip(np.asarray(waves) / cfg['norm_wave'], 0, cfg['clip_wave'])
    out = dict()
    for sid, sig in signals.items():
        waves = [signal_wave[sid]]
        for neighbor in sig.downstream.values():
            if neighbor is not None and supervisors[neighbor] == supervisors[sid]:
                waves.append(cfg['alpha'] * signal_wave[neighbor])
        waits = np.clip(np.asarray([sig.full_observation[l]['max_wait'] for l in sig.lanes]) / cfg['norm_wait'],
                        0, cfg['clip_wait'])
        out[sid] = np.concatenate([np.concatenate(waves), waits])
    out.update(managers)
    return out


def fma2c(signals):
    """states.py:162-229 -- worker obs (own + same-region neighbour waves, max waits) + manager obs."""
    return _fma2c(signals, 'FMA2C', full=False)


def fma2c_full(signals):
    """states.py:232-305."""
    return _fma2c(signals, 'FMA2CFull', full=True)


# ---- batched device views (N instances) ----------------------------------------------------------
def _b_mplight(env):
    return env.sim.obs_view()["mplight"]


def _b_wave(env):
    return env.sim.obs_view()["wave"]


def _b_drq(env, norm):
    """[N, n_sig_lanes, 5] rows in signal-major lane order (ragged per signal: see env.sig_lane_slices), written by the
    observe step of the kernel (RS_OUT_DRQ / RS_OUT_DRQ_NORM, switched on by MultiSignal for these state functions)."""
    name = "drq_norm" if norm else "drq"
    v = env.sim.obs_view()
    if name not in v:
        raise RuntimeError(f"states.{name}.batched: the '{name}' output is not selected (VecSim.select_outputs)")
    return v[name]


def _lane_rows(env):
    """lane name -> row of the per-lane observation tensors ([N, n_sig_lanes], signal-major); a lane listed by two
    signals resolves to the later one, like the reference's ``lane_wave`` dict (states.py:182-186)."""
    rows = dict()
    for s, sid in enumerate(env.signal_ids):
        q0 = env.sig_lane_slices[s].start
        for slot, lane in enumerate(env.signals[sid].lanes):
            rows[lane] = q0 + slot
    return rows


def fma2c_plan(env, key, full):
    """Static gather plan of states.fma2c / fma2c_full: every output element is
    ``coef * clip(F[src] / div, 0, clipmax)`` with F = [wave | total_wait/28 | speed_sum/20/28 | max_wait] over the
    per-lane rows.  Built once per env by walking the same loops as ``_fma2c`` (states.py:162-305)."""
    cfg = env.mdp_config(key)
    signals = env.signals
    supervisors, neighbors_of = cfg['supervisors'], cfg['management_neighbors']
    fringes = _region_fringes(signals, cfg)
    rows = _lane_rows(env)
    SL = env.sim.SL
    WAVE, TW, SPD, MW = 0, SL, 2 * SL, 3 * SL
    nw, cw = float(cfg['norm_wave']), float(cfg['clip_wave'])
    manager_obs = {mgr: [(WAVE + rows[l], nw, cw, 1.0) for l in lanes] for mgr, lanes in fringes.items()}
    plan = dict()
    signal_wave = dict()
    for sid, sig in signals.items():
        items = []
        for lane in sig.lanes:
            q = env.sig_lane_slices[env.signal_ids.index(sid)].start + sig.lanes.index(lane)
            items.append((WAVE + rows[lane], nw, cw, 1.0))
            if full:            # total_wait and the speed sum are read from the signal's OWN row of the lane
                items.append((TW + q, nw, cw, 1.0))
                items.append((SPD + q, nw, cw, 1.0))
        signal_wave[sid] = items
    alpha = float(cfg['alpha'])
    for sid, sig in signals.items():
        items = list(signal_wave[sid])
        for neighbor in sig.downstream.values():
            if neighbor is not None and supervisors[neighbor] == supervisors[sid]:
                items += [(a, b, c, alpha * d) for a, b, c, d in signal_wave[neighbor]]
        q0 = env.sig_lane_slices[env.signal_ids.index(sid)].start
        items += [(MW + q0 + slot, float(cfg['norm_wait']), float(cfg['clip_wait']), 1.0) for slot in range(len(sig.lanes))]
        plan[sid] = items
    for mgr in manager_obs:
        items = list(manager_obs[mgr])
        for nb in neighbors_of[mgr]:
            items += [(a, b, c, alpha * d) for a, b, c, d in manager_obs[nb]]
        plan[mgr] = items
    return plan


def _b_fma2c(env, key, full):
    """dict id -> [N, obs_dim] device tensor (workers, then the manager pseudo-agents), float32."""
    import torch
    cache = env.__dict__.setdefault('_fma2c_state_plans', dict())
    v = env.sim.obs_view()
    dev = v["lane_queue"].device
    if key not in cache:
        plan = fma2c_plan(env, key, full)
        cache[key] = {k: tuple(torch.tensor([it[j] for it in items], device=dev,
                                            dtype=torch.int64 if j == 0 else torch.float32) for j in range(4))
                      for k, items in plan.items()}
    F = torch.cat([v["lane_queue"] + v["lane_approach"], v["lane_total_wait"] / 28, v["lane_speed_sum"] / 20 / 28,
                   v["lane_max_wait"]], dim=1)
    out = dict()
    for k, (src, div, clipmax, coef) in cache[key].items():
        out[k] = torch.minimum(torch.clamp_min(F[:, src] / div, 0.0), clipmax) * coef
    return out


def _b_mplight_full(env):
    """[N, S, 1 + 48] device tensor written by the kernel (RS_OUT_MPLIGHT_FULL): phase, then per movement (pressure,
    sum(total_wait / 28), speed sum of the LAST lane of the movement -- the reference resets total_speed inside its
    lane loop, states.py:97 --, sum(approach / 28))."""
    v = env.sim.obs_view()
    if "mplight_full" not in v:
        raise RuntimeError("states.mplight_full.batched: the 'mplight_full' output is not selected (VecSim.select_outputs)")
    return v["mplight_full"]


mplight.batched = _b_mplight
mplight_full.batched = _b_mplight_full
wave.batched = _b_wave
fma2c.batched = lambda env: _b_fma2c(env, 'FMA2C', False)
fma2c_full.batched = lambda env: _b_fma2c(env, 'FMA2CFull', True)
drq.batched = lambda env: _b_drq(env, False)
drq_norm.batched = lambda env: _b_drq(env, True)
# optional kernel outputs a state function needs (MultiSignal switches them on: VecSim.select_outputs)
drq.kernel_outputs = ("drq",)
drq_norm.kernel_outputs = ("drq_norm",)
mplight_full.kernel_outputs = ("mplight_full",)
