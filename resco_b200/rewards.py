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


def _ma2c_cfg(signals):
    """mdp_configs['MA2C'] (rewards.py:51,66).  The reference ships no 'MA2C' entry in config/mdp_config.py, so its
    queue_maxwait* raise KeyError unless the caller has put one there (main.py:48-53 does for the agent it runs); same
    here: the entry is looked up in the scenario's mdp table (``env.scenario.meta['mdp']['MA2C']``: coef, coop_gamma)."""
    env = next(iter(signals.values()))._env
    return env.scenario.meta.get('mdp', {})['MA2C']


def queue_maxwait(signals):
    """rewards.py:44-53 -- -(sum over lanes of queue + coef * max_wait)."""
    coef = _ma2c_cfg(signals)['coef']
    out = dict()
    for sid, sig in signals.items():
        r = 0
        for lane in sig.lanes:
            r += sig.full_observation[lane]['queue']
            r += sig.full_observation[lane]['max_wait'] * coef
        out[sid] = -r
    return out


def queue_maxwait_neighborhood(signals):
    """rewards.py:56-69 -- own queue_maxwait + coop_gamma x that of every downstream neighbour."""
    own = queue_maxwait(signals)
    gamma = _ma2c_cfg(signals)['coop_gamma']
    out = dict()
    for sid, sig in signals.items():
        tot = own[sid]
        for neighbor in sig.downstream.values():
            if neighbor is not None:
                tot += gamma * own[neighbor]
        out[sid] = tot
    return out


def _b_queue_maxwait(env, neighborhood):
    """[N, S] device tensor from the kernel's per-lane queue / max_wait rows."""
    import torch
    cfg = env.scenario.meta.get('mdp', {})['MA2C']
    v = env.sim.obs_view()
    dev = v["lane_queue"].device
    cache = env.__dict__.setdefault('_ma2c_plan', None)
    if cache is None:
        S, SL = len(env.signal_ids), env.sim.SL
        lane_to_sig = torch.zeros((SL, S), device=dev)
        nb = torch.zeros((S, S), device=dev)
        for si, sid in enumerate(env.signal_ids):
            lane_to_sig[env.sig_lane_slices[si], si] = 1.0
            nb[si, si] += 1.0
            for neighbor in env.signals[sid].downstream.values():
                if neighbor is not None:
                    nb[env.signal_ids.index(neighbor), si] += float(cfg['coop_gamma'])
        cache = env.__dict__['_ma2c_plan'] = (lane_to_sig, nb)
    lane_to_sig, nb = cache
    own = -((v["lane_queue"] + v["lane_max_wait"] * float(cfg['coef'])) @ lane_to_sig)
    return own @ nb if neighborhood else own


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


def _b_fma2c(env, key):
    """dict id -> [N] device tensor.  Workers: -(queue + coef * max_wait) summed over the signal's lanes plus alpha x
    the same of same-region neighbours; managers: fringe arrivals + liquidity (departures - arrivals) plus alpha x
    the neighbouring regions' (rewards.py:72-136).  len(arrivals) comes from the kernel's per-lane arrival counts
    (``lane_arrivals``), len(departures) = vehicles at the previous observe - (vehicles now - arrivals); the previous
    count is kept by MultiSignal (``presence_counts``)."""
    import torch
    from .states import _region_fringes
    v = env.sim.obs_view()
    dev = v["lane_queue"].device
    cache = env.__dict__.setdefault('_fma2c_reward_plans', dict())
    if key not in cache:
        cfg = env.mdp_config(key)
        signals = env.signals
        supervisors, neighbors_of = cfg['supervisors'], cfg['management_neighbors']
        fringes = _region_fringes(signals, cfg)
        S, SL = len(env.signal_ids), env.sim.SL
        managers = list(cfg['management'])
        lane_to_sig = torch.zeros((SL, S), device=dev)            # per-lane -> per-signal sums
        fringe_w = torch.zeros((SL, len(managers)), device=dev)   # lanes counted in a manager's fringe arrivals
        region = torch.zeros((S, len(managers)), device=dev)      # signal -> its manager
        for si, sid in enumerate(env.signal_ids):
            sl = env.sig_lane_slices[si]
            lane_to_sig[sl, si] = 1.0
            mi = managers.index(supervisors[sid])
            region[si, mi] = 1.0
            for slot, lane in enumerate(signals[sid].lanes):
                if lane in fringes[supervisors[sid]]:
                    fringe_w[sl.start + slot, mi] = 1.0
        nb_sig = torch.zeros((S, S), device=dev)                   # own + alpha * same-region downstream neighbours
        for si, sid in enumerate(env.signal_ids):
            nb_sig[si, si] += 1.0
            for neighbor in signals[sid].downstream.values():
                if neighbor is not None and supervisors[neighbor] == supervisors[sid]:
                    nb_sig[env.signal_ids.index(neighbor), si] += float(cfg['alpha'])
        nb_mgr = torch.zeros((len(managers), len(managers)), device=dev)
        for mi, mgr in enumerate(managers):
            nb_mgr[mi, mi] += 1.0
            for nb in neighbors_of[mgr]:
                nb_mgr[managers.index(nb), mi] += float(cfg['alpha'])
        cache[key] = (float(cfg['coef']), managers, lane_to_sig, fringe_w, region, nb_sig, nb_mgr)
    coef, managers, lane_to_sig, fringe_w, region, nb_sig, nb_mgr = cache[key]
    own = -((v["lane_queue"] + v["lane_max_wait"] * coef) @ lane_to_sig)          # [N, S]
    worker = own @ nb_sig
    n_now, n_prev = env.presence_counts()
    arrivals = v["lane_arrivals"] @ lane_to_sig
    departures = (n_prev - (n_now - arrivals)) if n_prev is not None else torch.zeros_like(arrivals)
    mgr_own = v["lane_arrivals"] @ fringe_w + (departures - arrivals) @ region   # [N, n_managers]
    mgr = mgr_own @ nb_mgr
    out = {sid: worker[:, si] for si, sid in enumerate(env.signal_ids)}
    out.update({m: mgr[:, mi] for mi, m in enumerate(managers)})
    return out


fma2c.batched = lambda env: _b_fma2c(env, 'FMA2C')
fma2c_full.batched = lambda env: _b_fma2c(env, 'FMA2CFull')
queue_maxwait.batched = lambda env: _b_queue_maxwait(env, False)
queue_maxwait_neighborhood.batched = lambda env: _b_queue_maxwait(env, True)
wait.batched = lambda env: env.sim.obs_view()["reward_wait"]
wait_norm.batched = lambda env: env.sim.obs_view()["reward_wait_norm"]
pressure.batched = lambda env: env.sim.obs_view()["reward_pressure"]
