#!/usr/bin/env python
"""Generate golden vectors by running the REAL reference Python (dev container only).

The reference package is imported from /root/reference *unmodified* (``multi_signal.py``,
``traffic_signal.py``, ``states.py``, ``rewards.py``, ``agents/maxwave.py``, ``agents/maxpressure.py``)
with three stub modules in ``sys.modules`` -- ``traci`` (our TraCI-subset facade), ``sumolib``
(``checkBinary``) and ``gym`` (``Env``, ``spaces``) -- because SUMO / gym are not installed anywhere in
this environment.  The simulator behind the facade is the CPU oracle, so what these vectors pin is
everything the reference itself implements on this path:

  create_yellows (traffic_signal.py:7-24), green-phase discovery (multi_signal.py:52-59), Signal
  topology / lane order (traffic_signal.py:46-87), the phase machine prep_phase / set_phase
  (:176-187) incl. the static-program countdown between calls, the env-step schedule
  (multi_signal.py:164-197), Signal.observe with the waiting-time latch and detector range
  (traffic_signal.py:189-247), states.{mplight,mplight_full,wave,drq,drq_norm}, rewards.{wait,
  wait_norm,pressure}, calc_metrics (multi_signal.py:199-216) and the MAXPRESSURE / MAXWAVE agents.

The tests then drive OUR MultiSignal (fused env-step path of the oracle on CPU, of the CUDA kernel on
the GPU box) with the recorded actions and require identical observations / rewards / phases.

Writes tests/golden/<case>.npz.  Usage: python tools/make_golden.py
"""
import json
import os
import sys
import types

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'oracle'))
REF_PARENT = '/root/reference'
REF = os.path.join(REF_PARENT, 'resco_benchmark')

from resco_b200.multi_signal import load_scenario          # noqa: E402
from resco_b200.traci_facade import Phase, open_facade      # noqa: E402
from pyoracle import OracleSim                              # noqa: E402

CFG = dict(step_length=10, yellow_length=3, max_distance=200.0, next_seed=0, sigma=-1.0, speed_dev=-1.0)


def install_stubs():
    traci = types.ModuleType('traci')
    conns = {}

    def _map_of(cmd):
        for flag in ('-c', '-n'):
            if flag in cmd:
                return os.path.basename(os.path.dirname(cmd[cmd.index(flag) + 1]))
        raise ValueError(cmd)

    def start(cmd, label='default'):
        sc = load_scenario(_map_of(cmd))
        seed = CFG['next_seed']
        CFG['next_seed'] += 1
        conns[label] = open_facade(sc, lambda m: OracleSim(m, 1, seed=seed), step_length=CFG['step_length'],
                                   yellow_length=CFG['yellow_length'], max_distance=CFG['max_distance'], seed=seed,
                                   sigma=CFG['sigma'], speed_dev=CFG['speed_dev'])
        traci._current = label

    traci.start = start
    traci.getConnection = lambda label: conns[label]
    traci.switch = lambda label: setattr(traci, '_current', label)
    traci.close = lambda: conns.pop(traci._current).close() if traci._current in conns else None
    tl = types.ModuleType('traci.trafficlight')
    tl.Phase = Phase
    traci.trafficlight = tl
    sys.modules['traci'] = traci
    sys.modules['traci.trafficlight'] = tl
    sumolib = types.ModuleType('sumolib')
    sumolib.checkBinary = lambda name: name
    sys.modules['sumolib'] = sumolib
    gym = types.ModuleType('gym')

    class Env:
        pass

    class Box:
        def __init__(self, low, high, shape):
            self.low, self.high, self.shape = low, high, shape

    class Discrete:
        def __init__(self, n):
            self.n = n
    gym.Env = Env
    gym.spaces = types.SimpleNamespace(Box=Box, Discrete=Discrete)
    sys.modules['gym'] = gym
    # package shim: submodules resolve from the reference tree, __init__.py (sys.exit without SUMO_HOME) is skipped
    pkg = types.ModuleType('resco_benchmark')
    pkg.__path__ = [REF]
    sys.modules['resco_benchmark'] = pkg


def ref_env(map_name, state_name, reward_name, max_distance, yellow_length=None):
    import importlib
    ms = importlib.import_module('resco_benchmark.multi_signal')
    states = importlib.import_module('resco_benchmark.states')
    rewards = importlib.import_module('resco_benchmark.rewards')
    mc = importlib.import_module('resco_benchmark.config.map_config').map_configs[map_name]
    yl = mc['yellow_length'] if yellow_length is None else yellow_length
    CFG.update(step_length=mc['step_length'], yellow_length=yl, max_distance=float(max_distance), next_seed=0)
    if state_name.startswith('fma2c'):          # main.py:48-70: per-map mdp config + 'supervisors' reverse map
        mdp = importlib.import_module('resco_benchmark.config.mdp_config').mdp_configs
        key = 'FMA2CFull' if state_name == 'fma2c_full' else 'FMA2C'
        cfgm = dict(mdp[key][map_name]) if map_name in mdp[key] else dict(mdp[key])
        cfgm['supervisors'] = {w: mgr for mgr, ws in cfgm['management'].items() for w in ws}
        mdp[key] = cfgm
    net = os.path.join(REF, mc['net'])
    route = os.path.join(REF, mc['route']) if mc['route'] is not None else None
    if route is not None:
        raise NotImplementedError('zip-routed maps are generated through net=sumocfg-less path below')
    env = ms.MultiSignal('golden', map_name, net, getattr(states, state_name), getattr(rewards, reward_name),
                         route=route, step_length=mc['step_length'], yellow_length=yl, step_ratio=mc['step_ratio'],
                         end_time=mc['end_time'], max_distance=max_distance, lights=mc['lights'], log_dir='/tmp/golden_logs/',
                         libsumo=False, warmup=mc['warmup'])
    return env, mc


def record(map_name, state_name, reward_name, max_distance, n_steps, policy, seed=0, yellow_length=None):
    import importlib
    env, mc = ref_env(map_name, state_name, reward_name, max_distance, yellow_length)
    rng = np.random.default_rng(seed)
    obs = env.reset()
    order = list(env.ts_order)
    n_act = {ts: len(env.phases[ts]) for ts in order if ts in env.phases}
    agent = None
    if policy in ('MAXPRESSURE', 'MAXWAVE'):
        agent_mod = importlib.import_module('resco_benchmark.agents.' + ('maxpressure' if policy == 'MAXPRESSURE' else 'maxwave'))
        cls = getattr(agent_mod, policy)
        agent = cls({}, {ts: [env.obs_shape[ts], n_act[ts]] for ts in order}, map_name, 0)
    rec = dict(obs=[], rew=[], act=[], phase=[], queue_lengths=[], max_queues=[], step_time=[])
    rec_reset = [np.asarray(obs[ts], np.float64).ravel() for ts in order]
    hold = {ts: 0 for ts in order}
    cur = {ts: 0 for ts in order}
    for step in range(n_steps):
        if agent is not None:
            act = agent.act(obs)
        else:
            act = {}
            for ts in [t for t in order if t in n_act]:   # random, with some holding (both prep_phase branches)
                if hold[ts] <= 0:
                    cur[ts] = int(rng.integers(n_act[ts]))
                    hold[ts] = int(rng.integers(1, 4))
                hold[ts] -= 1
                act[ts] = cur[ts]
        obs, rew, done, info = env.step(act)
        rec['act'].append([int(act.get(ts, -1)) for ts in order])
        rec['obs'].append(np.concatenate([np.asarray(obs[ts], np.float64).ravel() for ts in order]))
        rec['rew'].append([float(rew[ts]) for ts in order])
        rec['phase'].append([int(env.signals[ts].phase) if ts in env.signals else -1 for ts in order])
        mt = env.metrics[-1]
        rec['queue_lengths'].append([mt['queue_lengths'].get(ts, -1) for ts in order])
        rec['max_queues'].append([mt['max_queues'].get(ts, -1) for ts in order])
        rec['step_time'].append(mt['step'])
    real = [ts for ts in order if ts in env.signals]
    yellows = {ts: env.signals[ts].yellow_dict for ts in real}
    programs = {ts: [[p.duration, p.state] for p in env.signals[ts].phases] for ts in real}
    meta = dict(map=map_name, state=state_name, reward=reward_name, max_distance=max_distance, policy=policy,
                ts_order=order, obs_shapes={ts: list(env.obs_shape[ts]) for ts in order}, yellow_dicts=yellows,
                programs=programs, lanes={ts: env.signals[ts].lanes for ts in real},
                step_length=mc['step_length'], yellow_length=CFG['yellow_length'], episode_seed=1,
                n_actions=n_act)
    env.close()
    out = os.path.join(ROOT, 'tests', 'golden', f"{map_name}_{state_name}_{reward_name}_{policy.lower()}.npz")
    np.savez_compressed(out, meta=np.frombuffer(json.dumps(meta).encode(), np.uint8),
                        reset_obs=np.concatenate(rec_reset), obs=np.asarray(rec['obs']), rew=np.asarray(rec['rew']),
                        act=np.asarray(rec['act'], np.int32), phase=np.asarray(rec['phase'], np.int32),
                        queue_lengths=np.asarray(rec['queue_lengths'], np.int32),
                        max_queues=np.asarray(rec['max_queues'], np.int32), step_time=np.asarray(rec['step_time']))
    print('wrote', os.path.relpath(out, ROOT), 'obs', np.asarray(rec['obs']).shape, f'{os.path.getsize(out) / 1024:.0f} KiB')


def main():
    install_stubs()
    os.makedirs(os.path.join(ROOT, 'tests', 'golden'), exist_ok=True)
    record('cologne1', 'mplight', 'wait', 200, 150, 'random')
    record('cologne1', 'drq_norm', 'wait_norm', 200, 100, 'random', seed=1)
    record('cologne1', 'mplight', 'wait', 200, 120, 'MAXPRESSURE')
    record('cologne8', 'mplight', 'pressure', 200, 150, 'random', seed=2)
    record('cologne8', 'mplight', 'wait', 200, 360, 'MAXPRESSURE')
    record('cologne8', 'wave', 'wait', 50, 120, 'MAXWAVE')
    record('cologne8', 'drq_norm', 'wait_norm', 200, 80, 'random', seed=3, yellow_length=4)   # EPyMARL registration values
    record('cologne8', 'mplight_full', 'pressure', 200, 60, 'random', seed=4)
    record('cologne3', 'drq', 'wait', 200, 60, 'random', seed=5)
    record('ingolstadt21', 'mplight', 'pressure', 200, 60, 'random', seed=6)
    record('ingolstadt21', 'drq_norm', 'wait_norm', 200, 40, 'random', seed=7)
    record('cologne8', 'fma2c', 'fma2c', 200, 80, 'random', seed=8)
    record('ingolstadt7', 'fma2c_full', 'fma2c_full', 200, 50, 'random', seed=9)


if __name__ == '__main__':
    main()
