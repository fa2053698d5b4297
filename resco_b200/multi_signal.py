"""``MultiSignal`` -- the reference's gym surface (multi_signal.py:9-234) on the vectorised backend.

Constructor signature, ``reset()`` / ``step(act)`` return shapes (dict keyed by signal id, or lists
in ``ts_order`` when ``gymma``), the attributes read by callers (``obs_shape``, ``phases``,
``all_ts_ids``, ``connection_name``, ``observation_space``, ``action_space``, ``n_agents``,
``ts_order``, ``signals``, ``metrics``) and ``calc_metrics`` / ``save_metrics`` follow the reference,
so ``agents/`` and the EPyMARL registration can construct it unmodified.

Extra keyword arguments (all optional) select the batched mode:
  n_env    number of lock-step instances (default 1 -> per-instance dict view, reference semantics);
           with n_env > 1 ``step`` takes ``[N, S]`` actions and returns device tensors
           (``state_fn.batched`` / ``reward_fn.batched``) and ``done`` as a bool.
  device   CUDA device index;   seed   base RNG seed of episode 1, +1 per reset (None -> the run index: episodes differ
           from each other like ``--random`` runs but repeat from one process to the next)
  vcap     vehicles an instance's store holds (0: from the map's size);  tile_vcap  vehicles of an instance the kernel
           keeps in shared memory (performance only, see RsScenario.tile_vcap)
  backend  factory ``Marshalled -> simulator`` (tests inject the CPU oracle); the default is the CUDA
           ``VecSim`` and raises if the extension or a GPU is missing -- there is no CPU fallback.
"""
from __future__ import annotations

import os
from typing import Callable, Dict, List, Optional

import numpy as np

from .abi import Marshalled, marshal
from .scenario.compiler import Scenario
from .traffic_signal import Signal

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")


class _Box:          # stand-ins so that the class works without gym installed (gym is optional here)
    def __init__(self, low, high, shape):
        self.low, self.high, self.shape = low, high, tuple(shape)


class _Discrete:
    def __init__(self, n):
        self.n = int(n)


try:                                  # pragma: no cover - depends on the environment
    import gym as _gym
    _EnvBase = _gym.Env
    _mk_box = lambda shape: _gym.spaces.Box(low=-np.inf, high=np.inf, shape=shape)   # noqa: E731
    _mk_discrete = lambda n: _gym.spaces.Discrete(n)                                  # noqa: E731
except Exception:                     # gym absent: plain object with the same attributes
    _EnvBase = object
    _mk_box = lambda shape: _Box(-np.inf, np.inf, shape)                              # noqa: E731
    _mk_discrete = lambda n: _Discrete(n)                                             # noqa: E731


def load_scenario(map_name: str) -> Scenario:
    path = os.path.join(_DATA, map_name + ".npz")
    if not os.path.exists(path):
        raise FileNotFoundError(f"no compiled scenario for map '{map_name}' ({path}); "
                                "run tools/compile_scenarios.py where the SUMO files are available")
    return Scenario.load(path)


def _default_backend(n_env: int, device: int):
    def make(m: Marshalled):
        from .sim import VecSim      # raises if the CUDA library is missing: no CPU fallback
        return VecSim(m, n_env, seed=0, device=device)
    return make


class MultiSignal(_EnvBase):
    def __init__(self, run_name, map_name, net, state_fn, reward_fn, route=None, gui=False, end_time=3600,
                 step_length=10, yellow_length=4, step_ratio=1, max_distance=200, lights=(), log_dir='/',
                 libsumo=False, warmup=0, gymma=False, *, n_env: int = 1, device: int = 0,
                 seed: Optional[int] = None, backend: Optional[Callable[[Marshalled], object]] = None,
                 scenario: Optional[Scenario] = None, vcap: int = 0, sigma: float = -1.0, speed_dev: float = -1.0,
                 tripinfo: Optional[bool] = None, tile_vcap: int = 0):
        if warmup != 0:
            raise NotImplementedError("warmup ticks before program installation are not supported (all shipped maps use 0)")
        if step_ratio != 1:
            raise NotImplementedError("step_ratio != 1 (sub-second SUMO steps) is not supported")
        self.libsumo = libsumo
        self.gymma = gymma
        self.log_dir = log_dir
        self.net, self.route, self.gui = net, route, gui
        self.state_fn, self.reward_fn = state_fn, reward_fn
        self.max_distance = max_distance
        self.warmup = warmup
        self.end_time = end_time
        self.step_length = step_length
        self.yellow_length = yellow_length
        self.step_ratio = step_ratio
        self.connection_name = run_name + '-' + map_name + '---' + state_fn.__name__ + '-' + reward_fn.__name__
        self.map_name = map_name
        self.n_env = int(n_env)
        self.seed = seed
        self.scenario = scenario if scenario is not None else load_scenario(map_name)
        sc = self.scenario
        # tripinfo_<run>.xml like `--tripinfo-output` (multi_signal.py:127-129): on by default for n_env == 1
        self.tripinfo = (self.n_env == 1 and log_dir is not None) if tripinfo is None else bool(tripinfo)
        self.marshalled = marshal(sc, step_length=step_length, yellow_length=yellow_length,
                                  max_distance=float(max_distance), end_time=float(end_time), vcap=vcap,
                                  sigma=sigma, speed_dev=speed_dev, record_trips=self.tripinfo, tile_vcap=tile_vcap)
        m = self.marshalled
        self.sim = (backend or _default_backend(self.n_env, device))(m)
        outs = tuple(getattr(state_fn, 'kernel_outputs', ())) + tuple(getattr(reward_fn, 'kernel_outputs', ()))
        if outs and hasattr(self.sim, 'select_outputs'):
            self.sim.select_outputs(*outs)
        self._begin = float(sc.meta["begin"])
        self._end_tick = m.struct.end_tick
        self._tick = 0

        sig_ids: List[str] = list(m.info["signal_ids"])
        if len(lights) > 0 and list(lights) != sig_ids:
            raise ValueError("lights differ from the compiled scenario's controlled signals")
        # green phases per signal (multi_signal.py:52-59): (duration, state) of phases w/o 'y' having g/G
        self.phases = {}
        for t in sc.meta["tls_ids"]:
            prog = sc.meta["programs"][t]
            self.phases[t] = [(d, s) for d, s in prog if 'y' not in s and 'g' in s.lower()]
        self.all_ts_ids = sig_ids
        self.ts_starter = len(self.all_ts_ids)
        self.signal_ids = list(sig_ids)
        self.signals: Dict[str, Signal] = dict()
        self.wait_metric = dict()
        self._make_signals()
        self._episode_seed = 0 if seed is None else int(seed)
        self.sim.reset(self._episode_seed, 0)
        self.sim.observe()
        self._observed_tick = 0
        self._refresh_views()
        self.obs_shape = dict()
        self.observation_space = list()
        self.action_space = list()
        observations = self.state_fn(self.signals)
        self.ts_order = list()
        for ts in observations:
            o_shape = observations[ts].shape
            self.obs_shape[ts] = o_shape
            self.ts_order.append(ts)
            self.observation_space.append(_mk_box(o_shape))
            if ts == 'top_mgr' or ts == 'bot_mgr':
                continue
            self.action_space.append(_mk_discrete(len(self.phases[ts])))
        self.n_agents = self.ts_starter
        self.run = 0
        self.demand_episode = 0
        self.metrics = []
        self.connection_name = (run_name + '-' + map_name + '-' + str(len(lights)) + '-' + state_fn.__name__ + '-'
                                + reward_fn.__name__)
        # row layout of the per-lane observation tensors ([N, n_sig_lanes], signal-major): used by calc_metrics and by
        # the .batched state / reward expressions
        a = sc.arrays
        self._lane_sig = np.concatenate([np.full(a["sig_lane_off"][s + 1] - a["sig_lane_off"][s], s, np.int64)
                                         for s in range(len(sig_ids))]) if sig_ids else np.zeros(0, np.int64)
        self._lane_slot = np.concatenate([np.arange(a["sig_lane_off"][s + 1] - a["sig_lane_off"][s])
                                          for s in range(len(sig_ids))]) if sig_ids else np.zeros(0, np.int64)
        self.sig_lane_slices = [slice(int(a["sig_lane_off"][s]), int(a["sig_lane_off"][s + 1]))
                                for s in range(len(sig_ids))]
        self._lane_sig_t = None

    # ------------------------------------------------------------------------------------------
    @property
    def lane_sig_t(self):
        if self._lane_sig_t is None:
            import torch
            dev = self.sim.obs_view()["phase"].device
            self._lane_sig_t = torch.as_tensor(self._lane_sig, device=dev)
            self._lane_slot_t = torch.as_tensor(self._lane_slot, device=dev).to(torch.int32)
        return self._lane_sig_t

    @property
    def lane_slot_t(self):
        self.lane_sig_t
        return self._lane_slot_t

    def _count_present(self):
        """[N, S] vehicles inside the detector range per signal at the last observe (len(all_vehicles),
        traffic_signal.py:214-224); a fresh tensor, the observation buffers are overwritten by the next step."""
        v = self.sim.obs_view()
        per_lane = v["lane_queue"] + v["lane_approach"]
        if getattr(self, "_lane_to_sig_t", None) is None:
            import torch
            m = torch.zeros((per_lane.shape[1], len(self.signal_ids)), device=per_lane.device)
            for s, sl in enumerate(self.sig_lane_slices):
                m[sl, s] = 1.0
            self._lane_to_sig_t = m
        return per_lane @ self._lane_to_sig_t

    def presence_counts(self):
        """(vehicles per signal now, at the previous observe or None before the first step) -- what the batched FMA2C
        manager reward needs to turn arrival counts into departures."""
        return self._n_now, self._n_prev

    def mdp_config(self, key):
        """mdp_configs[key][map] with the 'supervisors' reverse map (main.py:48-70)."""
        cfg = dict(self.scenario.meta.get('mdp', {}).get(key) or {})
        if not cfg:
            raise KeyError(f"no {key} configuration for map '{self.map_name}'")
        cfg['supervisors'] = {w: mgr for mgr, workers in cfg['management'].items() for w in workers}
        return cfg

    def _make_signals(self):
        for i, ts in enumerate(self.signal_ids):
            self.signals[ts] = Signal(self, ts, i)
            self.wait_metric[ts] = 0.0
        for ts in self.signal_ids:
            self.signals[ts].signals = self.signals

    def _phase_of(self, sig_index: int) -> int:
        return int(self._phases[sig_index])

    def _set_phase_one(self, sig_index: int, idx: int):
        S = len(self.signal_ids)
        ph = np.zeros((self.n_env, S), np.int32)
        mk = np.zeros((self.n_env, S), np.uint8)
        ph[:, sig_index] = idx
        mk[:, sig_index] = 1
        self.sim.set_phase(ph, mk)
        self._phases = self.sim.phases(0)[self.scenario.arrays["sig_tls"]]

    def _observe_into_signals(self):
        """Signal.observe() of a caller that drives the signals by hand (N = 1): the reference loops
        ``for ts in signal_ids: signals[ts].observe(...)`` (multi_signal.py:185-186) and every call latches only ITS
        signal's waiting times.  The device sweep covers all signals at once, so it runs once per tick: the first
        Signal.observe() after a tick does the sweep, the calls for the other signals in the same tick find it done."""
        if getattr(self, '_observed_tick', None) != self._tick:
            self.sim.observe()
            self._observed_tick = self._tick
            self._refresh_views()

    def _refresh_views(self, env: int = 0):
        """Per-instance dict view (Signal.full_observation) of instance `env` from the device buffers."""
        sc = self.scenario
        a = sc.arrays
        ob = self.sim.obs()
        veh = self.sim.vehicles(env)
        self._phases = self.sim.phases(env)[a["sig_tls"]]
        self._last_obs = ob
        lane_of = veh["lane"]
        trip_ids = sc.meta.get("trip_ids")
        vt_ids = sc.meta["vtype_ids"]
        md = np.float32(self.max_distance)
        for s, ts in enumerate(self.signal_ids):
            sig = self.signals[ts]
            full = dict()
            allv = set()
            q0 = int(a["sig_lane_off"][s])
            for slot, lane in enumerate(sig.lanes):
                q = q0 + slot
                li = int(a["sig_lane"][q])
                vehicles = []
                td = np.float32(a["lane_tls_dist"][li])
                if td >= 0:
                    for i in np.nonzero(lane_of == li)[0]:
                        dist = np.float32(np.float32(a["lane_len"][li]) - veh["pos"][i]) + td
                        if not (dist <= md):
                            continue
                        vid = int(veh["vid"][i])
                        name = trip_ids[vid] if trip_ids is not None and vid < len(trip_ids) else f"veh{vid}"
                        allv.add(name)
                        vehicles.append({'id': name, 'wait': float(veh["rwait"][i]), 'speed': float(veh["speed"][i]),
                                         'acceleration': float(veh["accel"][i]), 'position': float(veh["pos"][i]),
                                         'type': vt_ids[int(veh["vtype"][i])]})
                full[lane] = {'queue': int(ob["lane_queue"][env, q]), 'approach': int(ob["lane_approach"][env, q]),
                              'total_wait': float(ob["lane_total_wait"][env, q]),
                              'max_wait': float(ob["lane_max_wait"][env, q]), 'vehicles': vehicles}
            full['num_vehicles'] = allv
            if sig.last_step_vehicles is None:
                full['arrivals'] = allv
                full['departures'] = set()
            else:
                full['arrivals'] = allv.difference(sig.last_step_vehicles)
                full['departures'] = sig.last_step_vehicles.difference(allv)
            sig.last_step_vehicles = allv
            sig.full_observation = full
            sig.waiting_times = {v['id']: v['wait'] for lane in sig.lanes for v in full[lane]['vehicles'] if v['wait'] > 0}

    # ------------------------------------------------------------------------------------------
    def step_sim(self):
        for _ in range(self.step_ratio):
            self.sim.tick(1)
            self._tick += 1

    def save_tripinfo(self):
        """tripinfo_<run>.xml of instance 0 (readable by utils/readXML.py)."""
        from .metrics import write_tripinfo
        os.makedirs(self._log_path(), exist_ok=True)
        path = os.path.join(self._log_path(), 'tripinfo_' + str(self.run) + '.xml')
        write_tripinfo(path, self.scenario, self.sim.trip_records(0), self.sim.vehicles(0), self._tick)
        return path

    def reset(self):
        if self.run != 0:
            self.save_metrics()
            if self.tripinfo:
                self.save_tripinfo()
        self.metrics = []
        self.run += 1
        self._episode_seed = (self.run if self.seed is None else int(self.seed) + self.run - 1)
        # grid4x4 / arterial4x4: a different route file every episode (multi_signal.py:124); the compiled scenario
        # holds the first `n_demand_episodes` of them, run r uses file ((r - 1) mod n) + 1
        bank = self.scenario.arrays.get("bank_origin_off")
        self.demand_episode = 0
        if bank is not None and hasattr(self.sim, "set_demand_window"):
            self.demand_episode = (self.run - 1) % bank.shape[0]
            self.sim.set_demand_window(bank[self.demand_episode])
        self.sim.reset(self._episode_seed, 0)
        self._tick = 0
        self.signal_ids = list(self.all_ts_ids)
        self._make_signals()
        self.sim.observe()
        self._observed_tick = 0
        if self.n_env > 1:
            self._n_prev = None
            self._n_now = self._count_present()
            return self.state_fn.batched(self)
        self._refresh_views()
        states = self.state_fn(self.signals)
        if self.gymma:
            return [states[ts] for ts in self.ts_order]
        return states

    def step(self, act):
        S = len(self.signal_ids)
        if self.n_env > 1:
            self.sim.env_step(act)
            self._tick += self.step_length
            self._n_prev, self._n_now = self._n_now, self._count_present()
            done = self._begin + self._tick >= self.end_time
            return self.state_fn.batched(self), self.reward_fn.batched(self), done, {'eps': self.run}
        if self.gymma:
            act = {ts: act[i] for i, ts in enumerate(self.ts_order)}
        a = np.zeros((1, S), np.int32)
        for i, ts in enumerate(self.signal_ids):
            a[0, i] = int(act[ts])
        self.sim.env_step(a)                  # prep_phase -> yellow ticks -> set_phase -> green ticks -> observe
        self._tick += self.step_length
        self._observed_tick = self._tick
        self._refresh_views()
        observations = self.state_fn(self.signals)
        rewards = self.reward_fn(self.signals)
        self.calc_metrics(rewards)
        done = self._begin + self._tick >= self.end_time
        if self.gymma:
            return ([observations[ts] for ts in self.ts_order], [rewards[ts] for ts in self.ts_order], [done],
                    {'eps': self.run})
        return observations, rewards, done, {'eps': self.run}

    def calc_metrics(self, rewards):
        """multi_signal.py:199-216."""
        queue_lengths, max_queues = dict(), dict()
        for s, ts in enumerate(self.signal_ids):
            queue_lengths[ts] = int(self._last_obs["sig_queue_len"][0, s])
            max_queues[ts] = int(self._last_obs["sig_max_queue"][0, s])
        self.metrics.append({'step': self._begin + self._tick, 'reward': rewards, 'max_queues': max_queues,
                             'queue_lengths': queue_lengths})

    def _log_path(self):
        base = self.log_dir if self.log_dir not in (None, '/') else os.path.join('/tmp', 'resco_b200_logs')
        return os.path.join(base, self.connection_name)

    def save_metrics(self):
        """multi_signal.py:218-226 (same CSV line format, readable by utils/readCSV.py)."""
        os.makedirs(self._log_path(), exist_ok=True)
        log = os.path.join(self._log_path(), 'metrics_' + str(self.run) + '.csv')
        with open(log, 'w+') as f:
            for line in self.metrics:
                f.write(''.join(str(line[k]) + ', ' for k in ['step', 'reward', 'max_queues', 'queue_lengths']) + '\n')

    def episode_stats(self):
        """Per-instance episode average of timeLoss + departDelay the way the reference computes it from the tripinfo
        file (utils/readXML.py:27-77): over the trips that were INSERTED (finished or still running -- SUMO writes
        both with --tripinfo-output.write-unfinished).  Trips that never got into the network are not in a tripinfo
        file; readXML charges them ``end_time - depart`` only when the route file holds <vehicle> elements
        (grid4x4 / arterial4x4), never for <trip> demand (cologne*, ingolstadt*).  Here the backlog of a
        <vehicle>-demand map is charged ``now - depart`` (== the reference's rule at the end of the episode when
        insertion is first-in-first-out; ``metrics.avg_delay_from_tripinfo`` applies the reference's rule literally to a
        written tripinfo file).  Also returns the raw counters."""
        st = self.sim.stats()
        n = (st["n_arrived"] + st["n_active"]).astype(np.float64)
        total = st["sum_delay_arrived"].astype(np.float64) + st["sum_delay_running"]
        if self.map_name in ("grid4x4", "arterial4x4"):
            n = n + st["n_backlog"]
            total = total + st["sum_delay_pending"]
        return dict(avg_delay=total / np.maximum(n, 1), trips=n, stats=st)

    def render(self, mode='human'):
        pass

    def close(self):
        if self.run != 0:
            self.save_metrics()
            if self.tripinfo:
                self.save_tripinfo()
        if hasattr(self.sim, "close"):
            self.sim.close()
