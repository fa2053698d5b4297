"""The N = 1 seam: the TraCI-subset facade (resco_b200/traci_facade.py) driven the way the reference's classes drive
`self.sumo`, and our Signal objects driven by hand.

`RefStyleEnv` below restates, in a few lines, the call sequence of the reference's MultiSignal.step / Signal.observe
over the TraCI surface (multi_signal.py:164-197, traffic_signal.py:176-247): setPhase for the yellow, simulationStep x
yellow_length, setPhase, simulationStep x rest, then per signal and lane getLastStepVehicleIDs / getNextTLS /
getWaitingTime with the waiting-time latch.  (The unmodified reference classes themselves run over this facade in
tools/make_golden.py, in the dev container; /root/reference does not exist on the GPU box.)
"""
import numpy as np
import pytest

import resco_b200.rewards as rewards
import resco_b200.states as states
import util
from resco_b200.multi_signal import MultiSignal
from resco_b200.traci_facade import Phase, open_facade


class RefStyleEnv:
    def __init__(self, map_name, backend_factory, seed):
        self.sc = util.load(map_name)
        mc = self.sc.meta["map_config"]
        self.step_length, self.yellow_length, self.max_distance = mc["step_length"], mc["yellow_length"], 200.0
        self.sumo = open_facade(self.sc, backend_factory, step_length=self.step_length, yellow_length=self.yellow_length,
                                max_distance=self.max_distance, seed=seed)
        m = self.sumo.m
        self.ids = list(m.info["signal_ids"])
        self.lanes = {s: list(self.sc.meta["signals"][s]["lanes"]) for s in self.ids}
        self.yellow = {s: dict(m.info["yellow_dicts"][s]) for s in self.ids}
        self.next_phase = {s: 0 for s in self.ids}
        self.waiting = {s: dict() for s in self.ids}
        self.last = {s: None for s in self.ids}
        for s in self.ids:      # Signal.__init__ installs greens + yellows (traffic_signal.py:93-100)
            logic = self.sumo.trafficlight.getAllProgramLogics(s)[0]
            logic.type = 0
            logic.phases = [Phase(d, st) for d, st in m.info["programs_installed"][s]]
            self.sumo.trafficlight.setProgramLogic(s, logic)

    def observe(self):
        out = {}
        for s in self.ids:
            full, allv = {}, set()
            for lane in self.lanes[s]:
                q = a = 0
                tw = mw = 0.0
                for v in self.sumo.lane.getLastStepVehicleIDs(lane):
                    path = self.sumo.vehicle.getNextTLS(v)
                    if not (len(path) > 0 and path[0][2] <= self.max_distance):
                        continue
                    allv.add(v)
                    if v in self.waiting[s]:
                        self.waiting[s][v] += self.step_length
                    elif self.sumo.vehicle.getWaitingTime(v) > 0:
                        self.waiting[s][v] = self.sumo.vehicle.getWaitingTime(v)
                    w = self.waiting[s].get(v, 0)
                    if w > 0:
                        q += 1; tw += w; mw = max(mw, w)
                    else:
                        a += 1
                full[lane] = (q, a, tw, mw)
            if self.last[s] is not None:
                for v in self.last[s] - allv:
                    self.waiting[s].pop(v, None)
            self.last[s] = allv
            out[s] = full
        return out

    def step(self, act):
        tl = self.sumo.trafficlight
        for s in self.ids:                      # Signal.prep_phase
            cur = tl.getPhase(s)
            self.next_phase[s] = act[s]
            key = f"{cur}_{act[s]}"
            if cur != act[s] and key in self.yellow[s]:
                tl.setPhase(s, self.yellow[s][key])
        for _ in range(self.yellow_length):
            self.sumo.simulationStep()
        for s in self.ids:                      # Signal.set_phase
            tl.setPhase(s, int(self.next_phase[s]))
        for _ in range(self.step_length - self.yellow_length):
            self.sumo.simulationStep()
        return self.observe(), {s: tl.getPhase(s) for s in self.ids}


def _oracle(n=1):
    from pyoracle import OracleSim
    return lambda m: OracleSim(m, n, seed=0)


def _actions(env_ids, n_green, step):
    return {s: (step // 2 + i) % n_green[s] for i, s in enumerate(env_ids)}


def _check_against_fused(ref, env, steps):
    n_green = {s: len(env.phases[s]) for s in env.signal_ids}
    env.reset()
    ref.observe()
    for step in range(steps):
        act = _actions(env.signal_ids, n_green, step)
        env.step(act)
        full, phases = ref.step(act)
        for s in env.signal_ids:
            assert phases[s] == env.signals[s].phase, (step, s)
            for lane in env.signals[s].lanes:
                fo = env.signals[s].full_observation[lane]
                assert full[s][lane] == (fo['queue'], fo['approach'], fo['total_wait'], fo['max_wait']), (step, s, lane)


def test_facade_call_sequence_matches_fused_step_on_the_oracle():
    ref = RefStyleEnv("cologne3", _oracle(), seed=1)
    env = MultiSignal("f", "cologne3", None, states.mplight, rewards.wait, step_length=10, yellow_length=3, log_dir=None,
                      backend=_oracle(), seed=1)
    _check_against_fused(ref, env, 40)
    env.close()


@pytest.mark.gpu
def test_facade_over_the_cuda_backend():
    """The same per-call sequence with the CUDA library behind the facade (VecSim, N = 1: rs_tick, rs_set_phase,
    rs_dump_vehicles, rs_get_phases) against the fused rs_env_step path of a second CUDA instance."""
    from resco_b200.sim import VecSim
    ref = RefStyleEnv("cologne8", lambda m: VecSim(m, 1, seed=0), seed=1)
    env = MultiSignal("f", "cologne8", None, states.mplight, rewards.wait, step_length=10, yellow_length=3, log_dir=None, seed=1)
    _check_against_fused(ref, env, 30)
    env.close()


def _hand_driven_vs_step(backend):
    """`for ts in signal_ids: signals[ts].observe(...)` (multi_signal.py:185-186) must latch the waiting
    times once per tick, not once per signal."""
    kw = dict(step_length=10, yellow_length=3, log_dir=None, seed=3)
    if backend is not None:
        kw["backend"] = backend
    a = MultiSignal("h", "cologne3", None, states.drq_norm, rewards.wait, **kw)
    if backend is not None:
        kw["backend"] = backend
    b = MultiSignal("h", "cologne3", None, states.drq_norm, rewards.wait, **kw)
    a.reset(); b.reset()
    n_green = {s: len(a.phases[s]) for s in a.signal_ids}
    for step in range(30):
        act = _actions(a.signal_ids, n_green, step)
        oa, ra, _, _ = a.step(act)
        for s in b.signal_ids:                              # the reference's MultiSignal.step, spelled out
            b.signals[s].prep_phase(act[s])
        for _ in range(b.yellow_length):
            b.step_sim()
        for s in b.signal_ids:
            b.signals[s].set_phase()
        for _ in range(b.step_length - b.yellow_length):
            b.step_sim()
        for s in b.signal_ids:
            b.signals[s].observe(b.step_length, b.max_distance)
        ob, rb = b.state_fn(b.signals), b.reward_fn(b.signals)
        for s in a.signal_ids:
            np.testing.assert_array_equal(oa[s], ob[s], err_msg=f"step {step} {s}")
            assert ra[s] == rb[s], (step, s, ra[s], rb[s])
    a.close(); b.close()


def test_hand_driven_signals_match_env_step_on_the_oracle():
    _hand_driven_vs_step(_oracle())


@pytest.mark.gpu
def test_hand_driven_signals_match_env_step_on_cuda():
    _hand_driven_vs_step(None)


def test_queue_maxwait_rewards():
    """rewards.queue_maxwait / queue_maxwait_neighborhood (rewards.py:44-69): KeyError without an 'MA2C' mdp entry, like
    the reference as shipped; with one, dict view == batched view."""
    import torch  # noqa: F401
    from pyoracle import OracleSim
    env = MultiSignal("q", "cologne8", None, states.mplight, rewards.queue_maxwait_neighborhood, step_length=10,
                      yellow_length=3, log_dir=None, n_env=2, seed=2, backend=lambda m: OracleSim(m, 2, seed=0))
    env.reset()
    env._refresh_views(0)
    with pytest.raises(KeyError):
        rewards.queue_maxwait(env.signals)
    env.scenario.meta.setdefault('mdp', {})['MA2C'] = {'coef': 0.4, 'coop_gamma': 0.9}
    n_green = np.array([len(env.phases[s]) for s in env.signal_ids])
    for step in range(30):
        act = ((step // 2 + np.arange(2)[:, None] + np.arange(len(n_green))[None, :]) % n_green[None, :]).astype(np.int32)
        _, rew, _, _ = env.step(act)
    for inst in (0, 1):
        env._refresh_views(inst)
        own, nb = rewards.queue_maxwait(env.signals), rewards.queue_maxwait_neighborhood(env.signals)
        assert any(v != 0 for v in own.values())
        got_own = rewards.queue_maxwait.batched(env)[inst].numpy()
        np.testing.assert_allclose(got_own, [own[s] for s in env.signal_ids], rtol=1e-6)
        np.testing.assert_allclose(rew[inst].numpy(), [nb[s] for s in env.signal_ids], rtol=1e-6)
    env.close()
