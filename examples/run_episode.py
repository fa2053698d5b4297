#!/usr/bin/env python
"""One episode of a RESCO map on the vectorised backend with the reference's MAXPRESSURE / MAXWAVE rule.

    python examples/run_episode.py --map cologne8 --n-env 4096                  # B200: N lock-step instances

Prints the per-instance average delay (utils/readXML.py definition) and the env-step throughput.  The agent is the
library's batched front end on the device (`VecSim.policy_maxpressure`).  There is no CPU backend: without the CUDA
library or a Blackwell GPU the constructor raises (episodes on the CPU oracle, for development: tools/anchors.py)."""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import resco_b200.rewards as rewards                      # noqa: E402
import resco_b200.states as states                        # noqa: E402
from resco_b200.multi_signal import MultiSignal           # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--map", default="cologne8")
    ap.add_argument("--n-env", type=int, default=1024)
    ap.add_argument("--agent", default="MAXPRESSURE", choices=["MAXPRESSURE", "MAXWAVE"])
    ap.add_argument("--seed", type=int, default=1)
    a = ap.parse_args()
    use_wave = a.agent == "MAXWAVE"
    state_fn = states.wave if use_wave else states.mplight
    from resco_b200.multi_signal import load_scenario
    mc = load_scenario(a.map).meta["map_config"]
    env = MultiSignal("example", a.map, None, state_fn, rewards.pressure, end_time=mc["end_time"], step_length=mc["step_length"],
                      yellow_length=mc["yellow_length"], max_distance=50 if use_wave else 200, log_dir=None,
                      n_env=max(a.n_env, 2), seed=a.seed)
    sc = env.scenario
    pairs, va, sig = sc.meta["phase_pairs"], sc.meta["valid_acts"], env.signal_ids
    act = lambda obs: env.sim.policy_maxpressure(pairs, va, sig, use_wave=use_wave)     # noqa: E731
    obs = env.reset()
    steps, done, t0 = 0, False, time.perf_counter()
    while not done:
        obs, rew, done, info = env.step(act(obs))
        steps += 1
    dt = time.perf_counter() - t0
    st = env.episode_stats()
    d = st["avg_delay"]
    print(f"{a.map} {a.agent}: {env.n_env} instances x {steps} env steps in {dt:.2f} s "
          f"({env.n_env * steps / dt:,.0f} env steps/s incl. the Python loop)")
    print(f"avg delay per instance: mean {d.mean():.1f} s, min {d.min():.1f}, max {d.max():.1f} "
          f"(trips per instance {int(st['trips'].mean())})")
    env.close()


if __name__ == "__main__":
    main()
