"""Cycles per kernel phase (diagnostics).  Needs a library built with -DRS_PHASE_CLOCKS=1:
    tools/phase_clocks.sh            (builds ab/clocks.so here, run the printed command under gpurun)
usage: RESCO_B200_LIB=ab/clocks.so python tools/phase_clocks.py [map] [n_env] [tile_vcap] [steps] [synthetic_rate] [maxpressure|random]"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from resco_b200.sim import VecSim  # noqa: E402

NAMES = ["stage", "S0 tls", "S1 plan", "S2 move", "S3a cand", "S3b cap", "S4 prefix", "S5 merge", "S6 scatter", "S7 swap",
         "observe", "writeback", "sched"]
mp = sys.argv[1] if len(sys.argv) > 1 else "cologne8"
n_env = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
tile = int(sys.argv[3]) if len(sys.argv) > 3 else 128
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 100
rate = float(sys.argv[5]) if len(sys.argv) > 5 else 0.0
policy = sys.argv[6] if len(sys.argv) > 6 else "maxpressure"
sc, m = bench._marshal(mp, 4096 if rate > 0 else 0, tile, rate)
sim = VecSim(m, n_env, seed=1, device=0)
sim.reset(1, 0); sim.observe()
pairs, va, sig = sc.meta["phase_pairs"], sc.meta["valid_acts"], m.info["signal_ids"]
buf = (C.c_ulonglong * 24)()
sim.lib.rs_debug_phase_clocks.argtypes = [C.c_void_p, C.c_void_p]
act = (lambda: sim.policy_maxpressure(pairs, va, sig)) if policy == "maxpressure" else (lambda: sim.policy_random(1))
for _ in range(60):
    sim.env_step(act())
assert sim.lib.rs_debug_phase_clocks(sim._h, buf) == 1, "library was not built with -DRS_PHASE_CLOCKS=1"
a = list(buf)
for _ in range(steps):
    sim.env_step(act())
sim.lib.rs_debug_phase_clocks(sim._h, buf)
d = [y - x for x, y in zip(a, list(buf))]
tot = sum(d)
shape = sim.launch_shape()
print(mp, n_env, shape, "last k_run ms %.3f" % sim.last_step_ms())
for nm, v in zip(NAMES, d):
    print("  %-10s %6.2f %%   %10.0f cycles per CTA per env step" % (nm, 100.0 * v / tot, v / steps / shape["grid_ctas"]))
