import sys, time, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/oracle'); sys.path.insert(0,'/root/repo/tests')
import util
from pyoracle import OracleSim
from resco_b200.sim import VecSim
for map_name in ['cologne1','cologne8']:
    sc, m = util.marshal_map(map_name)
    n_env=2
    g = VecSim(m, n_env, seed=7); o = OracleSim(m, n_env, seed=7); g.reset(7,0); o.reset(7,0)
    g.observe(); o.observe()
    for step in range(60):
        act = util.cyclic_actions(m, n_env, step)
        g.env_step(act); o.env_step(act)
        try:
            util.assert_same_obs(g.obs(), o.obs(), f'step {step}')
            util.assert_same_state(g, o, 0, f'step {step}')
        except AssertionError as e:
            print(map_name, 'MISMATCH', str(e)[:600]); break
    else:
        print(map_name, 'OK 60 steps', g.stats()[0], 'ms/step', g.last_step_ms())
