import sys, os
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0,ROOT); sys.path.insert(0,os.path.join(ROOT,'tests')); sys.path.insert(0,os.path.join(ROOT,'oracle'))
import numpy as np, util
from pyoracle import OracleSim
from resco_b200.sim import VecSim
sc,m=util.marshal_map("cologne8", tile_vcap=128)
n=19
g=VecSim(m,n,seed=7); o=OracleSim(m,n,seed=7); g.reset(7,0); o.reset(7,0); g.observe(); o.observe()
print(g.tile_info(), g.launch_shape())
for step in range(120):
    act=util.cyclic_actions(m,n,step)
    g.env_step(act); o.env_step(act)
    og,oo=g.obs(),o.obs()
    sg,so=g.stats(),o.stats()
    bad=[k for k in util.OBS_EXACT if not np.array_equal(og[k],oo[k])]
    ti=g.tile_info()
    if bad or not np.array_equal(sg['n_active'],so['n_active']) or step%20==0:
        envs=sorted(set(np.argwhere(og['lane_queue']!=oo['lane_queue'])[:,0].tolist()))
        print('step',step,'bad',bad,'envs',envs,'n_active gpu',sg['n_active'].tolist(),'orc',so['n_active'].tolist(),'tick',sg['tick'].tolist()[:3], ti)
        if bad: break
