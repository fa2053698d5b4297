# golden for generate_config: run the REFERENCE Signal for grid4x4 under a map name that has no signal_configs entry
import sys, os, json, types, io, contextlib
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/oracle'); sys.path.insert(0,'/root/repo/tools')
import make_golden as mg
mg.install_stubs()
import importlib
from resco_b200.multi_signal import load_scenario
from resco_b200.traci_facade import open_facade
from pyoracle import OracleSim
ts = importlib.import_module('resco_benchmark.traffic_signal')
sc = load_scenario('grid4x4')
f = open_facade(sc, lambda m: OracleSim(m, 1, seed=0), step_length=10, yellow_length=3, max_distance=200.0, seed=0)
ms_phases = {t: [mg.Phase(d, s) for d, s in sc.meta['programs'][t] if 'y' not in s and 'g' in s.lower()] for t in sc.meta['tls_ids']}
out = {}
for t in sc.meta['tls_ids']:
    with contextlib.redirect_stdout(io.StringIO()):
        sig = ts.Signal('no_such_map', f, t, 3, ms_phases[t]) if False else None
    # signal_configs['no_such_map'] raises KeyError before generate_config: emulate `self.id not in myconfig` with an empty config
    cfgmod = importlib.import_module('resco_benchmark.config.signal_config')
    cfgmod.signal_configs['gridX'] = {}
    with contextlib.redirect_stdout(io.StringIO()):
        sig = ts.Signal('gridX', f, t, 3, ms_phases[t])
    out[t] = dict(lanes=sig.lanes, lane_sets=sig.lane_sets, downstream=sig.downstream)
json.dump(out, open('/root/repo/tests/golden/generate_config_grid4x4.json','w'))
print('wrote', len(out), 'signals; A0 lanes', out['A0']['lanes'][:4], out['A0']['downstream'])
