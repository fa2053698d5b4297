#!/usr/bin/env python
"""Compile the reference's scenario files into flat tables (run in the dev container only).

Reads  /root/reference/resco_benchmark/{environments,config}  (read-only inputs; nothing is copied
verbatim) and writes  resco_b200/data/<map>.npz.  The GPU box has no reference tree, so the compiled
tables are committed.  Usage:  python tools/compile_scenarios.py [map ...]
"""
import os
import runpy
import sys
import zipfile

sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..'))
from resco_b200.scenario.netxml import read_net, read_routes_xml, read_sumocfg  # noqa: E402
from resco_b200.scenario.compiler import compile_scenario  # noqa: E402

REF = os.environ.get('RESCO_REFERENCE', '/root/reference/resco_benchmark')
OUT = os.path.join(os.path.dirname(__file__), '..', 'resco_b200', 'data')
EPISODES = int(os.environ.get('RESCO_EPISODES', '10'))


def main(maps):
    signal_configs = runpy.run_path(os.path.join(REF, 'config', 'signal_config.py'))['signal_configs']
    map_configs = runpy.run_path(os.path.join(REF, 'config', 'map_config.py'))['map_configs']
    mdp_configs = runpy.run_path(os.path.join(REF, 'config', 'mdp_config.py'))['mdp_configs']
    os.makedirs(OUT, exist_ok=True)
    for m in maps:
        mc = map_configs[m]
        netp = os.path.join(REF, mc['net'])
        if netp.endswith('.sumocfg'):
            cfg = read_sumocfg(netp)
            net = read_net(cfg['net'])
            demand = read_routes_xml(cfg['routes'][0])
            begin = cfg['begin']
        else:
            net = read_net(netp)
            # multi_signal.py:33-37,124: route file <route>/<map>_<run>.rou.xml, run = 1 here
            zp = os.path.join(REF, 'environments', m, m + '.zip')
            # one demand table per episode: the first EPISODES route files (the zip ships 1400; MultiSignal.reset()
            # cycles through the compiled ones, run r -> file ((r - 1) % EPISODES) + 1)
            with zipfile.ZipFile(zp) as z:
                demand = [read_routes_xml(z.read(f'{m}_{r}.rou.xml').decode(), is_text=True) for r in range(1, EPISODES + 1)]
            begin = float(mc['start_time'])
        sc = compile_scenario(net, demand, m, mc, signal_configs[m], begin)
        # hyper-parameters + manager/worker regions of the FMA2C states/rewards (config/mdp_config.py)
        sc.meta['mdp'] = {k: mdp_configs[k][m] for k in ('FMA2C', 'FMA2CFull') if m in mdp_configs.get(k, {})}
        path = os.path.join(OUT, m + '.npz')
        sc.save(path)
        a = sc.arrays
        print(f"{m}: lanes={sc.n_lanes} links={sc.n_links} tls={sc.n_tls} signals={len(sc.meta['signal_ids'])} "
              f"routes={len(a['route_off']) - 1} trips={len(a['trip_depart'])} (file {sc.meta['n_trips_file']}, "
              f"unroutable {sc.meta['n_unroutable']}) origins={len(a['origin_lane'])} "
              f"foes={len(a['foe_link'])} -> {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == '__main__':
    main(sys.argv[1:] or ['cologne1', 'cologne3', 'cologne8', 'ingolstadt1', 'ingolstadt7', 'ingolstadt21',
                          'grid4x4', 'arterial4x4'])
