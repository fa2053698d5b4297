#!/usr/bin/env python
"""Golden vectors for the batched FRAP front end (dev container only): instantiate the UNMODIFIED reference class
``resco_benchmark.agents.mplight.FRAP`` (pfrl is not installed: its DiscreteActionValueHead is stubbed by a wrapper that
keeps the Q tensor, everything else is the reference's code), seed its parameters, run its own forward on random
``states.mplight`` rows and record parameters, inputs, Q-values and the greedy valid-action choice of
SharedDQN.batch_act (pfrl_dqn.py:124-163, evaluation branch).  Writes tests/golden/agents/frap_<map>.npz."""
import importlib
import json
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path.insert(0, ROOT)
REF = '/root/reference/resco_benchmark'


def install_stubs():
    pkg = types.ModuleType('resco_benchmark'); pkg.__path__ = [REF]
    sys.modules['resco_benchmark'] = pkg
    agents = types.ModuleType('resco_benchmark.agents'); agents.__path__ = [os.path.join(REF, 'agents')]
    sys.modules['resco_benchmark.agents'] = agents
    agent = types.ModuleType('resco_benchmark.agents.agent')
    agent.SharedAgent = type('SharedAgent', (), {})
    sys.modules['resco_benchmark.agents.agent'] = agent
    dqn = types.ModuleType('resco_benchmark.agents.pfrl_dqn')
    dqn.DQNAgent = type('DQNAgent', (), {})
    sys.modules['resco_benchmark.agents.pfrl_dqn'] = dqn
    pfrl = types.ModuleType('pfrl'); qf = types.ModuleType('pfrl.q_functions')

    class DiscreteActionValueHead(torch.nn.Module):      # pfrl wraps the tensor in an ActionValue; keep the tensor
        def forward(self, q):
            return q
    qf.DiscreteActionValueHead = DiscreteActionValueHead
    pfrl.q_functions = qf
    sys.modules['pfrl'] = pfrl; sys.modules['pfrl.q_functions'] = qf


def main():
    install_stubs()
    mod = importlib.import_module('resco_benchmark.agents.mplight')
    sigcfg = importlib.import_module('resco_benchmark.config.signal_config').signal_configs
    from resco_b200.agents.frap import competition_mask
    for map_name in ('cologne8', 'ingolstadt21'):
        pairs = sigcfg[map_name]['phase_pairs']
        valid = sigcfg[map_name]['valid_acts']
        torch.manual_seed(1234)
        cm = torch.from_numpy(competition_mask(pairs))
        model = mod.FRAP({'demand_shape': 1}, len(pairs), pairs, cm, torch.device('cpu'))
        for prm in model.parameters():                    # spread the parameters so that Q-values differ clearly
            torch.nn.init.normal_(prm, 0.0, 0.7)
        rng = np.random.default_rng(7)
        sig_ids = list(valid.keys())
        B = 64
        states = np.zeros((B, len(sig_ids), 13), np.float32)
        states[:, :, 0] = rng.integers(0, len(pairs), (B, len(sig_ids)))
        states[:, :, 1:] = rng.integers(-6, 25, (B, len(sig_ids), 12))
        with torch.no_grad():
            q = model(torch.from_numpy(states.reshape(-1, 13))).numpy().reshape(B, len(sig_ids), len(pairs))
        acts = np.zeros((B, len(sig_ids)), np.int32)
        for b in range(B):
            for s, sid in enumerate(sig_ids):             # SharedDQN.batch_act, evaluation branch
                max_val, max_idx = None, None
                for idx in valid[sid]:
                    if max_val is None or q[b, s, idx] > max_val:
                        max_val, max_idx = q[b, s, idx], idx
                acts[b, s] = valid[sid][max_idx]
        out = {'meta': np.frombuffer(json.dumps({'map': map_name, 'signal_ids': sig_ids}).encode(), np.uint8),
               'states': states, 'q': q.astype(np.float32), 'acts': acts}
        for k, v in model.state_dict().items():
            out['param.' + k] = v.numpy()
        path = os.path.join(ROOT, 'tests', 'golden', 'agents', 'frap_' + map_name + '.npz')
        np.savez_compressed(path, **out)
        print('wrote', path, 'q', q.shape, '%.1f KiB' % (os.path.getsize(path) / 1024))


if __name__ == '__main__':
    main()
