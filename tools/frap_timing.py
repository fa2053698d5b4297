"""Time rs_policy_frap alone (CUDA events, 30 launches after warm-up) on one half-batch of C4: 4096 ingolstadt21 instances
x 21 signals = 86016 rows of states.mplight.   usage: python tools/frap_timing.py   (RESCO_B200_LIB selects the build)"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import util  # noqa: E402
from resco_b200.sim import VecSim  # noqa: E402

z = np.load(os.path.join(ROOT, "tests", "golden", "agents", "frap_ingolstadt21.npz"))
sc, m = util.marshal_map("ingolstadt21")
sim = VecSim(m, 4096, seed=3)
sim.load_frap({k[len("param."):]: z[k] for k in z.files if k.startswith("param.")}, sc.meta["phase_pairs"],
              sc.meta["valid_acts"], m.info["signal_ids"])
g = torch.Generator(device="cuda").manual_seed(1)
obs = torch.randint(-8, 9, (4096, 21, 13), generator=g, device="cuda").float()
obs[:, :, 0] = torch.randint(0, 3, (4096, 21), generator=g, device="cuda").float()
for _ in range(5):
    sim.policy_frap(obs)
torch.cuda.synchronize()
ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(30)]
for a, b in ev:
    a.record(); sim.policy_frap(obs); b.record()
torch.cuda.synchronize()
ms = [a.elapsed_time(b) for a, b in ev]
acts = sim.policy_frap(obs).cpu().numpy()
print(json.dumps(dict(lib=os.path.basename(os.environ.get("RESCO_B200_LIB", "libresco_b200.so")), rows=4096 * 21,
                      ms_mean=float(np.mean(ms)), ms_min=float(np.min(ms)), action_checksum=int(acts.astype(np.int64).sum()))))
