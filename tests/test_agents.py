"""Batched agent front ends against vectors recorded from the reference's own classes (tools/make_golden_frap.py)."""
import glob
import json
import os

import numpy as np
import pytest
import torch

import util
from resco_b200.agents import BatchedFRAP, competition_mask

FIX = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "agents", "frap_*.npz")))


def _load(path):
    z = np.load(path)
    meta = json.loads(bytes(z["meta"]).decode())
    sc = util.load(meta["map"])
    model = BatchedFRAP(sc.meta["phase_pairs"], demand_shape=1, chunk_rows=100)    # odd chunk size: exercises chunking
    sd = {k[len("param."):]: torch.from_numpy(z[k]) for k in z.files if k.startswith("param.")}
    model.load_state_dict(sd)            # strict: the reference's parameter names and shapes
    return z, meta, sc, model


@pytest.mark.parametrize("path", FIX, ids=[os.path.basename(p)[:-4] for p in FIX])
def test_batched_frap_reproduces_the_reference_forward(path):
    z, meta, sc, model = _load(path)
    with torch.no_grad():
        q = model(torch.from_numpy(z["states"]))                                  # [B, S, 13] in one call
    assert q.shape == z["q"].shape
    np.testing.assert_allclose(q.numpy(), z["q"], rtol=2e-5, atol=2e-5)


@pytest.mark.parametrize("path", FIX, ids=[os.path.basename(p)[:-4] for p in FIX])
def test_batched_frap_greedy_actions_follow_valid_acts(path):
    z, meta, sc, model = _load(path)
    va = {sid: sc.meta["valid_acts"][sid] for sid in meta["signal_ids"]}
    acts = model.act(torch.from_numpy(z["states"]), va, meta["signal_ids"])
    np.testing.assert_array_equal(acts.numpy(), z["acts"])


def test_competition_mask_and_fixture_presence():
    assert len(FIX) == 2
    m = competition_mask([[0, 1], [1, 2], [3, 4]])
    assert m.tolist() == [[1, 0], [1, 0], [0, 0]]
