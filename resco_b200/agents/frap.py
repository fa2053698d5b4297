"""Batched forward of the reference's FRAP Q-network (agents/mplight.py:43-131), the model behind MPLight.

The reference evaluates it with a Python loop over the batch (`for i in range(batch_size)`, mplight.py:82-89), a Python
loop over the movements (:93-99) and materialises all n x (n - 1) concatenated pair embeddings (:105-113).  With one
observation row per (instance, signal) -- 65536 x 21 rows per env step in BASELINE configs[3] -- that is the bottleneck of
the act() side.  This module keeps the reference's parameters (same names and shapes: a reference ``state_dict`` loads
with ``load_state_dict``) and computes the same Q-values with batched tensor ops:

* the phase one-hot comes from a [n_pairs, n_movements] table lookup instead of the per-row loop;
* the 1x1 "lane" convolution over cat(pair_i, pair_j) is split into its two halves, W_a pair_i + W_b pair_j + b, so the
  [B, n, n-1, 32] tensor of concatenations is never built;
* the relation branch depends on the competition mask only, not on the batch: it is evaluated once per call;
* rows are processed in chunks so that the [chunk, n, n-1, 20] intermediates stay bounded.

Quirks kept as they are: the demand slice is ``states[:, i:i+demand_shape]`` (mplight.py:95 -- movement i, not
i*demand_shape; identical for the shipped demand_shape = 1), and a pair with two equal movements marks one position.
The value head (pfrl's DiscreteActionValueHead) is a wrapper around the Q tensor; ``forward`` returns the tensor and
``act`` its argmax restricted to the signal's valid actions (agents/agent.py SharedAgent + pfrl's greedy evaluation).
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F


def competition_mask(phase_pairs: Sequence[Sequence[int]]) -> np.ndarray:
    """[n, n-1] 0/1: pair i and the j-th OTHER pair share exactly one movement (agents/mplight.py:19-31)."""
    n = len(phase_pairs)
    out = np.zeros((n, n - 1), np.int64)
    for i in range(n):
        cnt = 0
        for j in range(n):
            if i == j:
                continue
            if len(set(list(phase_pairs[i]) + list(phase_pairs[j]))) == 3:
                out[i, cnt] = 1
            cnt += 1
    return out


class BatchedFRAP(nn.Module):
    def __init__(self, phase_pairs: Sequence[Sequence[int]], demand_shape: int = 1, chunk_rows: int = 32768):
        super().__init__()
        self.phase_pairs = [list(p) for p in phase_pairs]
        self.oshape = len(self.phase_pairs)
        self.demand_shape = int(demand_shape)
        self.chunk_rows = int(chunk_rows)
        self.d_out, self.p_out, self.lane_embed_units = 4, 4, 16
        relation_embed_size = 4
        # parameters: names and shapes of agents/mplight.py:59-71
        self.p = nn.Embedding(2, self.p_out)
        self.d = nn.Linear(self.demand_shape, self.d_out)
        self.lane_embedding = nn.Linear(self.p_out + self.d_out, self.lane_embed_units)
        self.lane_conv = nn.Conv2d(2 * self.lane_embed_units, 20, kernel_size=(1, 1))
        self.relation_embedding = nn.Embedding(2, relation_embed_size)
        self.relation_conv = nn.Conv2d(relation_embed_size, 20, kernel_size=(1, 1))
        self.hidden_layer = nn.Conv2d(20, 20, kernel_size=(1, 1))
        self.before_merge = nn.Conv2d(20, 1, kernel_size=(1, 1))
        n = self.oshape
        self.register_buffer("comp_mask", torch.from_numpy(competition_mask(self.phase_pairs)), persistent=False)
        self.register_buffer("pair_a", torch.tensor([p[0] for p in self.phase_pairs]), persistent=False)
        self.register_buffer("pair_b", torch.tensor([p[1] for p in self.phase_pairs]), persistent=False)
        # for pair i, the indices of the n-1 other pairs in the reference's (i, j != i) order
        others = torch.tensor([[j for j in range(n) if j != i] for i in range(n)])
        self.register_buffer("others", others, persistent=False)
        self._onehot: Optional[torch.Tensor] = None

    def _pair_onehot(self, num_movements: int, device) -> torch.Tensor:
        if self._onehot is None or self._onehot.shape[1] != num_movements or self._onehot.device != device:
            oh = torch.zeros((self.oshape, num_movements), dtype=torch.int64, device=device)
            for i, (a, b) in enumerate(self.phase_pairs):
                oh[i, a] = 1
                oh[i, b] = 1
            self._onehot = oh
        return self._onehot

    def _relations(self) -> torch.Tensor:
        """[n, n-1, 20]: relu(conv(relu(embed(mask)))) -- batch independent (mplight.py:116-119)."""
        r = F.relu(self.relation_embedding(self.comp_mask))                              # [n, n-1, 4]
        return F.relu(F.linear(r, self.relation_conv.weight.flatten(1), self.relation_conv.bias))

    def _chunk(self, states: torch.Tensor, relations: torch.Tensor) -> torch.Tensor:
        B = states.shape[0]
        M = int((states.shape[1] - 1) / self.demand_shape)
        acts = states[:, 0].to(torch.int64)
        x = states[:, 1:].float()
        ext = self._pair_onehot(M, states.device)[acts]                                  # [B, M]
        phase = torch.sigmoid(self.p(ext))                                               # [B, M, 4]
        if self.demand_shape == 1:
            dem_in = x[:, :M].unsqueeze(-1)
        else:                                                                           # states[:, i:i+ds] (reference slice)
            dem_in = torch.stack([x[:, i:i + self.demand_shape] for i in range(M)], 1)
        demand = torch.sigmoid(self.d(dem_in))                                           # [B, M, 4]
        pd = F.relu(self.lane_embedding(torch.cat((phase, demand), -1)))                 # [B, M, 16]
        pairs = pd[:, self.pair_a] + pd[:, self.pair_b]                                  # [B, n, 16]
        w = self.lane_conv.weight.flatten(1)                                             # [20, 32]
        E = self.lane_embed_units
        first = F.linear(pairs, w[:, :E], self.lane_conv.bias)                           # [B, n, 20]  W_a pair_i + b
        second = F.linear(pairs, w[:, E:])                                               # [B, n, 20]  W_b pair_j
        rot = F.relu(first.unsqueeze(2) + second[:, self.others])                        # [B, n, n-1, 20]
        comb = rot * relations.unsqueeze(0)
        comb = F.relu(F.linear(comb, self.hidden_layer.weight.flatten(1), self.hidden_layer.bias))
        comb = F.linear(comb, self.before_merge.weight.flatten(1), self.before_merge.bias)   # [B, n, n-1, 1]
        return comb.squeeze(-1).sum(-1)                                                  # [B, n]

    def forward(self, states: torch.Tensor) -> torch.Tensor:
        """states [B, 1 + n_movements * demand_shape] (row = states.mplight of one signal: phase index first) or
        [N, S, ...] (flattened internally) -> Q-values [B, n_pairs] / [N, S, n_pairs]."""
        lead = states.shape[:-1]
        flat = states.reshape(-1, states.shape[-1])
        rel = self._relations()
        out = [self._chunk(flat[i:i + self.chunk_rows], rel) for i in range(0, flat.shape[0], self.chunk_rows)]
        q = torch.cat(out, 0) if len(out) != 1 else out[0]
        return q.reshape(*lead, self.oshape)

    @torch.no_grad()
    def act(self, states: torch.Tensor, valid_acts: Optional[Dict[str, Dict]] = None,
            signal_ids: Optional[Sequence[str]] = None) -> torch.Tensor:
        """Greedy actions for an [N, S, 13] batch.  With the map's ``valid_acts`` the argmax runs over the signal's valid
        pair indices and is translated to the signal's local action index (agents/agent.py:46-60 with reverse_valid);
        without it the pair index is the action."""
        q = self.forward(states)
        if valid_acts is None:
            return q.argmax(-1).to(torch.int32)
        assert signal_ids is not None and q.dim() == 3 and q.shape[1] == len(signal_ids)
        # the reference scans the signal's valid pair indices in dict order with a strict '>' (pfrl_dqn.py:131-142):
        # gather the Q-values in that order and take the first maximum; rows are padded with their first entry,
        # which can never win a tie against itself at position 0
        kmax = max(len(valid_acts[sid]) for sid in signal_ids)
        ord_idx = torch.zeros((len(signal_ids), kmax), dtype=torch.int64, device=q.device)
        ord_act = torch.zeros((len(signal_ids), kmax), dtype=torch.int32, device=q.device)
        for s, sid in enumerate(signal_ids):
            items = [(int(k), int(a)) for k, a in valid_acts[sid].items()]
            items += [items[0]] * (kmax - len(items))
            ord_idx[s] = torch.tensor([k for k, _ in items])
            ord_act[s] = torch.tensor([a for _, a in items])
        qsel = q.gather(2, ord_idx.unsqueeze(0).expand(q.shape[0], -1, -1))               # [N, S, kmax]
        first_max = qsel.argmax(-1)                                                       # torch: first maximum
        return ord_act.unsqueeze(0).expand(q.shape[0], -1, -1).gather(2, first_max.unsqueeze(-1)).squeeze(-1)
