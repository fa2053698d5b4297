"""Batched agent front ends: the parts of the reference's agents that are evaluated once per
env step for every instance, restated for [N, S, ...] observation batches.  Agent-side code: torch library ops, not part
of the simulation path."""
from .frap import BatchedFRAP, competition_mask  # noqa: F401
