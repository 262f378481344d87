"""Sharding of oversample groups over GPUs / ranks.

Every batch ("oversample group") adds independently to the histograms
(``src/Analysis.cpp:821-833``; mixed-event partners come from the same batch,
``src/HBT_correlation.cpp:197,209``), so groups are dealt round-robin to ranks and the
histograms are summed once at the end.  The only cross-batch coupling is the shared mt19937
stream: its consumption per batch depends on the batch's event counts only, so every rank
replays the draws of the groups it does not own (``Random.skip_batch``) and arrives at exactly
the draws the single-process run would make for its own groups.
"""
from __future__ import annotations

from typing import Iterable, Iterator, Sequence, Tuple


def owner(group: int, world: int) -> int:
    return group % world


def my_groups(rank: int, world: int, n_groups: int) -> range:
    return range(rank, n_groups, world)


def walk(rank: int, world: int, event_counts: Sequence[Tuple[int, int]], rng) -> Iterator[int]:
    """Yield the indices of this rank's groups in stream order, fast-forwarding ``rng`` (an
    object with ``skip_batch(nev, nev_mixed)``) past every group owned by another rank.  The
    caller must consume the draws of each yielded group before asking for the next one."""
    for g, (nev, nev_mixed) in enumerate(event_counts):
        if owner(g, world) == rank:
            yield g
        else:
            rng.skip_batch(nev, nev_mixed)
