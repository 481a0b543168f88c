"""Sharding of the matching path over the GPUs of one box.

Views are split into equal contiguous blocks, one per rank (= the block whose 2D segments the rank holds before the
all-gather).  The deduplicated view-pair list is built exactly like the reference does (computeMatches,
line3D.cc:704-741: views ascending, a pair is taken the first time either side lists the other) and every pair goes to
the rank that owns its SOURCE view: no pair is evaluated twice, none is dropped, and the arithmetic of a pair does not
depend on which rank runs it (multi-GPU result == single-GPU result, bit for bit).
"""
from __future__ import annotations

import numpy as np


def view_range(rank: int, world: int, num_views: int):
    per = (num_views + world - 1) // world
    return rank * per, min(num_views, (rank + 1) * per)


def owner_of_view(view: np.ndarray, world: int, num_views: int) -> np.ndarray:
    per = (num_views + world - 1) // world
    return np.asarray(view) // per


def rank_pairs(pairs: np.ndarray, rank: int, world: int, num_views: int) -> np.ndarray:
    """rows of `pairs` (src, tgt) whose source view belongs to `rank`, in their original order"""
    pairs = np.asarray(pairs).reshape(-1, 2)
    return pairs[owner_of_view(pairs[:, 0], world, num_views) == rank]
