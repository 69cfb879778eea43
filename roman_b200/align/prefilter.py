"""
Prefilters of the all-pairs search — the tests the reference makes per submap pair BEFORE registering
it (roman/align/submap_align.py:96-149), here evaluated for all S0 x S1 pairs at once:

  * `submap_distances`   centre distance (:96-99) — decides `skip_distance` (:136) and "nearby" (:101);
  * `aabb_overlap`       bounding-box test used for "nearby" when submaps are force-filled or have no
                         radius (:102; roman/utils.py:160-169);
  * `submap_similarities` max pairwise cosine of the stacked frame descriptors (roman/map/map.py:145-162),
                         on the device (rb_submap_similarity) — a pair is registered only if this is
                         not below `submap_descriptor_thresh` (:144);
  * `candidate_mask`     the resulting work list.
"""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np


def submap_distances(positions0: np.ndarray, positions1: np.ndarray) -> np.ndarray:
    """(S0, S1) Euclidean distances between submap centres."""
    p0 = np.asarray(positions0, dtype=np.float64).reshape(len(positions0), -1)
    p1 = np.asarray(positions1, dtype=np.float64).reshape(len(positions1), -1)
    return np.linalg.norm(p0[:, None, :] - p1[None, :, :], axis=2)


def aabb_overlap(points0: Sequence[np.ndarray], points1: Sequence[np.ndarray]) -> np.ndarray:
    """(S0, S1) bool: do the axis-aligned boxes of point sets points0[i], points1[j] (each (n, 3), global frame)
    intersect (closed intervals, as roman/utils.py:160-169)."""
    def boxes(pts):
        lo = np.array([np.min(np.asarray(p, dtype=np.float64)[:, :3], axis=0) for p in pts]).reshape(-1, 3)
        hi = np.array([np.max(np.asarray(p, dtype=np.float64)[:, :3], axis=0) for p in pts]).reshape(-1, 3)
        return lo, hi
    lo0, hi0 = boxes(points0)
    lo1, hi1 = boxes(points1)
    return np.all((lo0[:, None, :] <= hi1[None, :, :]) & (hi0[:, None, :] >= lo1[None, :, :]), axis=2)


def nearby_mask(sm_params, dist: np.ndarray, points0=None, points1=None) -> np.ndarray:
    """The condition under which the reference fills robots_nearby_mat[i, j] (submap_align.py:101-103)."""
    by_radius = (not sm_params.force_fill_submaps) and sm_params.submap_radius is not None
    if by_radius:
        return dist < 2.0 * sm_params.submap_radius
    if points0 is None or points1 is None:
        raise ValueError("force-filled / radius-less submaps need the object centres for the bounding-box test")
    return aabb_overlap(points0, points1)


def submap_similarities(descriptors0: Sequence, descriptors1: Sequence, engine=None) -> np.ndarray:
    """(S0, S1) float64 similarity matrix on the device; descriptors*[i] is (N_i, D) stacked frame descriptors
    or a (D,) mean descriptor."""
    from ..engine import Engine
    eng = engine or Engine()
    return eng.submap_similarity(descriptors0, descriptors1).cpu().numpy()


def candidate_mask(sm_params, S0: int, S1: int, dist: Optional[np.ndarray] = None, skip_distance: float = np.inf,
                   similarity: Optional[np.ndarray] = None) -> np.ndarray:
    """Pairs the reference would register: centre distance not above `skip_distance` (submap_align.py:136) and,
    when a submap descriptor is in use, similarity not below the threshold (:131-134, :144)."""
    mask = np.ones((S0, S1), dtype=bool)
    if dist is not None:
        mask &= ~(dist > skip_distance)
    if sm_params.submap_descriptor is not None and similarity is not None:
        mask &= ~(similarity < sm_params.submap_descriptor_thresh)
    return mask
