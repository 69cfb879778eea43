"""
DistRegWithPruning ('clipper+prune') with the interface of the reference class of the same name
(reference roman/align/dist_reg_with_pruning.py:15-99): putative associations are the all-to-all
set minus pairs whose descriptor cosine or shape ratios fall below the thresholds; the survivors are
scored with the EuclideanDistance invariant and solved; with use_gravity the estimated rotation
must have |roll|, |pitch| below roll_pitch_thresh, else GravityConstraintError.

The pruning is expressed here as one boolean keep-mask over the (n1, n2) grid; row-major order of
the surviving pairs equals the order the reference obtains by deleting rows from the all-to-all list.
"""
from __future__ import annotations

from typing import List

import numpy as np

from .. import clipperpy
from .object_registration import ObjectRegistration


class GravityConstraintError(Exception):
    pass


def zyx_euler(R: np.ndarray):
    """(yaw, pitch, roll) with R = Rz(yaw) Ry(pitch) Rx(roll) — the 'ZYX' intrinsic convention the
    reference gets from scipy's Rotation.as_euler('ZYX') (dist_reg_with_pruning.py:41)."""
    return (float(np.arctan2(R[1, 0], R[0, 0])),
            float(-np.arcsin(np.clip(R[2, 0], -1.0, 1.0))),
            float(np.arctan2(R[2, 1], R[2, 2])))


class DistRegWithPruning(ObjectRegistration):

    def __init__(self, sigma, epsilon, mindist=0.0, shape_epsilon=0.0, cos_min=0.85,
                 dim=3, use_gravity=False, roll_pitch_thresh=np.deg2rad(5)):
        super().__init__(dim)
        assert not use_gravity or dim == 3, "Gravity can only be used with 3D points"
        self.sigma, self.epsilon, self.mindist = sigma, epsilon, mindist
        self.shape_epsilon, self.cos_min = shape_epsilon, cos_min
        self.use_gravity, self.roll_pitch_thresh = use_gravity, roll_pitch_thresh

    # ---- pieces the base class asks for
    def _setup_clipper(self):
        ip = clipperpy.invariants.EuclideanDistanceParams()
        ip.sigma, ip.epsilon, ip.mindist = self.sigma, self.epsilon, self.mindist
        return clipperpy.CLIPPER(clipperpy.invariants.EuclideanDistance(ip), clipperpy.Params())

    def _object_to_clipper_list(self, object):
        return np.asarray(object.center, dtype=np.float64).reshape(-1)[:self.dim].tolist()

    @staticmethod
    def _shape_table(object_map: List) -> np.ndarray:
        """(n, 4): volume, linearity, planarity, scattering"""
        return np.array([[o.volume, o.linearity, o.planarity, o.scattering] for o in object_map], dtype=np.float64)

    def putative_associations(self, map1: List, map2: List) -> np.ndarray:
        """(m, 2) int32 associations that survive the semantic and shape pruning."""
        d1 = np.stack([np.asarray(o.semantic_descriptor, dtype=np.float64).ravel() for o in map1])
        d2 = np.stack([np.asarray(o.semantic_descriptor, dtype=np.float64).ravel() for o in map2])
        keep = ~((d1 @ d2.T) < self.cos_min)                    # the reference drops `sim < cos_min`
        s1, s2 = self._shape_table(map1)[:, None, :], self._shape_table(map2)[None, :, :]
        with np.errstate(divide="ignore", invalid="ignore"):
            ratio = np.minimum(s1, s2) / np.maximum(s1, s2)     # (n1, n2, 4)
        keep &= ~np.any(ratio < self.shape_epsilon, axis=2)     # ... and any `ratio < shape_epsilon`
        i1, i2 = np.nonzero(keep)                               # row-major == order of the pruned all-to-all list
        return np.stack([i1, i2], axis=1).astype(np.int32)

    def _score_pruned_assoc(self, clipper, map1, map2):
        A_put = self.putative_associations(map1, map2)
        P1 = np.array([self._object_to_clipper_list(o) for o in map1])
        P2 = np.array([self._object_to_clipper_list(o) for o in map2])
        self._check_clipper_arrays(P1, P2)
        clipper.score_pairwise_consistency(P1.T, P2.T, A_put)
        return clipper, A_put

    # ---- public entry point
    def register(self, map1: List, map2: List, u0=None):
        if len(map1) == 0 or len(map2) == 0:
            return np.array([[]])
        clipper, _ = self._score_pruned_assoc(self._setup_clipper(), map1, map2)
        clipper.solve(u0)
        inliers = clipper.get_selected_associations()
        if self.use_gravity:
            R = self.T_align(map1, map2, inliers)[:self.dim, :self.dim]
            _, pitch, roll = zyx_euler(R)
            if not (abs(roll) < self.roll_pitch_thresh and abs(pitch) < self.roll_pitch_thresh):
                raise GravityConstraintError(f"Roll and pitch must be less than {self.roll_pitch_thresh} rad")
        return inliers
