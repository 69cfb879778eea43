"""
ObjectRegistration — same interface as reference roman/align/object_registration.py:17-129
(register, get_MCA, mno_clipper, T_align and InsufficientAssociationsException), running the
CLIPPER steps through roman_b200.clipperpy and T_align through the CUDA Arun kernel.
"""
from __future__ import annotations

from typing import List

import numpy as np

from .. import clipperpy


class InsufficientAssociationsException(Exception):
    """reference object_registration.py:8-15"""

    def __init__(self, map1_len, map2_len, n_associations=None):
        self.map1_len, self.map2_len, self.n_associations = map1_len, map2_len, n_associations
        super().__init__("Insufficient associations. " + " ".join(
            (f"Map 1 length: {map1_len}.", f"Map 2 length: {map2_len}.", f"Associations: {n_associations}")))


class ObjectRegistration:

    def __init__(self, dim=3):
        self.dim = dim

    # -- reference object_registration.py:22-29
    def register(self, map1: List, map2: List, u0=None):
        """Returns the (k, 2) int32 inlier associations.  `u0` (optional, not in the reference
        signature) fixes the solver's start vector, which the reference draws at random."""
        if min(len(map1), len(map2)) == 0:
            return np.array([[]])      # what the reference returns for an empty map
        solver, _ = self._clipper_score_all_to_all(self._setup_clipper(), map1, map2)
        solver.solve(u0)
        return solver.get_selected_associations()

    def _setup_clipper(self):
        raise NotImplementedError

    def _object_to_clipper_list(self, object):
        raise NotImplementedError

    def _check_clipper_arrays(self, map1_cl, map2_cl):
        return

    # -- reference object_registration.py:40-48
    def _feature_rows(self, objects: List) -> np.ndarray:
        return np.array([self._object_to_clipper_list(o) for o in objects])

    def _clipper_score_all_to_all(self, clipper, map1: List, map2: List):
        rows1, rows2 = self._feature_rows(map1), self._feature_rows(map2)
        self._check_clipper_arrays(rows1, rows2)
        every_pair = clipperpy.utils.create_all_to_all(len(map1), len(map2))
        clipper.score_pairwise_consistency(rows1.T, rows2.T, every_pair)   # the binding takes (d, n) matrices
        return clipper, every_pair

    # -- reference object_registration.py:50-55
    def get_MCA(self, map1: List, map2: List):
        scored, assoc = self._clipper_score_all_to_all(self._setup_clipper(), map1, map2)
        return scored.get_affinity_matrix(), scored.get_constraint_matrix(), assoc

    # -- reference object_registration.py:57-86
    def mno_clipper(self, map1: List, map2: List, num_solutions=2, u0s=None):
        """Multiple solutions: solve, record the clique, zero the affinities among its nodes
        (constraints stay), solve again.  Returns [(associations (k,2) int64, score)], where score
        is the Rayleigh quotient of the clique-restricted u on the ORIGINAL affinity matrix."""
        M, C, A = self.get_MCA(map1, map2)
        M_first = M.copy()
        resolver = clipperpy.CLIPPER(clipperpy.invariants.PairwiseInvariant(), clipperpy.Params())
        found = []
        for k in range(num_solutions):
            resolver.set_matrix_data(M=M, C=C)
            resolver.solve(u0s[k] if u0s is not None else None)
            sol = resolver.get_solution()
            nodes = np.asarray(sol.nodes, dtype=np.int64)
            u_clique = np.zeros_like(sol.u)
            u_clique[nodes] = sol.u[nodes]
            score = 0 if nodes.size == 0 else float(u_clique @ M_first @ u_clique / (u_clique @ u_clique))
            found.append((np.asarray(A)[nodes].astype(np.int64).reshape(-1, 2), score))
            if k + 1 < num_solutions and nodes.size:
                M[np.ix_(nodes, nodes)] = 0.0
        return found

    # -- reference object_registration.py:88-129
    def T_align(self, map1: List, map2: List, correspondences: np.ndarray = None):
        """Transformation that aligns map2 to map1 (Arun's method on the matched centres)."""
        n1, n2 = len(map1), len(map2)
        if min(n1, n2) == 0:
            raise InsufficientAssociationsException(n1, n2)
        correspondences = self.register(map1, map2) if correspondences is None else correspondences
        if len(correspondences) < self.dim:      # Arun's method needs at least `dim` point pairs
            raise InsufficientAssociationsException(n1, n2, len(correspondences))
        corr = np.asarray(correspondences, dtype=np.int64).reshape(-1, 2)
        if ((corr[:, 0] < -n1) | (corr[:, 0] >= n1) | (corr[:, 1] < -n2) | (corr[:, 1] >= n2)).any():
            raise IndexError("list index out of range")       # what map1[corr[i, 0]] raises in the reference (:110-111)
        correspondences = np.stack([corr[:, 0] % n1, corr[:, 1] % n2], axis=1)   # negative indices count from the end, as Python lists do
        pts1 = np.array([np.asarray(o.center).reshape(-1)[:self.dim] for o in map1], dtype=np.float64)
        pts2 = np.array([np.asarray(o.center).reshape(-1)[:self.dim] for o in map2], dtype=np.float64)
        T = _arun_gpu(pts1, pts2, np.asarray(correspondences), self.dim)
        return T


_arun_engine = None


def _arun_gpu(pts1: np.ndarray, pts2: np.ndarray, corr: np.ndarray, dim: int) -> np.ndarray:
    """One T_align through the CUDA Arun kernel (rb_arun)."""
    global _arun_engine
    from ..engine import Engine, make_invariant_params
    if _arun_engine is None:
        _arun_engine = Engine()
    e = _arun_engine
    e.set_invariant_params(make_invariant_params(kind=0, point_dim=dim))
    k = len(corr)
    e.plan(dim, [len(pts1)], [len(pts2)], m=[max(k, 1)])
    sel = np.zeros((max(k, 1), 2), dtype=np.int32)
    sel[:k] = corr
    T, status = e.arun(np.ascontiguousarray(pts1).reshape(-1), np.ascontiguousarray(pts2).reshape(-1),
                       sel.reshape(-1), np.array([k], dtype=np.int32))
    T = T[0].cpu().numpy()
    if dim == 3:
        return T
    out = np.eye(3)
    out[:2, :2] = T[:2, :2]
    out[:2, 2] = T[:2, 3]
    return out
