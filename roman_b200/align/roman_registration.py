"""
ROMANRegistration — same interface as reference roman/align/roman_registration.py:16-114, plus
`register_batch`, the batched entry point (many submap pairs per launch) the GPU path is built for.
"""
from __future__ import annotations

from dataclasses import dataclass
from enum import Enum
from typing import List, Optional, Sequence, Tuple

import numpy as np

from .. import clipperpy
from .object_registration import ObjectRegistration


class FusionMethod(Enum):
    GEOMETRIC_MEAN = clipperpy.invariants.ROMAN.GEOMETRIC_MEAN
    ARITHMETIC_MEAN = clipperpy.invariants.ROMAN.ARITHMETIC_MEAN
    PRODUCT = clipperpy.invariants.ROMAN.PRODUCT


@dataclass
class ROMANParams:
    """reference roman_registration.py:21-36"""
    point_dim: int = 3
    fusion_method = FusionMethod.GEOMETRIC_MEAN

    sigma: float = 0.4
    epsilon: float = 0.6
    mindist: float = 0.2

    gravity: bool = False
    volume: bool = False
    pca: bool = False
    extent: bool = False
    semantics_dim: int = 0
    gravity_unc_ang_rad: float = 0.0872665

    cos_min: float = 0.85
    cos_max: float = 1.0
    epsilon_shape: float = None


class ROMANRegistration(ObjectRegistration):
    """
    Maps ROMANParams onto the ROMAN invariant and packs each object into the feature row
        [centre(dim) | linearity, planarity, scattering | volume | sorted extent(3) | descriptor(D)]
    keeping only the enabled groups, in that order (reference roman_registration.py:38-108).
    """

    def __init__(self, params: ROMANParams):
        super().__init__(dim=params.point_dim)
        self.pca, self.volume, self.extent = bool(params.pca), bool(params.volume), bool(params.extent)
        self.semantics = params.semantics_dim > 0
        n_ratio = 3 * self.pca + 1 * self.volume + 3 * self.extent

        ip = clipperpy.invariants.ROMANParams()
        ip.point_dim, ip.ratio_feature_dim, ip.cos_feature_dim = params.point_dim, n_ratio, params.semantics_dim
        ip.sigma, ip.epsilon, ip.mindist = params.sigma, params.epsilon, params.mindist
        ip.distance_weight = ip.ratio_weight = ip.cosine_weight = 1.0
        ip.ratio_epsilon = np.full(n_ratio, 0.0 if params.epsilon_shape is None else params.epsilon_shape)
        ip.cosine_min, ip.cosine_max = params.cos_min, params.cos_max
        ip.gravity_guided = bool(params.gravity)
        ip.gravity_unc_ang_rad = params.gravity_unc_ang_rad if params.gravity else 0.0
        ip.drift_aware = False
        # params.fusion_method is deliberately not forwarded: the reference never does
        # (roman_registration.py:55-78), so the invariant always fuses with GEOMETRIC_MEAN.
        self.iparams = ip
        self._batch_engine = None

    def _setup_clipper(self):
        return clipperpy.CLIPPERPairwiseAndSingle(clipperpy.invariants.ROMAN(self.iparams), clipperpy.Params())

    def _object_to_clipper_list(self, object):
        row = [float(v) for v in np.asarray(object.center).reshape(-1)[:self.dim]]
        if self.pca:
            row.extend((object.linearity, object.planarity, object.scattering))
        if self.volume:
            row.append(object.volume)
        if self.extent:
            row.extend(sorted(object.extent))
        if self.semantics:
            row.extend(np.asarray(object.semantic_descriptor, dtype=np.float64).ravel().tolist())
        return row

    def _check_clipper_arrays(self, map1_cl, map2_cl):
        assert map1_cl.shape[1] == map2_cl.shape[1]

    def rows(self, objects: List) -> np.ndarray:
        """(n, d) float64 feature rows of an object list: the same numbers `_object_to_clipper_list` produces,
        assembled column-group-wise instead of through per-object Python lists (the reference's packing costs
        more than the GPU registration itself)."""
        n, d = len(objects), self.row_length()
        if n == 0:
            return np.zeros((0, d))
        cols = [np.array([np.asarray(o.center, dtype=np.float64).reshape(-1)[:self.dim] for o in objects]).reshape(n, self.dim)]
        if self.pca:
            cols.append(np.array([(o.linearity, o.planarity, o.scattering) for o in objects], dtype=np.float64).reshape(n, 3))
        if self.volume:
            cols.append(np.array([o.volume for o in objects], dtype=np.float64).reshape(n, 1))
        if self.extent:
            cols.append(np.sort(np.array([np.asarray(o.extent, dtype=np.float64).reshape(-1) for o in objects]).reshape(n, 3), axis=1))
        if self.semantics:
            cols.append(np.stack([np.asarray(o.semantic_descriptor, dtype=np.float64).ravel() for o in objects]).reshape(n, -1))
        out = np.concatenate(cols, axis=1)
        assert out.shape[1] == d, "object feature row does not match the invariant's dimensions"
        return out

    def _clipper_score_all_to_all(self, clipper, map1: List, map2: List):
        rows1 = self.rows(map1)
        rows2 = self.rows(map2)
        self._check_clipper_arrays(rows1, rows2)
        A_init = clipperpy.utils.create_all_to_all(len(map1), len(map2))
        clipper.score_pairwise_and_single_consistency(rows1.T, rows2.T, A_init)
        return clipper, A_init

    # ------------------------------------------------------------------ batched entry point (new)
    def row_length(self) -> int:
        ip = self.iparams
        return int(ip.point_dim + ip.ratio_feature_dim + ip.cos_feature_dim)

    def pack(self, maps: Sequence[List]) -> Tuple[np.ndarray, np.ndarray]:
        """Pack object lists into one (sum n, d) float64 row table + the per-map counts."""
        d = self.row_length()
        rows = [self.rows(m) for m in maps]
        table = np.concatenate(rows, axis=0) if rows else np.zeros((0, d))
        return np.ascontiguousarray(table), np.array([len(m) for m in maps], dtype=np.int32)

    def register_batch(self, pairs: Sequence[Tuple[List, List]], u0: Optional[Sequence[np.ndarray]] = None):
        """
        Register many (map1, map2) pairs in one launch: ObjectRegistration.register + T_align
        for each pair.  Returns (associations, T, status): a list of (k_b, 2) int32 arrays,
        a (B, dim+1, dim+1) float64 array (NaN rows where status != 0, as the NaN poses of
        submap_align.py:179-184) and a (B,) int32 status (0 ok, 1 insufficient associations).
        """
        from ..engine import Engine, default_solver_params
        B = len(pairs)
        if B == 0:
            return [], np.zeros((0, self.dim + 1, self.dim + 1)), np.zeros(0, dtype=np.int32)
        D1, n1 = self.pack([p[0] for p in pairs])
        D2, n2 = self.pack([p[1] for p in pairs])
        d = D1.shape[1]
        if self._batch_engine is None:
            self._batch_engine = Engine()
        e = self._batch_engine
        e.set_invariant_params(clipperpy.invariants.ROMAN(self.iparams)._rb_params())
        e.set_solver_params(default_solver_params())
        e.plan(d, n1, n2)
        m = n1.astype(np.int64) * n2.astype(np.int64)
        if u0 is None:
            u0cat = np.random.default_rng().uniform(0.0, 1.0, int(m.sum()))
        else:
            u0cat = np.concatenate([np.asarray(u, dtype=np.float64).reshape(-1) for u in u0]) if m.sum() else np.zeros(0)
            assert u0cat.size == m.sum()
        res = e.align_batch(D1.reshape(-1), D2.reshape(-1), u0cat)
        n_sel = res.n_sel.cpu().numpy()
        sel = res.sel.cpu().numpy()
        T = res.T.cpu().numpy()
        status = res.status.cpu().numpy().copy()
        self.last_scores = res.score.cpu().numpy().copy()      # objective value F of every pair of the last batch
        assoc = [sel[int(e.offsets[b]):int(e.offsets[b]) + int(n_sel[b])].copy() for b in range(B)]
        if self.dim == 2:
            T3 = np.tile(np.eye(3), (B, 1, 1))
            T3[:, :2, :2] = T[:, :2, :2]
            T3[:, :2, 2] = T[:, :2, 3]
            T3[status != 0] = np.nan
            T = T3
        return assoc, T, status

    def register_multistart(self, map1: List, map2: List, u0s: Sequence[np.ndarray]):
        """
        K solver restarts on ONE pair (CLIPPER is a local method and the reference
        draws a single random start per pair, clipper/src/utils.cpp:22-29).  The K problems share ONE affinity
        matrix (rb_plan_shared): K1 scores and stores the pair once, the K solves run side by side on it with their own
        start vectors.  Returns a dict with the best-scoring start's `associations` (k, 2), `T`, `status`, its index
        `best`, and the `scores` of all starts.
        """
        from ..engine import Engine, default_solver_params
        K = len(u0s)
        if K == 0 or len(map1) == 0 or len(map2) == 0:
            raise ValueError("register_multistart needs at least one start vector and two non-empty maps")
        D1, D2 = self.rows(map1), self.rows(map2)
        n1, n2 = len(map1), len(map2)
        if self._batch_engine is None:
            self._batch_engine = Engine()
        e = self._batch_engine
        e.set_invariant_params(clipperpy.invariants.ROMAN(self.iparams)._rb_params())
        e.set_solver_params(default_solver_params())
        zeros = np.zeros(K, dtype=np.int64)
        e.plan(D1.shape[1], [n1] * K, [n2] * K, obj1_offsets=zeros, obj2_offsets=zeros, total_obj1=n1, total_obj2=n2,
               share=np.zeros(K, dtype=np.int32))
        u0cat = np.concatenate([np.asarray(u, dtype=np.float64).reshape(-1) for u in u0s])
        assert u0cat.size == K * n1 * n2
        out = e.align_batch_host(np.ascontiguousarray(D1), np.ascontiguousarray(D2), u0cat)
        best = int(np.argmax(out["score"]))
        o = int(e.offsets[best])
        T = out["T"][best].copy()
        if out["status"][best] != 0:
            T[:] = np.nan
        return dict(associations=out["sel"][o:o + int(out["n_sel"][best])].copy(), T=T, status=int(out["status"][best]),
                    best=best, scores=out["score"].copy())

