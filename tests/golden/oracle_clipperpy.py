"""
TEST INFRASTRUCTURE — a CPU stand-in for the reference's pybind11 module `clipperpy`, backed by the
C oracle (oracle/clipper_oracle.c).  Its only user is tests/golden/make_reference_py_golden.py,
which imports the reference's own Python registration code from /root/reference (the compiled
`clipperpy` cannot be built here: no Eigen3) and needs something behind the ~15 names that code
touches (clipper/bindings/python/py_clipper.cpp:27-332).  Nothing under roman_b200/ imports this.

The reference calls `clipper.solve()` without a start vector and the C++ side draws one at random
(clipper/src/utils.cpp:22-29).  Here the start vector comes from `next_u0(m)`, which the generator
script seeds, and every vector handed out is appended to `U0_LOG` so that it can be stored next to
the results.
"""
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import oracle                      # noqa: E402
from oracle import clipper_np      # noqa: E402

_orc = None
U0_LOG = []
_rng = np.random.default_rng(0)


def seed_u0(seed):
    global _rng
    _rng = np.random.default_rng(seed)
    U0_LOG.clear()


def next_u0(m):
    u0 = _rng.uniform(0.0, 1.0, int(m))
    U0_LOG.append(u0)
    return u0


def _oracle():
    global _orc
    if _orc is None:
        _orc = oracle.COracle()
    return _orc


# ---------------------------------------------------------------- clipperpy.invariants
class PairwiseInvariant:
    pass


class ROMANParams:
    def __init__(self):
        d = clipper_np.InvariantParams()
        for k, v in d.__dict__.items():
            if k != "kind":
                setattr(self, k, v)


class ROMAN(PairwiseInvariant):
    GEOMETRIC_MEAN, ARITHMETIC_MEAN, PRODUCT = 0, 1, 2

    def __init__(self, params):
        self.params = params

    def _np(self):
        p = self.params
        return clipper_np.InvariantParams(
            kind=0, point_dim=int(p.point_dim), ratio_feature_dim=int(p.ratio_feature_dim),
            cos_feature_dim=int(p.cos_feature_dim), fusion_method=int(p.fusion_method), sigma=float(p.sigma),
            epsilon=float(p.epsilon), mindist=float(p.mindist), distance_weight=float(p.distance_weight),
            ratio_weight=float(p.ratio_weight), cosine_weight=float(p.cosine_weight),
            ratio_epsilon=np.asarray(p.ratio_epsilon, dtype=np.float64), cosine_min=float(p.cosine_min),
            cosine_max=float(p.cosine_max), gravity_guided=bool(p.gravity_guided),
            gravity_unc_ang_rad=float(p.gravity_unc_ang_rad), drift_aware=bool(p.drift_aware),
            drift_scale_sigma=bool(p.drift_scale_sigma), drift_scale=float(p.drift_scale))


class EuclideanDistanceParams:
    def __init__(self):
        self.sigma, self.epsilon, self.mindist = 0.01, 0.06, 0.0


class EuclideanDistance(PairwiseInvariant):
    def __init__(self, params):
        self.params = params

    def _np(self):
        p = self.params
        return clipper_np.InvariantParams(kind=1, sigma=float(p.sigma), epsilon=float(p.epsilon),
                                          mindist=float(p.mindist))


# ---------------------------------------------------------------- clipperpy
class Rounding:
    NONZERO, DSD, DSD_HEU = 0, 1, 2


class Params:
    def __init__(self):
        for k, v in clipper_np.SolverParams().__dict__.items():
            setattr(self, k, v)

    def _np(self):
        return clipper_np.SolverParams(**{k: getattr(self, k) for k in clipper_np.SolverParams().__dict__})


class Solution:
    pass


class CLIPPER:
    def __init__(self, invariant, params):
        self.invariant, self.params = invariant, params
        self.A = np.zeros((0, 2), dtype=np.int32)
        self.Mu = self.Cu = None
        self.soln = None

    def _score(self, D1, D2, A):
        # the binding receives (d, n) matrices; one column = one object
        D1, D2 = np.asarray(D1, dtype=np.float64).T, np.asarray(D2, dtype=np.float64).T
        Mu, _, A, _ = _oracle().score(self.invariant._np(), self.params.affinityeps, D1, D2, A)
        self.A, self.Mu, self.Cu = A, Mu, None

    def score_pairwise_consistency(self, D1, D2, A):
        self._score(D1, D2, A)

    def solve(self, u0=None):
        m = self.Mu.shape[0]
        if u0 is None or np.asarray(u0).size == 0:
            u0 = next_u0(m)
        r = _oracle().solve(self.params._np(), self.Mu, u0, self.Cu)
        s = Solution()
        s.nodes, s.u, s.score, s.ifinal, s.u0, s.t = list(r["nodes"]), r["u"], r["score"], r["ifinal"], u0, 0.0
        self.soln = s

    def get_solution(self):
        return self.soln

    def get_initial_associations(self):
        return self.A

    def get_selected_associations(self):
        return self.A[np.asarray(self.soln.nodes, dtype=np.int64)].reshape(-1, 2)

    def get_affinity_matrix(self):
        return self.Mu + self.Mu.T + np.eye(self.Mu.shape[0])

    def get_constraint_matrix(self):
        Cu = (self.Mu != 0).astype(np.float64) if self.Cu is None else self.Cu
        return Cu + Cu.T + np.eye(Cu.shape[0])

    def set_matrix_data(self, M, C):
        self.Mu = np.triu(np.asarray(M, dtype=np.float64), 1)
        self.Cu = np.triu(np.asarray(C, dtype=np.float64), 1)
        if self.A.shape[0] != self.Mu.shape[0]:
            self.A = np.zeros((self.Mu.shape[0], 2), dtype=np.int32)

    def set_parallelize(self, flag):
        pass


class CLIPPERPairwiseAndSingle(CLIPPER):
    def score_pairwise_and_single_consistency(self, D1, D2, A):
        self._score(D1, D2, A)


def install():
    """Register this module as `clipperpy` (+ .invariants, .utils) in sys.modules."""
    me = sys.modules[__name__]
    inv = types.ModuleType("clipperpy.invariants")
    for name in ("PairwiseInvariant", "ROMANParams", "ROMAN", "EuclideanDistanceParams", "EuclideanDistance"):
        setattr(inv, name, getattr(me, name))
    utils = types.ModuleType("clipperpy.utils")
    utils.create_all_to_all = clipper_np.create_all_to_all
    me.invariants, me.utils = inv, utils
    me.__version__ = "oracle"
    sys.modules["clipperpy"] = me
    sys.modules["clipperpy.invariants"] = inv
    sys.modules["clipperpy.utils"] = utils
    return me
