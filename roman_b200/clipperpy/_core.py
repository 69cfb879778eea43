"""CLIPPER / CLIPPERPairwiseAndSingle / Params / Solution with the reference's Python surface
(py_clipper.cpp:226-332), executing on the GPU through roman_b200.engine.Engine."""
from __future__ import annotations

import enum
import time

import numpy as np

from . import invariants as _inv
from .utils import create_all_to_all


class Rounding(enum.IntEnum):
    NONZERO = 0
    DSD = 1
    DSD_HEU = 2


NONZERO, DSD, DSD_HEU = Rounding.NONZERO, Rounding.DSD, Rounding.DSD_HEU  # export_values()


class Params:
    """clipper::Params (clipper.h:27-60)."""

    def __init__(self):
        self.tol_u = 1e-8
        self.tol_F = 1e-9
        self.tol_Fop = 1e-10
        self.maxiniters = 200
        self.maxoliters = 1000
        self.beta = 0.25
        self.maxlsiters = 99
        self.eps = 1e-9
        self.affinityeps = 1e-4
        self.rescale_u0 = True
        self.rounding = Rounding.DSD_HEU

    def __repr__(self):
        return "<CLIPPER Parameters>"


class Solution:
    """clipper::Solution (clipper.h:65-73)."""

    def __init__(self):
        self.t = 0.0
        self.ifinal = 0
        self.nodes = []
        self.u0 = np.zeros(0)
        self.u = np.zeros(0)
        self.score = 0.0

    def __repr__(self):
        return "<CLIPPER Solution>"


# ---- engine pool: the reference constructs a fresh CLIPPER per register() call
#      (roman_registration.py:82-86); handles (and their scratch) are recycled instead.
_pool = []


def _acquire_engine():
    from ..engine import Engine
    return _pool.pop() if _pool else Engine()


def _release_engine(e):
    if e is not None and len(_pool) < 8:
        _pool.append(e)


class CLIPPER:
    def __init__(self, invariant, params: Params):
        self._invariant = invariant
        self._params = params
        self._engine = None
        self._A = np.zeros((0, 2), dtype=np.int32)
        self._soln = Solution()
        self._m = 0
        self._scored = False

    def __repr__(self):
        return "<CLIPPER>"

    def __del__(self):
        try:
            _release_engine(self._engine)
        except Exception:
            pass

    # ------------------------------------------------------------------ helpers
    def _eng(self):
        if self._engine is None:
            self._engine = _acquire_engine()
        return self._engine

    def _solver_params(self):
        from ..engine import default_solver_params
        p, sp = self._params, default_solver_params()
        sp.tol_u, sp.tol_F, sp.tol_Fop = p.tol_u, p.tol_F, p.tol_Fop
        sp.maxiniters, sp.maxoliters = int(p.maxiniters), int(p.maxoliters)
        sp.beta, sp.maxlsiters = p.beta, int(p.maxlsiters)
        sp.eps, sp.affinityeps = p.eps, p.affinityeps
        sp.rescale_u0, sp.rounding = int(bool(p.rescale_u0)), int(p.rounding)
        return sp

    def _score(self, D1, D2, A, pairwise_only=False):
        inv = self._invariant
        if not hasattr(inv, "_rb_params"):
            raise NotImplementedError(
                "only the built-in ROMAN and EuclideanDistance invariants are evaluated on the GPU; "
                "Python-subclassed invariants (py_clipper.cpp trampolines) are not supported")
        # the reference receives (d, n) column-major matrices: one column = one object
        D1 = np.ascontiguousarray(np.asarray(D1, dtype=np.float64).T)
        D2 = np.ascontiguousarray(np.asarray(D2, dtype=np.float64).T)
        n1, d = D1.shape
        n2 = D2.shape[0]
        assert D2.shape[1] == d
        A = np.asarray(A) if A is not None else np.zeros((0, 2), dtype=np.int32)
        e = self._eng()
        ip = inv._rb_params()
        if pairwise_only and ip.kind == 0:
            # CLIPPER::scorePairwiseConsistency calls invariant->operator(), which for a PairwiseAndSingleInvariant is
            # pairwise_similarity alone (abstract.h:99-102): no single scores, no single-score gate, no fusion
            ip.kind = 2
        e.set_invariant_params(ip)
        e.set_solver_params(self._solver_params())
        if A.size == 0:  # clipper.cpp:32,81: empty A => all-to-all
            self._A = create_all_to_all(n1, n2)
            e.plan(d, [n1], [n2])
            e.affinity(D1.reshape(-1), D2.reshape(-1))
        else:
            self._A = np.ascontiguousarray(A, dtype=np.int32).reshape(-1, 2)
            e.plan(d, [n1], [n2], m=[len(self._A)])
            e.affinity(D1.reshape(-1), D2.reshape(-1), A=self._A.reshape(-1))
        self._m = len(self._A)
        self._scored = True

    # ------------------------------------------------------------------ reference API
    def score_pairwise_consistency(self, D1, D2, A=None):
        """CLIPPER::scorePairwiseConsistency (clipper.cpp:29-73)."""
        self._score(D1, D2, A, pairwise_only=True)

    def solve(self, u0=None):
        """CLIPPER::solve (clipper.cpp:140-149): u0 empty => U[0,1) random start (utils.cpp:22-29)."""
        if not self._scored:
            raise RuntimeError("solve() before scoring / set_matrix_data")
        m = self._m
        u0 = np.zeros(0) if u0 is None else np.asarray(u0, dtype=np.float64).reshape(-1)
        if u0.size == 0:
            u0 = np.random.default_rng().uniform(0.0, 1.0, m)
        assert u0.size == m
        e = self._eng()
        e.set_solver_params(self._solver_params())
        t0 = time.perf_counter()
        out = e.solve(u0, want_u=True)
        k = int(out["n_nodes"][0])
        s = Solution()
        s.nodes = out["nodes"][:k].cpu().numpy().astype(int).tolist()
        s.t = time.perf_counter() - t0
        s.u0 = u0.copy()
        s.u = out["u"][:m].cpu().numpy()
        s.score = float(out["score"][0])
        s.ifinal = int(out["ifinal"][0])
        self._soln = s

    def solve_as_maximum_clique(self, params=None):
        raise NotImplementedError("solveAsMaximumClique (PMC) is outside the roman.align hot path")

    def solve_as_msrc_sdr(self, params=None):
        raise NotImplementedError("solveAsMSRCSDR (SCS) is outside the roman.align hot path")

    def get_initial_associations(self):
        return self._A.copy()

    def get_selected_associations(self):
        """utils::selectInlierAssociations (utils.cpp:101-108): rows of A at the solution nodes."""
        if len(self._soln.nodes) == 0:
            return np.zeros((0, 2), dtype=np.int32)
        return self._A[np.asarray(self._soln.nodes, dtype=np.int64)]

    def get_solution(self):
        return self._soln

    def get_affinity_matrix(self):
        M, _ = self._eng().get_dense(0, want_M=True, want_C=False)
        return M.cpu().numpy()

    def get_constraint_matrix(self):
        _, C = self._eng().get_dense(0, want_M=False, want_C=True)
        return C.cpu().numpy()

    def set_matrix_data(self, M, C):
        """CLIPPER::setMatrixData (clipper.cpp:220-229): strict upper triangles of dense M, C."""
        M = np.ascontiguousarray(M, dtype=np.float64)
        C = np.ascontiguousarray(C, dtype=np.float64)
        m = M.shape[0]
        assert M.shape == (m, m) and C.shape == (m, m)
        e = self._eng()
        e.set_solver_params(self._solver_params())
        e.plan(1, [m], [1], m=[m])
        e.set_dense(0, M, C)
        self._m = m
        self._scored = True

    def set_parallelize(self, parallelize: bool):
        """No-op: the OpenMP switch of the reference (clipper.h:147) has no meaning here."""


class CLIPPERPairwiseAndSingle(CLIPPER):
    def __repr__(self):
        return "<CLIPPERPairwiseAndSingle>"

    def score_pairwise_and_single_consistency(self, D1, D2, A=None):
        """scorePairwiseAndSingleConsistency (clipper.cpp:77-136)."""
        if not isinstance(self._invariant, _inv.PairwiseAndSingleInvariant):
            raise TypeError("CLIPPERPairwiseAndSingle needs a PairwiseAndSingleInvariant (e.g. ROMAN)")
        self._score(D1, D2, A)
