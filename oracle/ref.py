"""
oracle.ref — ctypes front-end of oracle/_ref/libclipper_ref.so: the REFERENCE's own CLIPPER C++ sources
(dependencies/clipper/src/{clipper,utils,dsd}.cpp, src/invariants/{roman,euclidean_distance}.cpp), compiled unchanged
against the Eigen stand-in of oracle/eigen_standin by `make -C oracle ref` (needs /root/reference; the built library
travels to the GPU box).  TEST INFRASTRUCTURE: it pins the C / numpy restatements (tests/test_ref_pins_oracle.py) and
is the timed CPU baseline of bench.py (`cpu_baseline.kind == "reference"`).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from . import OrcInvariantParams, OrcSolverParams, to_c_invariant, to_c_solver, clipper_np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libclipper_ref.so")
REFERENCE_SRC = "/root/reference/dependencies/clipper"


def build() -> bool:
    """Compile the reference sources if they are present (this container); on the GPU box only the prebuilt file is used."""
    if os.path.isdir(os.path.join(REFERENCE_SRC, "src")):
        subprocess.check_call(["make", "-C", _HERE, "-s", "ref"])
    return os.path.exists(LIB_PATH)


def available() -> bool:
    return os.path.exists(LIB_PATH)


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


class RefClipper:
    def __init__(self):
        if not available():
            raise ImportError(f"{LIB_PATH} not built (run `make -C oracle ref` where /root/reference exists)")
        self.lib = L = C.CDLL(LIB_PATH)
        for name in ("ref_score", "ref_align_pair", "ref_solve_dense", "ref_k_largest", "ref_dsd"):
            getattr(L, name).restype = C.c_int
        for name in ("ref_single_similarity", "ref_pairwise_similarity", "ref_fusion"):
            getattr(L, name).restype = C.c_double

    @staticmethod
    def _kind(ip, pairwise_only):
        c = to_c_invariant(ip)
        if pairwise_only and c.kind == 0:
            c.kind = 2      # the ROMAN invariant through CLIPPER::scorePairwiseConsistency (no single scores, no fusion)
        return c

    def score(self, ip, sp, D1, D2, A=None, parallelize=True, pairwise_only=False):
        """(M + I, C + I) dense (m, m) as CLIPPER::getAffinityMatrix / getConstraintMatrix return them."""
        D1 = np.ascontiguousarray(D1, dtype=np.float64)
        D2 = np.ascontiguousarray(D2, dtype=np.float64)
        n1, d = D1.shape
        n2 = D2.shape[0]
        m = n1 * n2 if A is None else len(A)
        A = None if A is None else np.ascontiguousarray(A, dtype=np.int32)
        M, Cm = np.zeros((m, m)), np.zeros((m, m))
        cip, csp = self._kind(ip, pairwise_only), to_c_solver(sp)
        rc = self.lib.ref_score(C.byref(cip), C.byref(csp), C.c_int(d), C.c_int(n1), _dp(D1), C.c_int(n2), _dp(D2),
                                C.c_int64(m), _ip(A) if A is not None else None, C.c_int(int(parallelize)), _dp(M), _dp(Cm))
        assert rc == 0
        return M, Cm

    def align_pair(self, ip, sp, D1, D2, u0, A=None, parallelize=True, pairwise_only=False, want_T=True):
        D1 = np.ascontiguousarray(D1, dtype=np.float64)
        D2 = np.ascontiguousarray(D2, dtype=np.float64)
        n1, d = D1.shape
        n2 = D2.shape[0]
        m = n1 * n2 if A is None else len(A)
        A = None if A is None else np.ascontiguousarray(A, dtype=np.int32)
        u0 = np.ascontiguousarray(u0, dtype=np.float64)
        sel = np.zeros((max(m, 1), 2), dtype=np.int32)
        nodes = np.zeros(max(m, 1), dtype=np.int32)
        u = np.zeros(max(m, 1))
        T = np.zeros((4, 4))
        nsel, ifinal = C.c_int32(0), C.c_int32(0)
        score = C.c_double(0)
        cip, csp = self._kind(ip, pairwise_only), to_c_solver(sp)
        st = self.lib.ref_align_pair(C.byref(cip), C.byref(csp), C.c_int(d), C.c_int(n1), _dp(D1), C.c_int(n2), _dp(D2),
                                     C.c_int64(m), _ip(A) if A is not None else None, _dp(u0), C.c_int(int(parallelize)),
                                     _ip(sel), C.byref(nsel), _dp(T) if want_T else None, C.byref(score), C.byref(ifinal),
                                     _dp(u), _ip(nodes))
        assert st in (0, 1), st
        k = nsel.value
        return dict(sel=sel[:k].copy(), T=T, score=score.value, status=st, ifinal=ifinal.value, u=u[:m].copy(),
                    nodes=[int(x) for x in nodes[:k]])

    def solve_dense(self, sp, M, u0, Cm=None):
        """setMatrixData(M, C) + solve(u0); M (and C) dense symmetric incl. or excl. the diagonal (it is zeroed)."""
        M = np.ascontiguousarray(M, dtype=np.float64)
        m = M.shape[0]
        Cm = None if Cm is None else np.ascontiguousarray(Cm, dtype=np.float64)
        u0 = np.ascontiguousarray(u0, dtype=np.float64)
        nodes = np.zeros(max(m, 1), dtype=np.int32)
        u = np.zeros(max(m, 1))
        nn, ifinal = C.c_int32(0), C.c_int32(0)
        score = C.c_double(0)
        csp = to_c_solver(sp)
        rc = self.lib.ref_solve_dense(C.byref(csp), C.c_int64(m), _dp(M), _dp(Cm) if Cm is not None else None, _dp(u0),
                                      _ip(nodes), C.byref(nn), C.byref(score), C.byref(ifinal), _dp(u))
        assert rc == 0
        return dict(nodes=[int(x) for x in nodes[:nn.value]], u=u[:m].copy(), score=score.value, ifinal=ifinal.value)

    def single_similarity(self, ip, a, b):
        a = np.ascontiguousarray(a, dtype=np.float64)
        b = np.ascontiguousarray(b, dtype=np.float64)
        c = to_c_invariant(ip)
        return float(self.lib.ref_single_similarity(C.byref(c), C.c_int(len(a)), _dp(a), _dp(b)))

    def pairwise_similarity(self, ip, ai, aj, bi, bj):
        v = [np.ascontiguousarray(x, dtype=np.float64) for x in (ai, aj, bi, bj)]
        c = to_c_invariant(ip)
        return float(self.lib.ref_pairwise_similarity(C.byref(c), C.c_int(len(v[0])), *[_dp(x) for x in v]))

    def fusion(self, ip, pair_ij, si, sj):
        c = to_c_invariant(ip)
        return float(self.lib.ref_fusion(C.byref(c), C.c_double(pair_ij), C.c_double(si), C.c_double(sj)))

    def k_largest(self, x, k):
        x = np.ascontiguousarray(x, dtype=np.float64)
        out = np.zeros(max(int(k), 1), dtype=np.int32)
        n = self.lib.ref_k_largest(_dp(x), C.c_int64(len(x)), C.c_int(int(k)), _ip(out))
        return [int(v) for v in out[:n]]

    def k2ij(self, k, n):
        i, j = C.c_uint64(0), C.c_uint64(0)
        self.lib.ref_k2ij(C.c_uint64(k), C.c_uint64(n), C.byref(i), C.byref(j))
        return int(i.value), int(j.value)

    def dsd(self, W, S=None):
        W = np.ascontiguousarray(W, dtype=np.float64)
        n = W.shape[0]
        S = np.arange(n, dtype=np.int32) if S is None else np.ascontiguousarray(S, dtype=np.int32)
        out = np.zeros(n, dtype=np.int32)
        k = self.lib.ref_dsd(C.c_int64(n), _dp(W), C.c_int32(len(S)), _ip(S), _ip(out))
        assert k >= 0
        return [int(v) for v in out[:k]]
