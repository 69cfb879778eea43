"""
oracle — CPU restatement of the reference's CLIPPER registration path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.  The product (roman_b200/) never does.

`COracle` is a ctypes wrapper over oracle/_build/liboracle.so (built from clipper_oracle.c by
oracle/Makefile); `clipper_np` is the independent numpy restatement.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from . import clipper_np  # noqa: F401

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")
ORC_MAX_RATIO = 8


class OrcInvariantParams(C.Structure):
    _fields_ = [
        ("kind", C.c_int32), ("point_dim", C.c_int32), ("ratio_feature_dim", C.c_int32),
        ("cos_feature_dim", C.c_int32), ("fusion_method", C.c_int32), ("gravity_guided", C.c_int32),
        ("drift_aware", C.c_int32), ("drift_scale_sigma", C.c_int32),
        ("sigma", C.c_double), ("epsilon", C.c_double), ("mindist", C.c_double),
        ("distance_weight", C.c_double), ("ratio_weight", C.c_double), ("cosine_weight", C.c_double),
        ("ratio_epsilon", C.c_double * ORC_MAX_RATIO),
        ("cosine_min", C.c_double), ("cosine_max", C.c_double),
        ("gravity_unc_ang_rad", C.c_double), ("drift_scale", C.c_double),
    ]


class OrcSolverParams(C.Structure):
    _fields_ = [
        ("tol_u", C.c_double), ("tol_F", C.c_double), ("tol_Fop", C.c_double),
        ("maxiniters", C.c_int32), ("maxoliters", C.c_int32),
        ("beta", C.c_double), ("maxlsiters", C.c_int32),
        ("eps", C.c_double), ("affinityeps", C.c_double),
        ("rescale_u0", C.c_int32), ("rounding", C.c_int32),
    ]


class OrcCsc(C.Structure):
    _fields_ = [("n", C.c_int64), ("nnz", C.c_int64), ("colptr", C.POINTER(C.c_int64)),
                ("rowidx", C.POINTER(C.c_int32)), ("val", C.POINTER(C.c_double))]


def build(force: bool = False) -> str:
    """Compile oracle/clipper_oracle.c (building the checker is not using it)."""
    src = os.path.join(_HERE, "clipper_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


def _dptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _iptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


def to_c_invariant(p: clipper_np.InvariantParams) -> OrcInvariantParams:
    s = OrcInvariantParams()
    s.kind, s.point_dim = int(p.kind), int(p.point_dim)
    s.ratio_feature_dim, s.cos_feature_dim = int(p.ratio_feature_dim), int(p.cos_feature_dim)
    s.fusion_method = int(p.fusion_method)
    s.gravity_guided, s.drift_aware, s.drift_scale_sigma = int(p.gravity_guided), int(p.drift_aware), int(p.drift_scale_sigma)
    s.sigma, s.epsilon, s.mindist = p.sigma, p.epsilon, p.mindist
    s.distance_weight, s.ratio_weight, s.cosine_weight = p.distance_weight, p.ratio_weight, p.cosine_weight
    re = np.zeros(ORC_MAX_RATIO)
    r = np.asarray(p.ratio_epsilon, dtype=np.float64).reshape(-1)
    assert p.ratio_feature_dim <= ORC_MAX_RATIO
    re[: len(r)] = r
    for i in range(ORC_MAX_RATIO):
        s.ratio_epsilon[i] = re[i]
    s.cosine_min, s.cosine_max = p.cosine_min, p.cosine_max
    s.gravity_unc_ang_rad, s.drift_scale = p.gravity_unc_ang_rad, p.drift_scale
    return s


def to_c_solver(p: clipper_np.SolverParams) -> OrcSolverParams:
    s = OrcSolverParams()
    s.tol_u, s.tol_F, s.tol_Fop = p.tol_u, p.tol_F, p.tol_Fop
    s.maxiniters, s.maxoliters = int(p.maxiniters), int(p.maxoliters)
    s.beta, s.maxlsiters = p.beta, int(p.maxlsiters)
    s.eps, s.affinityeps = p.eps, p.affinityeps
    s.rescale_u0, s.rounding = int(p.rescale_u0), int(p.rounding)
    return s


class COracle:
    """ctypes front-end of the C restatement."""

    def __init__(self):
        self.lib = C.CDLL(build())
        L = self.lib
        L.orc_score.restype = C.c_int
        L.orc_solve.restype = C.c_int
        L.orc_arun.restype = C.c_int
        L.orc_align_pair.restype = C.c_int
        L.orc_align_batch.restype = C.c_int
        L.orc_k_largest.restype = C.c_int
        L.orc_max_threads.restype = C.c_int
        L.orc_csc_from_dense_upper.restype = C.c_int

    def max_threads(self) -> int:
        return int(self.lib.orc_max_threads())

    def score(self, ip, affinityeps, D1, D2, A=None, parallelize=True):
        """Returns (dense strict-upper M (m,m), single (m,), A (m,2), nnz)."""
        D1 = np.ascontiguousarray(D1, dtype=np.float64)
        D2 = np.ascontiguousarray(D2, dtype=np.float64)
        n1, d = D1.shape
        n2 = D2.shape[0]
        if A is None or np.asarray(A).size == 0:
            A = clipper_np.create_all_to_all(n1, n2)
        A = np.ascontiguousarray(A, dtype=np.int32)
        m = A.shape[0]
        dense = np.zeros((m, m))
        single = np.zeros(m)
        csc = OrcCsc()
        cip = to_c_invariant(ip)
        rc = self.lib.orc_score(C.byref(cip), C.c_double(affinityeps), C.c_int(d), C.c_int(n1), _dptr(D1),
                                C.c_int(n2), _dptr(D2), C.c_int64(m), _iptr(A), C.c_int(int(parallelize)),
                                C.byref(csc), _dptr(dense), _dptr(single))
        assert rc == 0
        nnz = int(csc.nnz)
        self.lib.orc_csc_free(C.byref(csc))
        return dense, single, A, nnz

    def solve(self, sp, Mupper, u0, Cupper=None):
        """Solve on a dense strict-upper M (and optional explicit C); Eigen accumulation order."""
        Mupper = np.ascontiguousarray(Mupper, dtype=np.float64)
        m = Mupper.shape[0]
        mc, cc = OrcCsc(), OrcCsc()
        self.lib.orc_csc_from_dense_upper(C.c_int64(m), _dptr(Mupper), C.byref(mc))
        pattern = Cupper is None
        if not pattern:
            Cupper = np.ascontiguousarray(Cupper, dtype=np.float64)
            self.lib.orc_csc_from_dense_upper(C.c_int64(m), _dptr(Cupper), C.byref(cc))
        u0 = np.ascontiguousarray(u0, dtype=np.float64)
        u = np.zeros(m)
        nodes = np.zeros(max(m, 1), dtype=np.int32)
        nn, ifinal, passes = C.c_int32(0), C.c_int32(0), C.c_int32(0)
        score = C.c_double(0)
        csp = to_c_solver(sp)
        rc = self.lib.orc_solve(C.byref(csp), C.byref(mc), C.byref(mc if pattern else cc), C.c_int(int(pattern)),
                                _dptr(u0), _dptr(u), _iptr(nodes), C.byref(nn), C.byref(score), C.byref(ifinal),
                                C.byref(passes))
        self.lib.orc_csc_free(C.byref(mc))
        if not pattern:
            self.lib.orc_csc_free(C.byref(cc))
        if rc == -2:
            raise NotImplementedError("exact DSD rounding is out of scope")
        assert rc == 0
        return dict(nodes=[int(x) for x in nodes[: nn.value]], u=u, score=score.value, ifinal=ifinal.value,
                    passes=passes.value)

    def k_largest(self, x, k):
        x = np.ascontiguousarray(x, dtype=np.float64)
        out = np.zeros(max(len(x), 1), dtype=np.int32)
        n = self.lib.orc_k_largest(_dptr(x), C.c_int64(len(x)), C.c_int(int(k)), _iptr(out))
        return [int(v) for v in out[:n]]

    def arun(self, pts1, pts2):
        pts1 = np.ascontiguousarray(pts1, dtype=np.float64)
        pts2 = np.ascontiguousarray(pts2, dtype=np.float64)
        T = np.zeros((4, 4))
        rc = self.lib.orc_arun(C.c_int(pts1.shape[0]), _dptr(pts1), _dptr(pts2), _dptr(T))
        return rc, T

    def align_pair(self, ip, sp, D1, D2, u0, A=None, parallelize=True):
        D1 = np.ascontiguousarray(D1, dtype=np.float64)
        D2 = np.ascontiguousarray(D2, dtype=np.float64)
        n1, d = D1.shape
        n2 = D2.shape[0]
        if A is not None:
            A = np.ascontiguousarray(A, dtype=np.int32)
            m = A.shape[0]
        else:
            m = n1 * n2
        u0 = np.ascontiguousarray(u0, dtype=np.float64)
        sel = np.zeros((max(m, 1), 2), dtype=np.int32)
        T = np.zeros((4, 4))
        nsel, passes = C.c_int32(0), C.c_int32(0)
        score = C.c_double(0)
        nnz = C.c_int64(0)
        cip, csp = to_c_invariant(ip), to_c_solver(sp)
        st = self.lib.orc_align_pair(C.byref(cip), C.byref(csp), C.c_int(d), C.c_int(n1), _dptr(D1), C.c_int(n2),
                                     _dptr(D2), C.c_int64(m), _iptr(A) if A is not None else None, _dptr(u0),
                                     C.c_int(int(parallelize)), _iptr(sel), C.byref(nsel), _dptr(T),
                                     C.byref(score), C.byref(passes), C.byref(nnz))
        assert st in (0, 1), st
        return dict(sel=sel[: nsel.value].copy(), T=T, score=score.value, passes=passes.value, status=st,
                    nnz=nnz.value)

    def align_batch(self, ip, sp, D1, D2, u0, threads=0):
        """D1 (B,n1,d), D2 (B,n2,d), u0 (B,m): one pair per thread.  Returns dict + threads used."""
        D1 = np.ascontiguousarray(D1, dtype=np.float64)
        D2 = np.ascontiguousarray(D2, dtype=np.float64)
        B, n1, d = D1.shape
        n2 = D2.shape[1]
        m = n1 * n2
        u0 = np.ascontiguousarray(u0, dtype=np.float64)
        sel = np.zeros((B, m, 2), dtype=np.int32)
        nsel = np.zeros(B, dtype=np.int32)
        T = np.zeros((B, 4, 4))
        score = np.zeros(B)
        passes = np.zeros(B, dtype=np.int32)
        status = np.zeros(B, dtype=np.int32)
        cip, csp = to_c_invariant(ip), to_c_solver(sp)
        used = self.lib.orc_align_batch(C.byref(cip), C.byref(csp), C.c_int(B), C.c_int(d), C.c_int(n1), _dptr(D1),
                                        C.c_int(n2), _dptr(D2), _dptr(u0), C.c_int(int(threads)), _iptr(sel),
                                        _iptr(nsel), _dptr(T), _dptr(score), _iptr(passes), _iptr(status))
        return dict(sel=[sel[b, : nsel[b]].copy() for b in range(B)], T=T, score=score, passes=passes,
                    status=status, threads=int(used))
