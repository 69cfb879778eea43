"""
clipperpy.invariants — parameter/invariant objects with the names and attributes of the reference's
pybind11 submodule (clipper/bindings/python/py_clipper.cpp:27-143).  They only carry parameters:
the scoring itself runs in the CUDA affinity kernel.
"""
from __future__ import annotations

import numpy as np

from .._lib import RB_INVARIANT_EUCLIDEAN, RB_INVARIANT_ROMAN
from ..engine import make_invariant_params


class Invariant:
    pass


class PairwiseInvariant(Invariant):
    """Bare base class; the reference uses an instance of it as a dummy invariant together with
    set_matrix_data (roman/align/object_registration.py:60).  Python subclasses that override
    __call__ cannot be evaluated by the CUDA affinity kernel and are rejected at scoring time."""

    def __call__(self, ai, aj, bi, bj):
        raise NotImplementedError("PairwiseInvariant.__call__ is abstract")


class PairwiseAndSingleInvariant(PairwiseInvariant):
    pass


class ROMANParams:
    """clipper::invariants::ROMAN::Params (roman.h:27-53; py_clipper.cpp:60-102)."""

    def __init__(self):
        self.point_dim = 3
        self.ratio_feature_dim = 0
        self.cos_feature_dim = 0
        self.fusion_method = ROMAN.GEOMETRIC_MEAN
        self.sigma = 0.01
        self.epsilon = 0.06
        self.mindist = 0.0
        self.distance_weight = 1.0
        self.ratio_weight = 1.0
        self.cosine_weight = 1.0
        self.ratio_epsilon = np.zeros(0)
        self.cosine_min = 0.85
        self.cosine_max = 0.95
        self.gravity_guided = False
        self.gravity_unc_ang_rad = 0.0
        self.drift_aware = False
        self.drift_scale_sigma = False
        self.drift_scale = 0.1

    def __repr__(self):
        return "<ROMANParams : " + " ".join(f"{k}={v}" for k, v in self.__dict__.items()) + ">"


class ROMAN(PairwiseAndSingleInvariant):
    """clipper::invariants::ROMAN (roman.h:19-104); enum values exported at class scope like
    py::enum_(...).export_values() does (py_clipper.cpp:54-58)."""
    GEOMETRIC_MEAN = 0
    ARITHMETIC_MEAN = 1
    PRODUCT = 2

    class SimilarityFusionMethod:
        GEOMETRIC_MEAN = 0
        ARITHMETIC_MEAN = 1
        PRODUCT = 2

    def __init__(self, params: ROMANParams):
        self.params = params

    def _rb_params(self):
        p = self.params
        re = np.asarray(p.ratio_epsilon, dtype=np.float64).reshape(-1)
        if p.ratio_feature_dim > 0 and len(re) != p.ratio_feature_dim:
            raise ValueError("ROMANParams.ratio_epsilon must have ratio_feature_dim entries")
        return make_invariant_params(
            kind=RB_INVARIANT_ROMAN, point_dim=int(p.point_dim), ratio_feature_dim=int(p.ratio_feature_dim),
            cos_feature_dim=int(p.cos_feature_dim), fusion_method=int(getattr(p.fusion_method, "value", p.fusion_method)),
            sigma=p.sigma, epsilon=p.epsilon, mindist=p.mindist, distance_weight=p.distance_weight,
            ratio_weight=p.ratio_weight, cosine_weight=p.cosine_weight, ratio_epsilon=re,
            cosine_min=p.cosine_min, cosine_max=p.cosine_max, gravity_guided=bool(p.gravity_guided),
            gravity_unc_ang_rad=p.gravity_unc_ang_rad, drift_aware=bool(p.drift_aware),
            drift_scale_sigma=bool(p.drift_scale_sigma), drift_scale=p.drift_scale)


class EuclideanDistanceParams:
    """clipper::invariants::EuclideanDistance::Params (euclidean_distance.h:24-29)."""

    def __init__(self):
        self.sigma = 0.01
        self.epsilon = 0.06
        self.mindist = 0.0

    def __repr__(self):
        return f"<EuclideanDistanceParams : sigma={self.sigma} epsilon={self.epsilon} mindist={self.mindist}>"


class EuclideanDistance(PairwiseInvariant):
    def __init__(self, params: EuclideanDistanceParams):
        self.params = params

    def _rb_params(self):
        p = self.params
        return make_invariant_params(kind=RB_INVARIANT_EUCLIDEAN, sigma=p.sigma, epsilon=p.epsilon, mindist=p.mindist)
