"""Shared test helpers: build matching parameter objects for the oracle and the CUDA path."""
import numpy as np

from oracle import clipper_np as cn
from roman_b200 import synth
from roman_b200.engine import make_invariant_params, method_invariant_kwargs, default_solver_params


def oracle_invariant(**kw) -> cn.InvariantParams:
    kw = dict(kw)
    re = kw.pop("ratio_epsilon", None)
    p = cn.InvariantParams(**kw)
    p.ratio_epsilon = np.zeros(p.ratio_feature_dim) if re is None else np.asarray(re, dtype=np.float64)
    return p


def params_for_method(method="roman", **over):
    kw = method_invariant_kwargs(method, **over)
    return oracle_invariant(**kw), make_invariant_params(**kw)


def rel_fro(T, Tref):
    return float(np.linalg.norm(T - Tref) / np.linalg.norm(Tref))


def same_set(a, b):
    sa = sorted(map(tuple, np.asarray(a).reshape(-1, 2).tolist()))
    sb = sorted(map(tuple, np.asarray(b).reshape(-1, 2).tolist()))
    return sa == sb
