"""
ctypes binding of libroman_b200.so (include/roman_b200.h).  There is NO CPU fallback: if the
CUDA library is missing this module raises at import time.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RB_LIB_PATH") or os.path.join(_HERE, "csrc", "libroman_b200.so")   # RB_LIB_PATH: dev builds of the same library
RB_MAX_RATIO_FEATURES = 8

RB_INVARIANT_ROMAN = 0
RB_INVARIANT_EUCLIDEAN = 1
RB_INVARIANT_ROMAN_PAIRWISE = 2
RB_PAIR_OK = 0
RB_PAIR_INSUFFICIENT = 1


class RbInvariantParams(C.Structure):
    """POD mirror of clipper::invariants::ROMAN::Params (reference roman.h:27-53)."""
    _fields_ = [
        ("kind", C.c_int32), ("point_dim", C.c_int32), ("ratio_feature_dim", C.c_int32),
        ("cos_feature_dim", C.c_int32), ("fusion_method", C.c_int32), ("gravity_guided", C.c_int32),
        ("drift_aware", C.c_int32), ("drift_scale_sigma", C.c_int32),
        ("sigma", C.c_double), ("epsilon", C.c_double), ("mindist", C.c_double),
        ("distance_weight", C.c_double), ("ratio_weight", C.c_double), ("cosine_weight", C.c_double),
        ("ratio_epsilon", C.c_double * RB_MAX_RATIO_FEATURES),
        ("cosine_min", C.c_double), ("cosine_max", C.c_double),
        ("gravity_unc_ang_rad", C.c_double), ("drift_scale", C.c_double),
    ]


class RbSolverParams(C.Structure):
    """POD mirror of clipper::Params (reference clipper.h:27-60)."""
    _fields_ = [
        ("tol_u", C.c_double), ("tol_F", C.c_double), ("tol_Fop", C.c_double),
        ("maxiniters", C.c_int32), ("maxoliters", C.c_int32),
        ("beta", C.c_double), ("maxlsiters", C.c_int32),
        ("eps", C.c_double), ("affinityeps", C.c_double),
        ("rescale_u0", C.c_int32), ("rounding", C.c_int32),
    ]


# name -> (restype, argtypes); every symbol include/roman_b200.h declares
_vp, _i32, _i64, _dp, _ip, _fp = C.c_void_p, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p, C.POINTER(C.c_float)
SIGNATURES = {
    "rb_default_invariant_params": (None, [C.POINTER(RbInvariantParams)]),
    "rb_default_solver_params": (None, [C.POINTER(RbSolverParams)]),
    "rb_create": (C.c_int, [C.POINTER(_vp), C.c_int]),
    "rb_destroy": (C.c_int, [_vp]),
    "rb_last_error": (C.c_char_p, [_vp]),
    "rb_set_invariant_params": (C.c_int, [_vp, C.POINTER(RbInvariantParams)]),
    "rb_set_solver_params": (C.c_int, [_vp, C.POINTER(RbSolverParams)]),
    "rb_plan": (C.c_int, [_vp, _i32, _i32, _ip, _ip, _ip]),
    "rb_plan_indexed": (C.c_int, [_vp, _i32, _i32, _ip, _ip, _ip, _vp, _vp, _i64, _i64]),
    "rb_plan_shared": (C.c_int, [_vp, _i32, _i32, _ip, _ip, _ip, _vp, _vp, _i64, _i64, _ip]),
    "rb_assoc_offset": (_i64, [_vp, _i32]),
    "rb_obj1_offset": (_i64, [_vp, _i32]),
    "rb_obj2_offset": (_i64, [_vp, _i32]),
    "rb_packed_bytes": (_i64, [_vp]),
    "rb_affinity": (C.c_int, [_vp, _dp, _dp, _ip, _vp]),
    "rb_solve": (C.c_int, [_vp, _dp, _ip, _ip, _ip, _dp, _dp, _ip, _ip, _vp]),
    "rb_dsd_host": (C.c_int, [_i64, _i32, _ip, _dp, _ip, _ip]),
    "rb_submap_similarity": (C.c_int, [_vp, C.c_int32, C.c_int32, C.c_int32, _vp, _vp, _dp, _dp, _dp, _vp]),
    "rb_arun": (C.c_int, [_vp, _dp, _dp, _ip, _ip, _dp, _ip, _vp]),
    "rb_align_batch": (C.c_int, [_vp, _dp, _dp, _ip, _dp, _ip, _ip, _dp, _ip, _dp, _ip, _vp]),
    "rb_get_dense": (C.c_int, [_vp, _i32, _dp, _dp, _vp]),
    "rb_set_dense": (C.c_int, [_vp, _i32, _dp, _dp, _vp]),
    "rb_set_solver_cluster": (C.c_int, [_vp, _i32]),
    "rb_get_solver_cluster": (C.c_int, [_vp]),
    "rb_get_solver_max_clusters": (C.c_int, [_vp]),
    "rb_set_tile_skip": (C.c_int, [_vp, _i32]),
    "rb_edge_counts": (C.c_int, [_vp, _ip, _vp]),
    "rb_stored_tiles": (C.c_int, [_vp, _ip, _vp]),
    "rb_live_counts": (C.c_int, [_vp, _ip, _vp]),
    "rb_host_copy_bytes": (_i64, [_vp, _i32]),
    "rb_solver_resident": (C.c_int, [_vp, _ip, _vp]),
    "rb_solver_streamed_passes": (C.c_int, [_vp, _ip, _vp]),
    "rb_set_solver_promote": (C.c_int, [_vp, _i32]),
    "rb_set_profiling": (C.c_int, [_vp, _i32]),
    "rb_get_stage_ms": (C.c_int, [_vp, _fp]),
    "rb_align_batch_host": (C.c_int, [_vp, _dp, _dp, _ip, _dp, _ip, _ip, _dp, _ip, _dp, _ip]),
    "rb_debug_solver_cycles": (C.c_int, [_vp, _vp, C.c_int]),
    "rb_version": (C.c_char_p, []),
    "rb_device_sm_count": (C.c_int, [_vp]),
}


def load() -> C.CDLL:
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"roman_b200: CUDA library not built ({LIB_PATH} missing). Run `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C roman_b200/csrc`. There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    return lib


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        _lib = load()
    return _lib
