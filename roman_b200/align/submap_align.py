"""
All-pairs submap registration — the batched counterpart of the double loop of the reference's
`submap_align` (roman/align/submap_align.py:93-200), SURVEY §8(f) rank 1 "caller-side batching".

The reference walks every (submap i of run 0, submap j of run 1) pair serially: deep-copies both
submaps, rebuilds their feature rows, constructs a fresh CLIPPER object, registers, fits T, applies
the gravity post-filters.  Here every submap is packed into a feature table ONCE, the table is
uploaded once (rb_plan_indexed: problems reference row ranges of shared tables), all candidate
pairs are registered in one launch sequence, and the post-filters of submap_align.py:166-184 run
vectorised on the returned poses.

The tests the reference makes before registering a pair — centre distance against `skip_distance`
(:136) and the submap-descriptor similarity against `submap_descriptor_thresh` (:131-134, :144) — are
evaluated for all pairs at once (roman_b200.align.prefilter; the similarity on the device) and
produce the work list.

Out of scope (other subsystems of the reference): loading ROMANMap pickles, cutting submaps,
ground-truth pose handling.  Result writers: roman_b200.align.results.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

from . import prefilter
from .dist_reg_with_pruning import zyx_euler
from .params import SubmapAlignParams
from .roman_registration import ROMANRegistration


@dataclass
class SubmapPairResults:
    """The per-pair result matrices of the reference's SubmapAlignResults that come out of the
    registration itself (roman/align/results.py:18-44), indexed [i, j]."""
    T_ij_hat_mat: np.ndarray                 # (S0, S1, 4, 4) float64, NaN where no pose
    clipper_num_associations: np.ndarray     # (S0, S1) float64, NaN where the pair was not attempted
    clipper_percent_associations: np.ndarray # (S0, S1)
    associated_objs_mat: List[List[np.ndarray]]
    status_mat: np.ndarray                   # (S0, S1) int32: -1 skipped, 0 ok, 1 insufficient, 2 gravity-rejected,
                                             #   3 beyond skip_distance, 4 similarity below the threshold
    similarity_mat: Optional[np.ndarray] = None  # (S0, S1) submap-descriptor similarity (None: no descriptor in use;
                                                 #   NaN beyond skip_distance, as in the reference)
    submap_distance_mat: Optional[np.ndarray] = None
    robots_nearby_mat: Optional[np.ndarray] = None  # centre distance where the submaps are "nearby" (:101-103), else NaN
    # filled when the submap poses are given (pair_error_metrics): the reference's evaluation matrices
    T_ij_mat: Optional[np.ndarray] = None           # (S0, S1, 4, 4) relative pose of the two submap frames from their poses
    clipper_angle_mat: Optional[np.ndarray] = None  # registration rotation error, nearby pairs only
    clipper_dist_mat: Optional[np.ndarray] = None   # registration translation error, nearby pairs only
    submap_yaw_diff_mat: Optional[np.ndarray] = None
    pairs: np.ndarray = field(default_factory=lambda: np.zeros((0, 2), dtype=np.int64))


def rm_roll_pitch(T: np.ndarray) -> np.ndarray:
    """Keep only the yaw of the rotation (reference roman/utils.py:128-130: Rz(as_euler('ZYX')[0]))."""
    out = T.copy()
    yaw = np.arctan2(T[..., 1, 0], T[..., 0, 0])
    c, s = np.cos(yaw), np.sin(yaw)
    out[..., :3, :3] = 0.0
    out[..., 0, 0], out[..., 0, 1], out[..., 1, 0], out[..., 1, 1], out[..., 2, 2] = c, -s, s, c, 1.0
    return out


def roll_pitch(T: np.ndarray):
    """roll, pitch of R = Rz(yaw) Ry(pitch) Rx(roll) — the angles the reference reads from
    robotdatapy.transform.transform_to_xyzrpy (submap_align.py:168-169)."""
    pitch = -np.arcsin(np.clip(T[..., 2, 0], -1.0, 1.0))
    roll = np.arctan2(T[..., 2, 1], T[..., 2, 2])
    return roll, pitch


def pair_error_metrics(T_ij_hat_mat: np.ndarray, poses0: np.ndarray, poses1: np.ndarray, nearby: np.ndarray,
                       attempted: np.ndarray):
    """
    The evaluation matrices of the reference loop (roman/align/submap_align.py:117-129, 158-192) from the estimated
    poses: T_ij = inv(T_wi) T_wj of the (gravity-aligned, possibly ground-truth) submap poses; for NEARBY pairs the
    absolute relative yaw in degrees, and — for pairs that went through the loop body (`attempted`: not beyond
    skip_distance) — the rotation / translation error of T_ij_hat against T_ij.  Pairs without a pose get the
    reference's sentinels theta = 180.0 and dist = 1e6; theta then passes through the same rad2deg as a real angle
    (:187), so the matrix holds 10313.24 for them — kept as is.
    Returns (T_ij_mat, clipper_angle_mat, clipper_dist_mat, submap_yaw_diff_mat).
    """
    from scipy.spatial.transform import Rotation as Rot
    poses0, poses1 = np.asarray(poses0, dtype=np.float64), np.asarray(poses1, dtype=np.float64)
    S0, S1 = len(poses0), len(poses1)
    T_ij = np.einsum("iab,jbc->ijac", np.linalg.inv(poses0), poses1)
    angle = np.full((S0, S1), np.nan)
    dist = np.full((S0, S1), np.nan)
    yaw = np.full((S0, S1), np.nan)
    yaw[nearby] = np.abs(np.rad2deg(np.arctan2(T_ij[..., 1, 0], T_ij[..., 0, 0])))[nearby]
    for i in range(S0):
        for j in range(S1):
            if not (nearby[i, j] and attempted[i, j]):
                continue
            That = T_ij_hat_mat[i, j]
            if np.isnan(That).any():
                theta, dd = 180.0, 1e6
            else:
                T_err = np.linalg.inv(That) @ T_ij[i, j]
                theta = Rot.from_matrix(T_err[:3, :3]).magnitude()
                dd = np.linalg.norm(T_err[:3, 3])
            angle[i, j] = np.abs(np.rad2deg(theta))
            dist[i, j] = dd
    return T_ij, angle, dist, yaw


def align_submap_pairs(sm_params: SubmapAlignParams, submaps0: Sequence[List], submaps1: Sequence[List],
                       pair_mask: Optional[np.ndarray] = None, u0: Optional[dict] = None,
                       exclude_common_ids: Optional[bool] = None, rank: int = 0, world: int = 1,
                       positions0=None, positions1=None, skip_distance: float = np.inf,
                       descriptors0=None, descriptors1=None, poses0=None, poses1=None) -> SubmapPairResults:
    """
    Register every masked (i, j) pair of object lists.  `u0` optionally fixes the solver's start vectors (the
    reference draws them at random): a dict {(i, j): vector} or a callable (i, j, m) -> vector, called once per
    registered pair in the reference's loop order (i outer, j inner).  `exclude_common_ids` (default:
    sm_params.single_robot_lc) drops objects whose `id` occurs in both submaps of a pair before
    registering it (submap_align.py:108-115).  With world > 1 only the pairs g with g mod world == rank
    are computed (see roman_b200.dist for the all-gather of the records).

    Prefilters (all optional): `positions0/1` (S, 3) submap centres + `skip_distance` reproduce the skip of
    submap_align.py:136-142 (0 associations, NaN pose); `descriptors0/1` (per submap: (N, D) stacked frame
    descriptors or a (D,) mean descriptor) with `sm_params.submap_descriptor` set reproduce the similarity
    gate of :131-134, :144-149.  Both are evaluated for every pair before anything is registered.
    `poses0/1` (S, 4, 4): gravity-aligned submap poses T_odom_centre (ground truth where available) — with them (and
    positions) the result also carries the reference's evaluation matrices, see pair_error_metrics.
    """
    from ..engine import Engine, default_solver_params
    from .. import clipperpy
    reg = sm_params.get_object_registration()
    if not isinstance(reg, ROMANRegistration):
        raise NotImplementedError("batched all-pairs registration is implemented for the ROMAN-invariant methods")
    if sm_params.dim != 3:
        raise NotImplementedError("the gravity post-filters of submap_align are 3-D")
    if exclude_common_ids is None:
        exclude_common_ids = bool(sm_params.single_robot_lc)
    S0, S1 = len(submaps0), len(submaps1)
    mask = np.ones((S0, S1), dtype=bool) if pair_mask is None else np.asarray(pair_mask, dtype=bool)

    dist = sim = None
    if positions0 is not None and positions1 is not None:
        dist = prefilter.submap_distances(positions0, positions1)
    if sm_params.submap_descriptor is not None:
        if descriptors0 is None or descriptors1 is None:
            raise ValueError("sm_params.submap_descriptor is set: descriptors0 / descriptors1 are required")
        reg._batch_engine = reg._batch_engine or Engine()
        sim = prefilter.submap_similarities(descriptors0, descriptors1, engine=reg._batch_engine)
    too_far = (dist > skip_distance) if dist is not None else np.zeros((S0, S1), dtype=bool)
    dissimilar = (sim < sm_params.submap_descriptor_thresh) if sim is not None else np.zeros((S0, S1), dtype=bool)

    D1, n1_all = reg.pack(submaps0)   # every submap packed once
    D2, n2_all = reg.pack(submaps1)
    off1 = np.concatenate([[0], np.cumsum(n1_all)]).astype(np.int64)
    off2 = np.concatenate([[0], np.cumsum(n2_all)]).astype(np.int64)

    res = SubmapPairResults(
        T_ij_hat_mat=np.full((S0, S1, 4, 4), np.nan), clipper_num_associations=np.full((S0, S1), np.nan),
        clipper_percent_associations=np.full((S0, S1), np.nan),
        associated_objs_mat=[[np.zeros((0, 2), dtype=np.int32) for _ in range(S1)] for _ in range(S0)],
        status_mat=np.full((S0, S1), -1, dtype=np.int32), submap_distance_mat=dist,
        similarity_mat=None if sim is None else np.where(too_far, np.nan, sim))   # :136-142 `continue`s before :196
    if dist is not None and not sm_params.force_fill_submaps and sm_params.submap_radius is not None:
        res.robots_nearby_mat = np.where(prefilter.nearby_mask(sm_params, dist), dist, np.nan)
    for code, gate in ((3, mask & too_far), (4, mask & ~too_far & dissimilar)):   # attempted by the reference, not registered
        res.status_mat[gate] = code
        res.clipper_num_associations[gate] = 0
        res.clipper_percent_associations[gate] = 0.0
    mask = mask & ~too_far & ~dissimilar

    cand = [(i, j) for i in range(S0) for j in range(S1) if mask[i, j]]
    cand = cand[rank::world]
    pairs, A_list, kept = [], [], {}
    for (i, j) in cand:
        k1, k2 = np.arange(n1_all[i]), np.arange(n2_all[j])
        if exclude_common_ids:
            ids_i = {getattr(o, "id", None) for o in submaps0[i]}
            ids_j = {getattr(o, "id", None) for o in submaps1[j]}
            common = ids_i & ids_j
            k1 = np.array([k for k, o in enumerate(submaps0[i]) if getattr(o, "id", None) not in common], dtype=np.int64)
            k2 = np.array([k for k, o in enumerate(submaps1[j]) if getattr(o, "id", None) not in common], dtype=np.int64)
        if len(k1) == 0 or len(k2) == 0:   # register() returns no associations for an empty map
            res.clipper_num_associations[i, j] = 0
            res.clipper_percent_associations[i, j] = 0.0
            res.status_mat[i, j] = 1
            continue
        pairs.append((i, j))
        kept[(i, j)] = (k1, k2)
        if exclude_common_ids:  # all-to-all over the kept objects, as an explicit association list
            A_list.append(np.stack([np.repeat(k1, len(k2)), np.tile(k2, len(k1))], axis=1).astype(np.int32))
    res.pairs = np.array(pairs, dtype=np.int64).reshape(-1, 2)
    def finish(res):
        if poses0 is not None and poses1 is not None and res.robots_nearby_mat is not None:
            (res.T_ij_mat, res.clipper_angle_mat, res.clipper_dist_mat, res.submap_yaw_diff_mat) = pair_error_metrics(
                res.T_ij_hat_mat, poses0, poses1, ~np.isnan(res.robots_nearby_mat), ~too_far)
        return res

    if not pairs:
        return finish(res)

    pi, pj = res.pairs[:, 0], res.pairs[:, 1]
    eng = reg._batch_engine = reg._batch_engine or Engine()
    eng.set_invariant_params(clipperpy.invariants.ROMAN(reg.iparams)._rb_params())
    eng.set_solver_params(default_solver_params())
    m = [len(a) for a in A_list] if exclude_common_ids else None
    eng.plan(D1.shape[1], n1_all[pi], n2_all[pj], m=m, obj1_offsets=off1[pi], obj2_offsets=off2[pj],
             total_obj1=int(off1[-1]), total_obj2=int(off2[-1]))
    msz = np.diff(eng.offsets)
    rng = np.random.default_rng()
    def start_vector(b, i, j):
        if callable(u0):
            v = u0(i, j, int(msz[b]))
        elif u0 is not None and (i, j) in u0:
            v = u0[(i, j)]
        else:
            return rng.uniform(0.0, 1.0, int(msz[b]))
        v = np.asarray(v, dtype=np.float64).reshape(-1)
        assert v.size == msz[b], f"start vector of pair ({i}, {j}) must have {msz[b]} entries"
        return v
    u0cat = np.concatenate([start_vector(b, int(i), int(j)) for b, (i, j) in enumerate(pairs)])
    A_cat = np.ascontiguousarray(np.concatenate(A_list)) if exclude_common_ids else None
    out = eng.align_batch_host(np.ascontiguousarray(D1), np.ascontiguousarray(D2), u0cat, A=A_cat)

    T = out["T"].copy()
    status = out["status"].astype(np.int32).copy()
    n_sel = out["n_sel"]
    ok = status == 0
    # post-filters of submap_align.py:166-172
    if sm_params.force_rm_upside_down and ok.any():
        roll, pitch = roll_pitch(T)
        bad = ok & ((np.abs(roll) > np.deg2rad(90.0)) | (np.abs(pitch) > np.deg2rad(90.0)))
        status[bad] = 2          # GravityConstraintError branch: NaN pose, no associations
        ok &= ~bad
    if sm_params.force_rm_lc_roll_pitch and ok.any():
        T[ok] = rm_roll_pitch(T[ok])
    T[~ok] = np.nan
    for b, (i, j) in enumerate(pairs):
        o = int(eng.offsets[b])
        a = out["sel"][o:o + int(n_sel[b])].copy() if ok[b] else np.zeros((0, 2), dtype=np.int32)
        k1, k2 = kept[(i, j)]
        if exclude_common_ids and len(a):   # the reference indexes the submaps AFTER the common-id removal
            a[:, 0] = np.searchsorted(k1, a[:, 0])
            a[:, 1] = np.searchsorted(k2, a[:, 1])
        res.associated_objs_mat[i][j] = a
        res.T_ij_hat_mat[i, j] = T[b]
        res.clipper_num_associations[i, j] = len(a)
        mean_len = 0.5 * (len(k1) + len(k2))  # lengths after the common-id removal
        res.clipper_percent_associations[i, j] = len(a) / mean_len if mean_len > 0 else 0.0
        res.status_mat[i, j] = status[b]
    return finish(res)
