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


ALL_PAIRS_SOLVER_CLUSTER = 1     # CTAs per problem used by the all-pairs search (see align_submap_pairs)


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
    sel = nearby & attempted
    has_pose = ~np.isnan(T_ij_hat_mat).any(axis=(2, 3))
    ok = sel & has_pose
    angle[sel & ~has_pose] = np.abs(np.rad2deg(180.0))     # the reference's sentinel goes through rad2deg too (:187)
    dist[sel & ~has_pose] = 1e6
    if ok.any():
        T_err = np.linalg.inv(T_ij_hat_mat[ok]) @ T_ij[ok]                     # batched: no Python loop over pairs
        angle[ok] = np.abs(np.rad2deg(Rot.from_matrix(T_err[:, :3, :3]).magnitude()))
        dist[ok] = np.linalg.norm(T_err[:, :3, 3], axis=1)
    return T_ij, angle, dist, yaw


def align_submap_pairs(sm_params: SubmapAlignParams, submaps0: Sequence[List], submaps1: Sequence[List],
                       pair_mask: Optional[np.ndarray] = None, u0: Optional[dict] = None,
                       exclude_common_ids: Optional[bool] = None, rank: int = 0, world: int = 1,
                       positions0=None, positions1=None, skip_distance: float = np.inf,
                       descriptors0=None, descriptors1=None, poses0=None, poses1=None,
                       points0=None, points1=None, timing: Optional[dict] = None) -> SubmapPairResults:
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
    `points0/1`: per submap the object centres in the odometry frame ((n, 3) each) — needed for the "nearby" test of
    force-filled / radius-less submaps, which is a bounding-box intersection (submap_align.py:102; roman/utils.py:160-169).
    `timing`: a dict that receives `registration_seconds` (wall time of the batched registration: upload, kernels,
    read-back) and `n_registered`.
    """
    import time as _time
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
    by_radius = not sm_params.force_fill_submaps and sm_params.submap_radius is not None
    if dist is not None and (by_radius or (points0 is not None and points1 is not None)):
        res.robots_nearby_mat = np.where(prefilter.nearby_mask(sm_params, dist, points0, points1), dist, np.nan)
    for code, gate in ((3, mask & too_far), (4, mask & ~too_far & dissimilar)):   # attempted by the reference, not registered
        res.status_mat[gate] = code
        res.clipper_num_associations[gate] = 0
        res.clipper_percent_associations[gate] = 0.0
    mask = mask & ~too_far & ~dissimilar

    ci, cj = np.nonzero(mask)                      # row-major = the reference's loop order (i outer, j inner)
    ci, cj = ci[rank::world], cj[rank::world]
    pairs, A_list, kept = [], [], {}
    if exclude_common_ids:
        ids0 = [np.array([getattr(o, "id", None) for o in sm], dtype=object) for sm in submaps0]
        ids1 = [np.array([getattr(o, "id", None) for o in sm], dtype=object) for sm in submaps1]
        for i, j in zip(ci.tolist(), cj.tolist()):
            common = set(ids0[i].tolist()) & set(ids1[j].tolist())
            k1 = np.array([k for k, x in enumerate(ids0[i].tolist()) if x not in common], dtype=np.int64)
            k2 = np.array([k for k, x in enumerate(ids1[j].tolist()) if x not in common], dtype=np.int64)
            if len(k1) == 0 or len(k2) == 0:   # register() returns no associations for an empty map
                res.clipper_num_associations[i, j] = 0
                res.clipper_percent_associations[i, j] = 0.0
                res.status_mat[i, j] = 1
                continue
            pairs.append((i, j))
            kept[(i, j)] = (k1, k2)
            # all-to-all over the kept objects, as an explicit association list
            A_list.append(np.stack([np.repeat(k1, len(k2)), np.tile(k2, len(k1))], axis=1).astype(np.int32))
    else:
        empty = (n1_all[ci] == 0) | (n2_all[cj] == 0)
        res.clipper_num_associations[ci[empty], cj[empty]] = 0
        res.clipper_percent_associations[ci[empty], cj[empty]] = 0.0
        res.status_mat[ci[empty], cj[empty]] = 1
        pairs = list(zip(ci[~empty].tolist(), cj[~empty].tolist()))
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
    # all-pairs preset: most submap pairs of a search do not overlap, and a problem without a consistent clique
    # iterates for hundreds of passes — out of the shared-memory CSR after the first few.  One CTA per problem then
    # beats the library's default of two (which exists to shorten the tail of a batch whose problems STREAM their
    # matrix): measured on the 20 x 20 search of bench.py --config cfg5, 7.8 ms against 11.2 ms of solver time per search.
    eng.set_solver_cluster(ALL_PAIRS_SOLVER_CLUSTER)
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
    t0 = _time.perf_counter()
    out = eng.align_batch_host(np.ascontiguousarray(D1), np.ascontiguousarray(D2), u0cat, A=A_cat)
    if timing is not None:
        timing["registration_seconds"] = _time.perf_counter() - t0
        timing["n_registered"] = len(pairs)
        timing["n_empty"] = int((res.status_mat == 1).sum())   # pairs with an empty map: never handed to register()'s solver

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
    # result matrices: one vectorised scatter; only the ragged association lists need a loop
    n_ok = np.where(ok, n_sel, 0).astype(np.float64)
    res.T_ij_hat_mat[pi, pj] = T
    res.clipper_num_associations[pi, pj] = n_ok
    res.status_mat[pi, pj] = status
    if exclude_common_ids:
        len1 = np.array([len(kept[p][0]) for p in pairs], dtype=np.float64)
        len2 = np.array([len(kept[p][1]) for p in pairs], dtype=np.float64)
    else:
        len1, len2 = n1_all[pi].astype(np.float64), n2_all[pj].astype(np.float64)
    mean_len = 0.5 * (len1 + len2)           # lengths after the common-id removal
    res.clipper_percent_associations[pi, pj] = np.where(mean_len > 0, n_ok / np.where(mean_len > 0, mean_len, 1.0), 0.0)
    offs, sel = eng.offsets, out["sel"]
    for b in np.nonzero(ok & (n_sel > 0))[0].tolist():
        i, j = pairs[b]
        o = int(offs[b])
        a = sel[o:o + int(n_sel[b])].copy()
        if exclude_common_ids:   # the reference indexes the submaps AFTER the common-id removal
            k1, k2 = kept[(i, j)]
            a[:, 0] = np.searchsorted(k1, a[:, 0])
            a[:, 1] = np.searchsorted(k2, a[:, 1])
        res.associated_objs_mat[i][j] = a
    return finish(res)


# ---------------------------------------------------------------------------------------------------------------------
# The entry point `demo/loop_closures.py` calls: submap_align(sm_params, sm_io)      (reference submap_align.py:28-220)
class _PickledStub:
    """Stand-in for a class of the reference (roman.*) when a pickle is read without the reference installed: takes
    whatever state was pickled and answers the attribute names `roman.align` reads from map segments."""
    _ALIASES = {"center": "centroid", "volume": "_volume", "linearity": "_linearity", "planarity": "_planarity",
                "scattering": "_scattering", "descriptor": "_descriptor"}

    def __setstate__(self, state):
        if isinstance(state, tuple) and len(state) == 2:   # (dict, slots)
            for part in state:
                if part:
                    self.__dict__.update(part)
        else:
            self.__dict__.update(state)

    def __getattr__(self, name):
        alias = _PickledStub._ALIASES.get(name)
        if alias is not None and alias in self.__dict__:
            return self.__dict__[alias]
        raise AttributeError(name)


def _unpickle_tolerant(f):
    import pickle

    class _U(pickle.Unpickler):
        def find_class(self, module, name):
            try:
                return super().find_class(module, name)
            except (ImportError, AttributeError):
                if module.split(".")[0] in ("roman", "robotdatapy", "open3d"):
                    return type(name, (_PickledStub,), {"__module__": module})
                raise
    return _U(f).load()


@dataclass
class MapSegment:
    """One object of a robot's map, in the odometry frame — the attributes of the reference's Segment /
    SegmentMinimalData that the alignment path reads."""
    id: int
    center: np.ndarray
    volume: float
    linearity: float
    planarity: float
    scattering: float
    extent: np.ndarray
    semantic_descriptor: Optional[np.ndarray]
    first_seen: float
    last_seen: float
    points: Optional[np.ndarray] = None


def load_roman_map(source):
    """
    A robot's map as `MapData` from: a path to a pickled `ROMANMap` (reference roman/map/map.py:18-58, 194-208), an
    already loaded ROMANMap-like object (.segments, .trajectory, .times, .descriptors) or a MapData.  Pickles are read
    with the reference's own classes when `roman` is importable; otherwise with stand-in classes, which works for maps
    saved as minimal data (`ROMANMap.minimal_data()`: every segment carries its volume / shape attributes as numbers).
    A pickle of full Segments needs the reference (and open3d) to turn point clouds into volumes: a clear error says so.
    """
    import os
    from .submaps import MapData
    if isinstance(source, MapData):
        return source
    obj = source
    if isinstance(source, (str, bytes, os.PathLike)):
        with open(os.path.expanduser(source), "rb") as f:
            obj = _unpickle_tolerant(f)
    segs = []
    for sg in obj.segments:
        try:
            c = np.asarray(sg.center, dtype=np.float64).reshape(-1)
            segs.append(MapSegment(id=int(sg.id), center=c, volume=float(sg.volume), linearity=float(sg.linearity),
                                   planarity=float(sg.planarity), scattering=float(sg.scattering),
                                   extent=np.asarray(sg.extent, dtype=np.float64).reshape(-1),
                                   semantic_descriptor=None if getattr(sg, "semantic_descriptor", None) is None
                                   else np.asarray(sg.semantic_descriptor, dtype=np.float64).reshape(-1),
                                   first_seen=float(sg.first_seen), last_seen=float(sg.last_seen),
                                   points=getattr(sg, "points", None)))
        except AttributeError as ex:
            raise ValueError(
                "this ROMANMap holds full Segment objects whose volume / extent / shape attributes are computed from "
                "point clouds by the reference (open3d); save it as roman_map.minimal_data(), or load it with the "
                f"reference installed and pass the object instead of the path ({ex})") from None
    desc = getattr(obj, "descriptors", None)
    return MapData(segments=segs, trajectory=np.asarray(obj.trajectory, dtype=np.float64).reshape(-1, 4, 4),
                   times=np.asarray(obj.times, dtype=np.float64), descriptors=None if desc is None else np.vstack(desc))


def submap_align(sm_params: SubmapAlignParams, sm_io, u0=None):
    """
    Breaks two maps into submaps and registers every submap of the first against every submap of the second —
    the reference's `submap_align(sm_params, sm_io)` (roman/align/submap_align.py:28-220) with the same inputs
    (`SubmapAlignInputOutput`: two ROMANMap pickles) and the same outputs in `sm_io.output_dir` (result pickles, timing
    and parameter files, `<run>.g2o`, `<run>.json`, `<robot>.sm.json`), computed as ONE batched launch sequence on the
    GPU instead of a Python loop of per-pair solves.  Returns the `SubmapAlignResults` it saved.

    `timing_list` keeps the reference's meaning for the timing file (sum = time spent registering, mean = per pair):
    every registered pair is booked with an equal share of the batch's wall time.
    `u0` (not in the reference signature): start vectors as in `align_submap_pairs` — the reference draws them at random.
    Ground-truth pose files (`sm_io.input_gt_pose_yaml`) are read by robotdatapy in the reference, which this image
    lacks: asking for them raises.
    """
    import os
    import time as _time
    from .results import SubmapAlignResults, save_submap_align_results
    from .submaps import cut_submaps
    assert sm_io.input_type_json or sm_io.input_type_pkl, "Invalid input type"
    assert sm_io.input_type_json != sm_io.input_type_pkl, "Only one input type allowed"
    assert not sm_io.input_type_json, "Not currently supported"      # as in the reference (:67-68)
    if any(y is not None for y in sm_io.input_gt_pose_yaml):
        raise NotImplementedError("ground-truth pose files are loaded through robotdatapy.PoseData in the reference; "
                                  "robotdatapy is not available here")
    roman_maps = [load_roman_map(sm_io.inputs[r]) for r in range(2)]
    submaps = [cut_submaps(roman_maps[r], sm_params) for r in range(2)]
    # Kept for identical output files: in the reference `Submap.pose_gravity_aligned` strips roll and pitch from
    # `pose_flu` IN PLACE (roman/utils.py:128-130 writes into its argument; map.py:105-107), so by the time the loop
    # closures are written `pose_flu` IS the gravity-aligned pose and T_pi_pj (results.py:171) relates those frames.
    for run in submaps:
        for sm in run:
            sm.pose_flu = sm.pose_gravity_aligned
    S0, S1 = len(submaps[0]), len(submaps[1])
    total_t0 = _time.time()

    def global_points(run):   # Submap.segments_as_global_points (map.py:131-137): centres back in the odometry frame
        return [np.array([np.asarray(s.center, dtype=np.float64).reshape(-1)[:3] for s in sm.segments]) @ sm.pose_gravity_aligned[:3, :3].T
                + sm.pose_gravity_aligned[:3, 3] for sm in run]
    need_boxes = sm_params.force_fill_submaps or sm_params.submap_radius is None
    descr = [[sm.descriptor for sm in run] for run in submaps] if sm_params.submap_descriptor is not None else [None, None]
    timing = {}
    if S0 and S1:
        pr = align_submap_pairs(
            sm_params, [sm.segments for sm in submaps[0]], [sm.segments for sm in submaps[1]], u0=u0,
            positions0=np.array([sm.position for sm in submaps[0]]), positions1=np.array([sm.position for sm in submaps[1]]),
            skip_distance=sm_io.skip_distance, descriptors0=descr[0], descriptors1=descr[1],
            poses0=np.array([sm.pose_gravity_aligned for sm in submaps[0]]),
            poses1=np.array([sm.pose_gravity_aligned for sm in submaps[1]]),
            points0=global_points(submaps[0]) if need_boxes else None,
            points1=global_points(submaps[1]) if need_boxes else None, timing=timing)
        nan = np.full((S0, S1), np.nan)
        skipped = pr.status_mat == 3               # beyond skip_distance: the reference `continue`s before filling these (:136-142)
        angle = np.where(skipped, np.nan, pr.clipper_angle_mat if pr.clipper_angle_mat is not None else nan)
        dist = np.where(skipped, np.nan, pr.clipper_dist_mat if pr.clipper_dist_mat is not None else nan)
        assoc = [[pr.associated_objs_mat[i][j] if len(pr.associated_objs_mat[i][j]) else [] for j in range(S1)] for i in range(S0)]
        mats = dict(robots_nearby_mat=pr.robots_nearby_mat if pr.robots_nearby_mat is not None else nan, clipper_angle_mat=angle,
                    clipper_dist_mat=dist, clipper_num_associations=pr.clipper_num_associations,
                    similarity_mat=pr.similarity_mat if sm_params.submap_descriptor is not None else None,
                    submap_yaw_diff_mat=pr.submap_yaw_diff_mat if pr.submap_yaw_diff_mat is not None else nan,
                    associated_objs_mat=assoc, T_ij_mat=pr.T_ij_mat, T_ij_hat_mat=pr.T_ij_hat_mat)
    else:
        z = np.zeros((S0, S1)) * np.nan
        mats = dict(robots_nearby_mat=z, clipper_angle_mat=z.copy(), clipper_dist_mat=z.copy(), clipper_num_associations=z.copy(),
                    similarity_mat=z.copy() if sm_params.submap_descriptor is not None else None, submap_yaw_diff_mat=z.copy(),
                    associated_objs_mat=[[[] for _ in range(S1)] for _ in range(S0)],
                    T_ij_mat=np.zeros((S0, S1, 4, 4)) * np.nan, T_ij_hat_mat=np.zeros((S0, S1, 4, 4)) * np.nan)
    # one entry per register() call — and, as in the reference, a second one for a pair whose T_align / gravity check
    # raised afterwards (submap_align.py:157 and :180 both append)
    n_reg = int(timing.get("n_registered", 0))
    n_calls = n_reg + int(timing.get("n_empty", 0))          # an empty map still goes through register() (it returns at once)
    n_failed = int(np.isin(pr.status_mat, (1, 2)).sum()) if (S0 and S1) else 0
    share_t = timing.get("registration_seconds", 0.0) / n_reg if n_reg else 0.0
    timing_list = [share_t] * (n_calls + n_failed)
    results = SubmapAlignResults(timing_list=timing_list, submap_align_params=sm_params, submap_io=sm_io,
                                 total_time=_time.time() - total_t0, **mats)
    os.makedirs(sm_io.output_dir, exist_ok=True)
    save_submap_align_results(results, submaps, roman_maps)
    return results
