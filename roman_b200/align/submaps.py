"""
Cutting a robot's object map into submaps — the producer of the lists `align_submap_pairs` consumes; counterpart of
the reference's `submaps_from_roman_map` / `extract_submap_descriptors` (roman/map/map.py:210-357) with the parameters
`SubmapParams.from_submap_align_params` derives from `SubmapAlignParams` (:178-190).  Host-side numpy; no GPU involved.

Input is plain data (`MapData`): the map's objects in the odometry frame (anything with .center, .first_seen,
.last_seen plus the feature attributes the registration reads), the trajectory (4x4 poses, FLU) with its time stamps
and, optionally, one frame descriptor per trajectory pose.

Behaviour carried over from the reference, because it decides which objects end up in which submap:
  * a new submap starts at the first pose and whenever the position has moved more than `submap_center_dist` from the
    last submap's pose or more than `submap_center_time` seconds have passed (:292-301);
  * an object joins every submap whose pose is closer than `submap_radius` to its centre (no limit if None) unless it
    was last seen more than `submap_center_time` before the previous submap's time or first seen that long after the
    next one's (:307-319);
  * objects are expressed in the submap's gravity-aligned frame (pose with roll and pitch removed, roman/utils.py:128-130)
    (:321-323), then the submap keeps the `submap_max_size` best by the pruning key — time distance to the submap or
    distance to its origin — using a stable sort (:325-333);
  * `force_fill_submaps`: instead, objects sorted by reference time are cut into windows of `submap_max_size` with
    `submap_overlap` shared objects; the submap pose is the trajectory pose closest in time to the window's mean (:264-289);
  * submaps without objects are dropped, ids are NOT renumbered (:335);
  * submap descriptor: mean of the objects' semantic descriptors, the mean of the frame descriptors seen between the
    submap's first and last object sighting, or those frame descriptors stacked (optionally thinned so that consecutive
    kept frames are at least `frame_descriptor_dist` apart) (:337-349, 210-242).
One reference quirk is NOT carried over: there `Submap.pose_gravity_aligned` strips roll and pitch from `pose_flu` IN PLACE
(the same array as the trajectory entry), so after cutting, `pose_flu` equals the gravity-aligned pose.  Here `pose_flu`
stays what the trajectory said and `pose_gravity_aligned` is separate.
"""
from __future__ import annotations

import copy
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

from .params import SubmapAlignParams
from .submap_align import rm_roll_pitch


@dataclass
class MapData:
    segments: Sequence            # objects in the odometry frame
    trajectory: np.ndarray        # (N, 4, 4) FLU poses
    times: np.ndarray             # (N,)
    descriptors: Optional[np.ndarray] = None   # (N, D) one frame descriptor per pose


@dataclass
class SubmapData:
    id: int
    time: float
    pose_flu: np.ndarray
    pose_gravity_aligned: np.ndarray
    segments: List = field(default_factory=list)     # copies of the map's objects, centres in the submap frame
    descriptor: Optional[np.ndarray] = None

    @property
    def position(self) -> np.ndarray:
        return self.pose_flu[:3, 3]

    @property
    def first_seen(self) -> float:
        return min(s.first_seen for s in self.segments)

    @property
    def last_seen(self) -> float:
        return max(s.last_seen for s in self.segments)

    def __len__(self):
        return len(self.segments)


def reference_time(seg) -> float:
    """roman/object/segment.py:57-59 (use_avg_time=True)"""
    return (seg.first_seen + seg.last_seen) / 2.0


def _into_frame(seg, T_center_odom: np.ndarray):
    """Copy of `seg` with its centre moved by T (the reference deep-copies and calls seg.transform)."""
    out = copy.deepcopy(seg)
    c = np.asarray(seg.center, dtype=np.float64).reshape(-1)
    moved = T_center_odom[:3, :3] @ c[:3] + T_center_odom[:3, 3]
    shaped = moved.reshape(np.asarray(seg.center).shape)
    if hasattr(out, "centroid"):
        out.centroid = shaped
    else:
        out.center = shaped
    return out


def _new_submap(k: int, t: float, pose: np.ndarray) -> SubmapData:
    pose = np.array(pose, dtype=np.float64)
    return SubmapData(id=k, time=float(t), pose_flu=pose, pose_gravity_aligned=rm_roll_pitch(pose))


def cut_submaps(map_data: MapData, sm_params: SubmapAlignParams) -> List[SubmapData]:
    traj = np.asarray(map_data.trajectory, dtype=np.float64)
    times = np.asarray(map_data.times, dtype=np.float64)
    max_size = sm_params.submap_max_size
    submaps: List[SubmapData] = []

    if sm_params.force_fill_submaps:
        ordered = sorted(map_data.segments, key=reference_time)
        for start in range(0, len(ordered), max_size - sm_params.submap_overlap):
            window = ordered[start:start + max_size]
            if not window:
                continue
            k = int(np.argmin(np.abs(times - np.average([reference_time(s) for s in window]))))
            sm = _new_submap(len(submaps), times[k], traj[k])
            T = np.linalg.inv(sm.pose_gravity_aligned)
            sm.segments = [_into_frame(s, T) for s in window]
            submaps.append(sm)
    else:
        for k in range(len(traj)):
            if (not submaps or np.linalg.norm(traj[k, :3, 3] - submaps[-1].position) > sm_params.submap_center_dist
                    or times[k] - submaps[-1].time > sm_params.submap_center_time):
                submaps.append(_new_submap(len(submaps), times[k], traj[k]))
        dt = sm_params.submap_center_time
        for k, sm in enumerate(submaps):
            t_prev = submaps[k - 1].time if k > 0 else -np.inf
            t_next = submaps[k + 1].time if k + 1 < len(submaps) else np.inf
            T = np.linalg.inv(sm.pose_gravity_aligned)
            members = []
            for s in map_data.segments:
                near = sm_params.submap_radius is None or \
                    np.linalg.norm(np.asarray(s.center, dtype=np.float64).reshape(-1)[:3] - sm.position) < sm_params.submap_radius
                in_time = not (s.first_seen > t_next + dt or s.last_seen < t_prev - dt)
                if near and in_time:
                    members.append(_into_frame(s, T))
            if max_size is not None:
                if sm_params.submap_pruning_method == "time":
                    key = lambda s, t0=sm.time: abs(reference_time(s) - t0)
                else:
                    key = lambda s: np.linalg.norm(np.asarray(s.center, dtype=np.float64).reshape(-1))
                members = sorted(members, key=key)[:max_size]
            sm.segments = members

    submaps = [sm for sm in submaps if len(sm.segments) > 0]

    kind = sm_params.submap_descriptor
    if kind == "mean_semantic":
        for sm in submaps:
            sm.descriptor = np.mean([s.semantic_descriptor for s in sm.segments], axis=0).flatten()
    elif kind is not None:
        if map_data.descriptors is None:
            raise ValueError("the map carries no frame descriptors to build submap descriptors from")
        frames = np.vstack(map_data.descriptors)
        for sm in submaps:
            seen = (times >= sm.first_seen) & (times <= sm.last_seen)
            if kind == "mean_frame_descriptor":
                sm.descriptor = frames[seen].mean(axis=0)
            elif kind == "stacked_frame_descriptors":
                if sm_params.frame_descriptor_dist is None:
                    sm.descriptor = frames[seen]
                else:
                    keep, last = [], None
                    for f, p in zip(frames[seen], traj[seen][:, :3, 3]):
                        if last is None or np.linalg.norm(p - last) >= sm_params.frame_descriptor_dist:
                            keep.append(f)
                            last = p
                    sm.descriptor = np.vstack(keep)
            else:
                raise ValueError(f"unknown submap_descriptor '{kind}'")
    return submaps
