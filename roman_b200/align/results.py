"""
Loop-closure wire formats produced from the registration results — the part of the reference's
`save_submap_align_results` that feeds the pose-graph stage (reference roman/align/results.py:149-200):
`align.g2o` (EDGE_SE3:QUAT lines) and `align.json`.  SURVEY §8(f) rank 3.

The reference takes times and poses from its `Submap` objects and vertex indices from
robotdatapy's `PoseData.idx`; here they arrive as plain `SubmapMeta` records, so this module has no
dependency outside numpy/scipy.  Quaternions are scipy's `Rotation.from_matrix(R).as_quat()`
([x, y, z, w]), which is what robotdatapy's `transform_to_xyz_quat` (un-vendored) wraps.
"""
from __future__ import annotations

import json
from dataclasses import dataclass
from typing import List, Sequence

import numpy as np
from scipy.spatial.transform import Rotation as Rot

from .submap_align import SubmapPairResults


@dataclass
class SubmapMeta:
    time: float                      # Submap.time [s]
    pose_gravity_aligned: np.ndarray  # T_odom_centre, 4x4 (Submap.pose_gravity_aligned)
    pose_flu: np.ndarray             # T_odom_pose, 4x4 (Submap.pose_flu)
    pose_index: int                  # index of the trajectory pose at `time` (PoseData.idx(time))


def transform_to_xyz_quat(T: np.ndarray):
    t = np.asarray(T, dtype=np.float64)[:3, 3].copy()
    q = Rot.from_matrix(np.asarray(T, dtype=np.float64)[:3, :3]).as_quat()
    return t, q


def loop_closures(results: SubmapPairResults, meta0: Sequence[SubmapMeta], meta1: Sequence[SubmapMeta],
                  lc_association_thresh: int = 4, single_robot_lc: bool = False,
                  single_robot_lc_time_thresh: float = 50.0) -> List[dict]:
    """
    The accepted loop closures in the reference's order (i outer, j inner; results.py:157-165):
    pairs with fewer than `lc_association_thresh` associations are dropped, and for single-robot
    runs pairs closer in time than `single_robot_lc_time_thresh`.  Each record carries the pose of
    frame j expressed in frame i, T_pi_pj = inv(T_odomi_pi) T_odomi_ci T_ci_cj inv(T_odomj_cj) T_odomj_pj.
    """
    out = []
    n = results.clipper_num_associations
    for i in range(len(meta0)):
        for j in range(len(meta1)):
            if not (n[i, j] >= lc_association_thresh):   # NaN (pair not attempted) is dropped like a small count
                continue
            if single_robot_lc and abs(meta0[i].time - meta1[j].time) < single_robot_lc_time_thresh:
                continue
            T_ci_cj = results.T_ij_hat_mat[i, j]
            T_pi_pj = (np.linalg.inv(meta0[i].pose_flu) @ meta0[i].pose_gravity_aligned @ T_ci_cj
                       @ np.linalg.inv(meta1[j].pose_gravity_aligned) @ meta1[j].pose_flu)
            t, q = transform_to_xyz_quat(T_pi_pj)
            out.append(dict(i=i, j=j, n_associations=int(n[i, j]), T_pi_pj=T_pi_pj, t=t, q=q,
                            times=(meta0[i].time, meta1[j].time), idx=(meta0[i].pose_index, meta1[j].pose_index)))
    return out


def write_g2o(path: str, closures: Sequence[dict], g2o_t_std: float = 0.5, g2o_r_std: float = np.deg2rad(0.5)):
    """`align.g2o` exactly as results.py:181-193 lays it out (tabs and trailing spaces included)."""
    I_t = 1 / (g2o_t_std ** 2)
    I_r = 1 / (g2o_r_std ** 2)
    info = np.diag([I_t, I_t, I_t, I_r, I_r, I_r])
    with open(path, "w") as f:
        for c in closures:
            t, q = c["t"], c["q"]
            f.write(f"# LC: {c['n_associations']}\n")
            f.write(f"EDGE_SE3:QUAT a{c['idx'][0]} b{c['idx'][1]} \t")
            f.write(f"{t[0]} {t[1]} {t[2]} \t")
            f.write(f"{q[0]} {q[1]} {q[2]} {q[3]} \t")
            for ii in range(6):
                f.write("".join(f"{info[ii, jj]} " for jj in range(ii, 6)))
                f.write("\t")
            f.write("\n")


def write_lc_json(path: str, closures: Sequence[dict], robot_names: Sequence[str] = ("0", "1")):
    """`align.json` (results.py:167-174, 195-197)."""
    recs = [{
        "seconds": [int(c["times"][0]), int(c["times"][1])],
        "nanoseconds": [int((c["times"][0] % 1) * 1e9), int((c["times"][1] % 1) * 1e9)],
        "names": list(robot_names),
        "translation": np.asarray(c["t"]).tolist(),
        "rotation": np.asarray(c["q"]).tolist(),
        "rotation_convention": "xyzw",
    } for c in closures]
    with open(path, "w") as f:
        json.dump(recs, f, indent=4)


# ---------------------------------------------------------------------------------------------------------------------
# Submap description file (`<robot>.sm.json`, reference roman/align/results.py:199-246)
def _secs_nsecs(t: float) -> dict:
    """reference results.py:47-53 (time_to_secs_nsecs, as_dict=True)"""
    return {"seconds": int(t), "nanoseconds": int((t - int(t)) * 1e9)}


def write_submaps_json(path: str, robot_name: str, segments: Sequence[dict], submaps: Sequence[dict]):
    """
    `segments`: one dict per map segment with keys id, centroid_odom (3,), volume, linearity, planarity, scattering,
    first_seen, last_seen.  `submaps`: one dict per submap with keys time, pose_gravity_aligned (4x4, T_odom_submap),
    segment_ids.  Layout, key order and the time splitting follow the reference writer, so that downstream consumers
    of `*.sm.json` read either file.
    """
    doc = {"segments": [], "submaps": []}
    for sg in segments:
        doc["segments"].append({
            "robot_name": robot_name,
            "segment_index": sg["id"],
            "centroid_odom": np.asarray(sg["centroid_odom"], dtype=np.float64).tolist(),
            "shape_attributes": {"volume": sg["volume"], "linearity": sg["linearity"], "planarity": sg["planarity"],
                                 "scattering": sg["scattering"]},
            "first_seen": _secs_nsecs(sg["first_seen"]),
            "last_seen": _secs_nsecs(sg["last_seen"]),
        })
    for j, sm in enumerate(submaps):
        t, q = transform_to_xyz_quat(sm["pose_gravity_aligned"])
        doc["submaps"].append({
            "submap_index": j,
            "T_odom_submap": {"tx": t[0], "ty": t[1], "tz": t[2], "qx": q[0], "qy": q[1], "qz": q[2], "qw": q[3]},
            "robot_name": robot_name,
            "seconds": int(sm["time"]),
            "nanoseconds": int((sm["time"] % 1) * 1e9),
            "segment_indices": list(sm["segment_ids"]),
        })
    with open(path, "w") as f:
        json.dump(doc, f, indent=4)
