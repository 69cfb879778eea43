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
import pickle
from dataclasses import dataclass
from typing import Any, List, Optional, Sequence

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


# ---------------------------------------------------------------------------------------------------------------------
# The result object of `submap_align` and its writer (reference roman/align/results.py:18-44, 122-246)
@dataclass
class SubmapAlignResults:
    """Field names, order and meaning of the reference dataclass (roman/align/results.py:18-44)."""
    robots_nearby_mat: np.ndarray
    clipper_angle_mat: np.ndarray
    clipper_dist_mat: np.ndarray
    clipper_num_associations: np.ndarray
    similarity_mat: Optional[np.ndarray]
    submap_yaw_diff_mat: np.ndarray
    associated_objs_mat: Any
    T_ij_mat: np.ndarray
    T_ij_hat_mat: np.ndarray
    timing_list: List[float]
    submap_align_params: Any
    submap_io: Any
    total_time: float = -np.inf

    def save(self):
        with open(self.submap_io.output_pkl, "wb") as f:
            pickle.dump(self, f)

    @classmethod
    def load(cls, file_path) -> "SubmapAlignResults":
        with open(file_path, "rb") as f:
            return pickle.load(f)


def plot_align_results(results: SubmapAlignResults, dpi=500) -> bool:
    """The heat-map figure of reference results.py:55-120 — drawn only where matplotlib exists (it is not part of
    this image); returns whether a figure was made."""
    try:
        import matplotlib.pyplot as plt
    except ImportError:
        return False
    show_sim = results.similarity_mat is not None
    names = results.submap_io.robot_names
    if None in results.submap_io.input_gt_pose_yaml:   # no ground truth: only the association counts (and similarity)
        fig, ax = plt.subplots(2 if show_sim else 1, 1, figsize=(8 if show_sim else 4, 4), dpi=dpi)
        ax = np.array(ax).reshape(-1, 1)
        mp = ax[0, 0].imshow(results.clipper_num_associations, cmap="viridis", vmin=0)
        fig.colorbar(mp, fraction=0.04, pad=0.04)
        ax[0, 0].set_title("Number of Associations")
        if show_sim:
            mp = ax[1, 0].imshow(results.similarity_mat, cmap="viridis", vmin=0.0, vmax=1.0)
            fig.colorbar(mp, fraction=0.04, pad=0.04)
            ax[1, 0].set_title("Similarity Score")
        fig.suptitle(f"{results.submap_io.run_name}: {names[0]}, {names[1]}")
        return True
    fig, ax = plt.subplots(3, 2, figsize=(8, 12), dpi=dpi)
    fig.subplots_adjust(wspace=.3)
    fig.suptitle(f"{results.submap_io.run_name}: {names[0]}, {names[1]}")
    angle_thresh, dist_thresh = 10.0, 5.0
    ang, dst = results.clipper_angle_mat.copy(), results.clipper_dist_mat.copy()
    ang[(dst > dist_thresh) & ~np.isnan(ang)] = angle_thresh
    dst[(ang > angle_thresh) & ~np.isnan(dst)] = dist_thresh
    panels = [(0, 0, results.robots_nearby_mat, "magma", dict(vmin=0), "Submaps Center Distance (m)"),
              (0, 1, results.submap_yaw_diff_mat, "magma", dict(vmin=0), "Submap Center Yaw Diff. (deg)"),
              (1, 0, dst, "viridis_r", dict(vmax=dist_thresh, vmin=0.0), "Registration Translation Error (m)"),
              (1, 1, ang, "viridis_r", dict(vmax=angle_thresh, vmin=0.0), "Registration Angle Error (deg)"),
              (2, 0, results.clipper_num_associations, "viridis", dict(vmin=0), "Number of Associations")]
    if show_sim:
        panels.append((2, 1, results.similarity_mat, "viridis", dict(vmin=0.0, vmax=1.0), "Similarity Score"))
    for i, j, mat, cmap, kw, title in panels:
        mp = ax[i, j].imshow(mat, cmap=cmap, **kw)
        fig.colorbar(mp, fraction=0.04, pad=0.04)
        ax[i, j].set_title(title)
    for a in ax.reshape(-1):
        a.set_xlabel("submap index (robot 2)")
        a.set_ylabel("submap index (robot 1)")
        a.grid(False)
    if not show_sim:
        fig.delaxes(ax[2, 1])
    return True


def save_submap_align_results(results: SubmapAlignResults, submaps, roman_maps):
    """
    Everything the reference's writer leaves in `submap_io.output_dir` (roman/align/results.py:122-246): the figure
    (where matplotlib exists), `<run>.matrix.pkl`, `<run>.pkl`, `<run>.timing.txt`, `<run>.params.txt`, `<run>.g2o`,
    `<run>.json` and one `<robot>.sm.json` per robot.  `submaps[r]`: the robot's submaps (time, pose_gravity_aligned,
    pose_flu, segments with .id); `roman_maps[r]`: its map (times, trajectory, segments).
    """
    io, sp = results.submap_io, results.submap_align_params
    if plot_align_results(results):
        import matplotlib.pyplot as plt
        plt.savefig(io.output_img)
    with open(io.output_matrix, "wb") as f:
        pickle.dump([results.robots_nearby_mat, results.clipper_angle_mat, results.clipper_dist_mat,
                     results.clipper_num_associations, results.submap_yaw_diff_mat], f)
    with open(io.output_pkl, "wb") as f:
        pickle.dump(results, f)
    sizes = [len(sm) for sm in list(submaps[0]) + list(submaps[1])]
    with open(io.output_timing, "w") as f:
        f.write(f"Total number of submaps: {len(submaps[0])} x {len(submaps[1])} = {len(submaps[0])*len(submaps[1])}\n")
        f.write(f"Average time per registration: {np.mean(results.timing_list):.4f} seconds\n")
        f.write(f"Total time: {np.sum(results.timing_list):.4f} seconds\n")
        f.write(f"Total number of objects: {np.sum(sizes)}\n")
        f.write(f"Average number of obects per map: {np.mean(sizes):.2f}\n")
    with open(io.output_params, "w") as f:
        f.write(f"{sp}")

    def pose_index(rm, t):   # robotdatapy PoseData.idx(t, force_single=True): the trajectory pose closest in time
        return int(np.argmin(np.abs(np.asarray(rm.times, dtype=np.float64) - t)))
    meta = [[SubmapMeta(sm.time, sm.pose_gravity_aligned, sm.pose_flu, pose_index(rm, sm.time)) for sm in run]
            for run, rm in zip(submaps, roman_maps)]
    pair_res = SubmapPairResults(T_ij_hat_mat=results.T_ij_hat_mat, clipper_num_associations=results.clipper_num_associations,
                                 clipper_percent_associations=None, associated_objs_mat=None, status_mat=None)
    lcs = loop_closures(pair_res, meta[0], meta[1], lc_association_thresh=io.lc_association_thresh,
                        single_robot_lc=sp.single_robot_lc, single_robot_lc_time_thresh=sp.single_robot_lc_time_thresh)
    write_g2o(io.output_g2o, lcs, g2o_t_std=io.g2o_t_std, g2o_r_std=io.g2o_r_std)
    write_lc_json(io.output_lc_json, lcs, io.robot_names)
    for r, path in enumerate(io.output_submaps):
        if path is None:
            continue
        segs = []
        for sg in roman_maps[r].segments:   # the reference skips segments without a point cloud (minimal data), results.py:211-226
            pts = getattr(sg, "points", None)
            if pts is None:
                continue
            segs.append(dict(id=sg.id, centroid_odom=np.mean(np.asarray(pts), axis=0), volume=sg.volume, linearity=sg.linearity,
                             planarity=sg.planarity, scattering=sg.scattering, first_seen=sg.first_seen, last_seen=sg.last_seen))
        write_submaps_json(path, io.robot_names[r], segs,
                           [dict(time=sm.time, pose_gravity_aligned=sm.pose_gravity_aligned,
                                 segment_ids=[s.id for s in sm.segments]) for sm in submaps[r]])
