#!/usr/bin/env python
"""
Two synthetic robots drive through the same street of objects; each one's object map is cut into submaps, every
submap pair is registered on the GPU in one launch sequence, accepted loop closures are written as align.g2o / align.json
and fused with (made-up) odometry into one pose-graph file — the chain of the reference's demo/loop_closures.py +
offline_rpgo/g2o_file_fusion.py, on stand-in data (the Kimera-Multi demo data is not available offline).

    python examples/all_pairs_loop_closures.py --out /tmp/lc            # needs a B200
    python examples/all_pairs_loop_closures.py --dry-run                # CPU only: stops before the registration
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from roman_b200.align import (MapData, SimpleObject, SubmapAlignParams, SubmapMeta, align_submap_pairs, create_config,  # noqa: E402
                              cut_submaps, g2o_file_fusion, loop_closures, write_g2o, write_lc_json)
from roman_b200.align import prefilter  # noqa: E402

SEM = 64


def make_world(rng, n_obj=140):
    """Objects along a 120 m street: centre, shape attributes, unit descriptor."""
    objs = []
    for k in range(n_obj):
        d = rng.normal(size=SEM)
        e = np.sort(rng.uniform(0.02, 1.0, 3))[::-1]
        objs.append(dict(center=np.array([rng.uniform(0, 120), rng.uniform(-8, 8), rng.uniform(-0.5, 2.5)]),
                         volume=float(np.exp(rng.uniform(np.log(0.05), np.log(5)))), linearity=(e[0] - e[1]) / e[0],
                         planarity=(e[1] - e[2]) / e[0], scattering=e[2] / e[0], descriptor=d / np.linalg.norm(d)))
    return objs


def robot_map(world, rng, T_odom_world, t0, noise=0.04):
    """What one robot mapped: the world's objects in ITS odometry frame (noisy), a trajectory down the street."""
    n = 60
    traj = np.tile(np.eye(4), (n, 1, 1))
    traj[:, 0, 3] = np.linspace(0, 120, n)
    traj = T_odom_world @ traj
    times = t0 + 2.0 * np.arange(n)
    segs = []
    for k, w in enumerate(world):
        if rng.uniform() < 0.15:       # not every object is seen by both robots
            continue
        c = T_odom_world[:3, :3] @ w["center"] + T_odom_world[:3, 3] + rng.normal(0, noise, 3)
        seen = t0 + 2.0 * (w["center"][0] / 120 * (n - 1))
        d = w["descriptor"] + 0.2 * rng.normal(size=SEM) / np.sqrt(SEM)
        o = SimpleObject(center=c, volume=w["volume"] * (1 + rng.normal(0, 0.05)), linearity=w["linearity"],
                         planarity=w["planarity"], scattering=w["scattering"], semantic_descriptor=d / np.linalg.norm(d), id=k)
        o.first_seen, o.last_seen = seen - 3.0, seen + 3.0
        segs.append(o)
    return MapData(segments=segs, trajectory=traj, times=times)


def yaw_pose(yaw, t):
    T = np.eye(4)
    T[:2, :2] = [[np.cos(yaw), -np.sin(yaw)], [np.sin(yaw), np.cos(yaw)]]
    T[:3, 3] = t
    return T


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="/tmp/roman_b200_lc")
    ap.add_argument("--dry-run", action="store_true", help="stop before the GPU registration (no CUDA device needed)")
    args = ap.parse_args()
    rng = np.random.default_rng(0)
    world = make_world(rng)
    frames = [yaw_pose(0.3, [5.0, -2.0, 0.0]), yaw_pose(-1.1, [-30.0, 12.0, 0.5])]   # each robot's odometry frame
    maps = [robot_map(world, rng, np.linalg.inv(F), 100.0 + 1000.0 * r) for r, F in enumerate(frames)]
    sp = SubmapAlignParams(method="roman", semantics_dim=SEM, submap_radius=15.0, submap_center_dist=10.0,
                           submap_center_time=1e9, submap_max_size=40)
    subs = [cut_submaps(m, sp) for m in maps]
    print("submaps per robot:", [len(s) for s in subs], "objects per submap:", [[len(x) for x in s] for s in subs])
    # candidate pairs: the robots' odometry frames differ, so a position prefilter is meaningless here — all pairs
    S0, S1 = len(subs[0]), len(subs[1])
    print(f"{S0} x {S1} = {S0 * S1} submap pairs, largest problem {max(len(a) for a in subs[0]) * max(len(b) for b in subs[1])} associations")
    if args.dry_run:
        d = prefilter.submap_distances([s.position for s in subs[0]], [s.position for s in subs[1]])
        print("dry run: centre distances between odometry frames range", float(d.min()), "to", float(d.max()))
        return
    res = align_submap_pairs(sp, [s.segments for s in subs[0]], [s.segments for s in subs[1]])
    n = res.clipper_num_associations
    print("associations per pair:\n", n.astype(int))
    meta = [[SubmapMeta(s.time, s.pose_gravity_aligned, s.pose_flu, int(np.argmin(np.abs(m.times - s.time)))) for s in run]
            for run, m in zip(subs, maps)]
    lcs = loop_closures(res, meta[0], meta[1], lc_association_thresh=6)
    os.makedirs(os.path.join(args.out, "r0_r1"), exist_ok=True)
    os.makedirs(os.path.join(args.out, "odom"), exist_ok=True)
    write_g2o(os.path.join(args.out, "r0_r1", "align.g2o"), lcs)
    write_lc_json(os.path.join(args.out, "r0_r1", "align.json"), lcs, ["r0", "r1"])
    # the true transform between the two odometry frames is inv(F0) F1; every loop closure implies an estimate of it
    T_true = np.linalg.inv(frames[0]) @ frames[1]
    errs = []
    for c in lcs:
        Ti, Tj = meta[0][c["i"]].pose_flu, meta[1][c["j"]].pose_flu
        T_est = Ti @ c["T_pi_pj"] @ np.linalg.inv(Tj)
        errs.append(np.linalg.norm(T_est[:3, 3] - T_true[:3, 3]))
    print(f"{len(lcs)} loop closures; translation error of the implied frame alignment: median {np.median(errs):.3f} m")
    from scipy.spatial.transform import Rotation as Rot
    for r, m in enumerate(maps):       # odometry files for the fusion step
        with open(os.path.join(args.out, "odom", f"r{r}.g2o"), "w") as f:
            for k, T in enumerate(m.trajectory):
                q = Rot.from_matrix(T[:3, :3]).as_quat()
                f.write("VERTEX_SE3:QUAT " + " ".join(str(v) for v in [k, *T[:3, 3], *q]) + "\n")
    cfg = create_config(["r0", "r1"], os.path.join(args.out, "odom"), args.out, "align")
    cfg["single_lc"] = []              # this example has no single-robot loop closures
    g2o_file_fusion(cfg, os.path.join(args.out, "fused.g2o"), thresh=6)
    print("wrote", os.path.join(args.out, "fused.g2o"))


if __name__ == "__main__":
    main()
