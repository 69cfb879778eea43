#!/usr/bin/env python
"""
End-to-end golden for `submap_align(sm_params, sm_io)`: the reference's OWN function (roman/align/submap_align.py:28-220,
with roman/map/map.py's map loading + submap cutting and roman/align/results.py's writer) run on two pickled ROMANMaps.

    python tests/golden/make_submap_align_e2e_golden.py
        -> tests/golden/roman_map_a.pkl, roman_map_b.pkl      (pickles of the reference's ROMANMap / SegmentMinimalData)
        -> tests/golden/reference_submap_align_e2e.json       (every file the reference wrote + its result matrices)

Stand-ins, as in make_reference_py_golden.py (whose stubs are reused): matplotlib / open3d / robotdatapy / shapely are
absent -> inert stubs; robotdatapy's transform_to_xyzrpy / transform_to_xyz_quat / PoseData.idx -> scipy one-liners;
the compiled clipperpy -> tests/golden/oracle_clipperpy.py (C oracle), start vectors default_rng(900) in call order.
The reference converts full Segments (point clouds, open3d) to SegmentMinimalData on the fly; the pickles here hold
SegmentMinimalData already, so the three Segment methods the cutting code calls (set_center_ref, minimal_data,
transform) are added to that class for the run.
"""
import copy
import json
import os
import pickle
import tempfile

import numpy as np

import make_reference_py_golden as G      # installs the stubs and the oracle-backed clipperpy, imports the reference
from scipy.spatial.transform import Rotation as Rot

import roman.align.results as ref_results
import roman.align.submap_align as ref_sa
from roman.map.map import ROMANMap
from roman.object.segment import SegmentMinimalData
from roman.params.submap_align_params import SubmapAlignInputOutput, SubmapAlignParams

HERE = os.path.dirname(os.path.abspath(__file__))
SEM = 8


def two_robot_scenario():
    """Robot a: golden_inputs.map_scenario().  Robot b: the same street seen from another odometry frame, driven in the
    same direction a few minutes later: 85 % of the objects, 3 cm noise, its own ids; plus 10 objects only b saw."""
    sc = G.golden_inputs.map_scenario()
    rng = np.random.default_rng(901)
    T_ba = np.eye(4)                      # odometry frame of b <- odometry frame of a
    T_ba[:3, :3] = Rot.from_euler("z", 0.8).as_matrix()
    T_ba[:3, 3] = [12.0, -7.0, 0.4]
    segs_b = []
    for g in sc["segments"]:
        if rng.uniform() < 0.15:
            continue
        c = T_ba[:3, :3] @ g["center"] + T_ba[:3, 3] + rng.normal(0, 0.03, 3)
        d = g["semantic_descriptor"] + 0.05 * rng.normal(size=SEM)
        segs_b.append(dict(id=1000 + g["id"], center=c, volume=g["volume"] * (1 + rng.normal(0, 0.03)), linearity=g["linearity"],
                           planarity=g["planarity"], scattering=g["scattering"], extent=g["extent"], semantic_descriptor=d,
                           first_seen=g["first_seen"] + 400.0, last_seen=g["last_seen"] + 400.0))
    n = len(sc["times"])
    for j in range(10):
        k = int(rng.integers(0, n))
        c = (T_ba @ sc["trajectory"][k])[:3, 3] + rng.normal(0, 5.0, 3) * np.array([1.0, 1.0, 0.2])
        first = sc["times"][k] + 400.0 - rng.uniform(0, 6)
        segs_b.append(dict(id=2000 + j, center=c, volume=float(rng.uniform(0.1, 3)), linearity=float(rng.uniform(0, 1)),
                           planarity=float(rng.uniform(0, 1)), scattering=float(rng.uniform(0, 1)), extent=rng.uniform(0.2, 2, 3),
                           semantic_descriptor=rng.normal(size=SEM), first_seen=float(first), last_seen=float(first + rng.uniform(1, 15))))
    traj_b = np.array([T_ba @ T for T in sc["trajectory"]])
    b = dict(trajectory=traj_b, times=sc["times"] + 400.0 + rng.uniform(0, 0.2, n), segments=segs_b,
             descriptors=sc["descriptors"] + 0.02 * rng.normal(size=sc["descriptors"].shape))
    return sc, b


def to_roman_map(m):
    segs = [SegmentMinimalData(g["id"], np.asarray(g["center"], dtype=np.float64).copy(), float(g["volume"]), float(g["linearity"]),
                               float(g["planarity"]), float(g["scattering"]), np.asarray(g["extent"], dtype=np.float64),
                               np.asarray(g["semantic_descriptor"], dtype=np.float64), float(g["first_seen"]), float(g["last_seen"]))
            for g in m["segments"]]
    return ROMANMap(segments=segs, trajectory=[np.array(T) for T in m["trajectory"]], times=np.array(m["times"]),
                    descriptors=[d for d in m["descriptors"]])


def nan_list(a):
    return [[None if (v is None or np.isnan(v)) else float(v) for v in row] for row in np.asarray(a, dtype=np.float64)]


def main():
    a, b = two_robot_scenario()
    paths = [os.path.join(HERE, "roman_map_a.pkl"), os.path.join(HERE, "roman_map_b.pkl")]
    for m, path in zip((a, b), paths):
        with open(path, "wb") as f:
            pickle.dump(to_roman_map(m), f, protocol=4)

    # the three methods of a full Segment the cutting code calls (robotdatapy.transform.transform is absent)
    SegmentMinimalData.set_center_ref = lambda self, ref: None
    SegmentMinimalData.minimal_data = lambda self: copy.deepcopy(self)

    def _transform(self, T):
        self.centroid = (T[:3, :3] @ self.centroid.reshape(3) + T[:3, 3]).reshape(3, 1)
    SegmentMinimalData.transform = _transform

    class _PoseData:
        def __init__(self, times):
            self.times = np.asarray(times, dtype=np.float64)

        @classmethod
        def from_times_and_poses(cls, times, poses):
            return cls(times)

        def idx(self, t, force_single=True):
            return int(np.argmin(np.abs(self.times - t)))

    def xyz_quat(T, separate=False):
        t, q = np.asarray(T)[:3, 3].copy(), Rot.from_matrix(np.asarray(T)[:3, :3]).as_quat()
        return (t, q) if separate else np.concatenate([t, q]).tolist()

    def xyzrpy(T):
        return np.concatenate([np.asarray(T)[:3, 3], Rot.from_matrix(np.asarray(T)[:3, :3]).as_euler("ZYX")[::-1]])

    ref_results.PoseData = _PoseData
    ref_results.transform_to_xyz_quat = xyz_quat
    ref_results.plot_align_results = lambda *a_, **k_: None
    ref_results.plt.savefig = lambda *a_, **k_: None
    ref_sa.transform_to_xyzrpy = xyzrpy
    ref_sa.tqdm = lambda it: it
    import roman.map.map as ref_map
    ref_map.transform = lambda T, pts, axis=0: (np.asarray(pts) @ T[:3, :3].T + T[:3, 3])   # segments_as_global_points

    cases = []
    for name, kw, iokw in (
            ("default", dict(method="roman", semantics_dim=SEM), dict()),
            ("descriptor_gate_and_skip", dict(method="roman", semantics_dim=SEM, submap_descriptor="mean_frame_descriptor",
                                              submap_descriptor_thresh=0.2, submap_max_size=12), dict(skip_distance=45.0, lc_association_thresh=5)),
            ("force_fill_geometry_only", dict(method="gravity", force_fill_submaps=True, submap_max_size=14, submap_overlap=5), dict())):
        G.shim.seed_u0(900)
        del G.shim.U0_LOG[:]
        with tempfile.TemporaryDirectory() as tmp:
            io = SubmapAlignInputOutput(inputs=paths, output_dir=tmp, run_name="align", robot_names=["alpha", "beta"], **iokw)
            sp = SubmapAlignParams(**kw)
            ref_sa.submap_align(sp, io)
            res = ref_results.SubmapAlignResults.load(io.output_pkl)
            files = {k: open(getattr(io, k)).read() for k in ("output_g2o", "output_lc_json", "output_timing", "output_params")}
            files["output_submaps"] = [open(p).read() for p in io.output_submaps]
            with open(io.output_matrix, "rb") as f:
                mats = pickle.load(f)
        cases.append(dict(
            name=name, sm_kwargs=kw, io_kwargs=iokw, u0_seed=900, u0_sizes=[int(len(u)) for u in G.shim.U0_LOG], files=files,
            n_timing=len(res.timing_list), matrix_pkl=[nan_list(m) for m in mats],
            clipper_num_associations=nan_list(res.clipper_num_associations), robots_nearby_mat=nan_list(res.robots_nearby_mat),
            clipper_angle_mat=nan_list(res.clipper_angle_mat), clipper_dist_mat=nan_list(res.clipper_dist_mat),
            submap_yaw_diff_mat=nan_list(res.submap_yaw_diff_mat),
            similarity_mat=None if res.similarity_mat is None else nan_list(res.similarity_mat),
            T_ij_mat=np.asarray(res.T_ij_mat).tolist(),
            T_ij_hat_mat=[[None if np.isnan(T).any() else np.asarray(T).tolist() for T in row] for row in res.T_ij_hat_mat],
            associated_objs_mat=[[np.asarray(x).reshape(-1, 2).tolist() for x in row] for row in res.associated_objs_mat]))
        print(name, "submaps", np.asarray(res.clipper_num_associations).shape, "registered", len(res.timing_list),
              "closures", files["output_g2o"].count("EDGE_SE3"))
    with open(os.path.join(HERE, "reference_submap_align_e2e.json"), "w") as f:
        json.dump(dict(source="roman/align/submap_align.py:28-220 run end to end by tests/golden/make_submap_align_e2e_golden.py",
                       cases=cases), f)
    print("wrote", len(cases), "cases")


if __name__ == "__main__":
    main()
