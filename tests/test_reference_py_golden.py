"""
CPU tests against tests/golden/reference_py_vectors.json — outputs of the reference's own Python code
(tests/golden/make_reference_py_golden.py imports it from /root/reference and runs it unmodified, with the
C oracle standing in for the compiled clipperpy).  They pin (i) the host layer of this repository
(parameter factory, feature packing, pruning mask, post-filter helpers) and (ii) the oracle's restatement
of the Python-side numerics (Arun/SVD, submap similarity, bounding boxes) to the real thing.
The GPU counterparts are in tests/test_gpu_reference_py.py.
"""
import json
import os
import sys

import numpy as np
import pytest

from oracle import clipper_np as cn
from roman_b200 import synth
from roman_b200.align import DistRegWithPruning, ROMANRegistration, SubmapAlignParams
from roman_b200.align import prefilter
from roman_b200.align.objects import objects_from_rows
from roman_b200.align.submap_align import rm_roll_pitch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import golden_inputs  # noqa: E402

TUNED = dict(sigma=0.3, epsilon=0.5, mindist=0.1, epsilon_shape=0.2, cosine_min=0.6, cosine_max=0.9,
             semantics_dim=64, gravity_unc_ang_rad=0.05)


@pytest.fixture(scope="module")
def gold():
    with open(os.path.join(HERE, "golden", "reference_py_vectors.json")) as f:
        return json.load(f)


def test_factory_equals_reference_factory(gold):
    """SubmapAlignParams.get_object_registration -> the invariant parameters the reference builds
    (roman/params/submap_align_params.py:86-150 + roman/align/roman_registration.py:38-80)."""
    assert len(gold["factory"]) == 22
    for key, want in gold["factory"].items():
        name, tag = key.split("/")
        reg = SubmapAlignParams(method=name, **(TUNED if tag == "tuned" else {})).get_object_registration()
        if want["cls"] == "DistRegWithPruning":
            assert isinstance(reg, DistRegWithPruning)
            for k in ("sigma", "epsilon", "mindist", "shape_epsilon", "cos_min", "use_gravity", "dim"):
                assert getattr(reg, k) == want[k], (key, k)
            assert float(reg.roll_pitch_thresh) == want["roll_pitch_thresh"]
            continue
        assert isinstance(reg, ROMANRegistration)
        for k in ("dim", "volume", "extent", "pca", "semantics"):
            assert getattr(reg, k) == want[k], (key, k)
        for k, v in want["iparams"].items():
            got = getattr(reg.iparams, k)
            if k == "ratio_epsilon":
                assert np.array_equal(np.asarray(got, dtype=np.float64), np.asarray(v, dtype=np.float64)), key
            else:
                assert got == v, (key, k, got, v)


def test_feature_rows_equal_reference_packing(gold):
    """_object_to_clipper_list / rows() / pack() against roman_registration.py:98-108."""
    for case in gold["registration"]:
        sem = case["semantics_dim"]
        p = synth.make_pair(case["seed"], case["n1"], case["n2"], method=case["method"], semantics_dim=sem)
        m1 = objects_from_rows(p.D1, case["method"], sem)
        reg = SubmapAlignParams(method=case["method"], semantics_dim=sem).get_object_registration()
        want = np.array(case["packed_rows"])
        assert np.array_equal(np.array([reg._object_to_clipper_list(o) for o in m1[:2]]), want)
        assert np.array_equal(reg.rows(m1)[:2], want)
        table, counts = reg.pack([m1, m1[:3]])
        assert counts.tolist() == [len(m1), 3] and np.array_equal(table[:2], want)


def test_oracle_end_to_end_equals_reference_python_flow(gold, coracle):
    """The oracle's one-call pair alignment == reference register() + T_align() (numpy LAPACK SVD)."""
    for case in gold["registration"]:
        sem = case["semantics_dim"]
        p = synth.make_pair(case["seed"], case["n1"], case["n2"], method=case["method"], semantics_dim=sem)
        reg = SubmapAlignParams(method=case["method"], semantics_dim=sem).get_object_registration()
        ip = cn.InvariantParams(kind=0, **{k: getattr(reg.iparams, k) for k in (
            "point_dim", "ratio_feature_dim", "cos_feature_dim", "sigma", "epsilon", "mindist", "cosine_min",
            "cosine_max", "gravity_guided", "gravity_unc_ang_rad")})
        ip.ratio_epsilon = np.asarray(reg.iparams.ratio_epsilon, dtype=np.float64)
        u0 = np.random.default_rng(case["u0_seed"]).uniform(0.0, 1.0, case["n1"] * case["n2"])
        for res in (coracle.align_pair(ip, cn.SolverParams(), p.D1, p.D2, u0),
                    cn.align_pair(ip, cn.SolverParams(), p.D1, p.D2, u0)):
            assert np.asarray(res["sel"]).tolist() == case["associations"], case["method"]
            T = np.asarray(case["T"])
            assert np.linalg.norm(np.asarray(res["T"]) - T) / np.linalg.norm(T) < 1e-10


def test_oracle_arun_equals_reference_t_align(gold, coracle):
    """orc_arun (Jacobi) and the numpy restatement against ObjectRegistration.T_align
    (object_registration.py:88-129) incl. the reflection fix and rank-deficient inputs."""
    seen = set()
    for case in gold["t_align"]:
        if case["dim"] != 3 or case["T"] is None:
            continue
        corr = np.asarray(case["correspondences"])
        P1 = np.asarray(case["pts1"])[corr[:, 0]]
        P2 = np.asarray(case["pts2"])[corr[:, 1]]
        T = np.asarray(case["T"])
        rc, Tc = coracle.arun(P1, P2)
        assert rc == 0
        for got in (Tc, cn.T_align(P1, P2)):
            assert np.linalg.norm(got - T) / np.linalg.norm(T) < 1e-9, case["name"]
        assert abs(np.linalg.det(T[:3, :3]) - 1.0) < 1e-9
        seen.add(case["name"])
    assert {"rigid_k3", "rigid_k40", "mirror_z", "coplanar"} <= seen


def test_prune_mask_equals_reference_pruning(gold):
    """DistRegWithPruning.putative_associations against the reference's two np.delete passes
    (dist_reg_with_pruning.py:70-91), executed by the reference itself."""
    for case in gold["clipper_prune"]:
        p = synth.make_pair(case["seed"], case["n1"], case["n2"], method="roman", semantics_dim=case["semantics_dim"],
                            desc_common=case["desc_common"])
        m1 = objects_from_rows(p.D1, "roman", case["semantics_dim"])
        m2 = objects_from_rows(p.D2, "roman", case["semantics_dim"])
        reg = SubmapAlignParams(method="clipper+prune", semantics_dim=case["semantics_dim"],
                                **case["params"]).get_object_registration()
        A = reg.putative_associations(m1, m2)
        assert A.tolist() == case["putative"]
        assert 0 < len(A) < case["n1"] * case["n2"]


def test_yaw_only_projection_equals_scipy_path(gold):
    """rm_roll_pitch against roman.utils.transform_rm_roll_pitch (scipy as_euler('ZYX')[0] -> Rz)."""
    for case in gold["helpers"]["rm_roll_pitch"]:
        got = rm_roll_pitch(np.asarray(case["T"]))
        assert np.allclose(got, np.asarray(case["out"]), rtol=0, atol=1e-12)


def test_similarity_and_aabb_restatements(gold):
    sims = gold["helpers"]["similarity"]
    n_seeded = len(golden_inputs.SIMILARITY_SHAPES)
    for case in sims[:n_seeded]:
        d1, d2 = golden_inputs.similarity_inputs(case["case"])
        assert abs(cn.submap_similarity(d1, d2) - case["similarity"]) < 1e-12
    for case in sims[n_seeded:]:   # 1-d (mean) descriptors, one of them all-zero
        assert abs(cn.submap_similarity(case["desc1"], case["desc2"]) - case["similarity"]) < 1e-12
    assert sims[-1]["similarity"] == 0.0
    hits = 0
    for case in gold["helpers"]["aabb"]:
        p1, p2 = np.asarray(case["p1"]), np.asarray(case["p2"])
        assert cn.aabb_intersects(p1, p2) == case["intersects"]
        assert bool(prefilter.aabb_overlap([p1], [p2])[0, 0]) == case["intersects"]
        hits += case["intersects"]
    assert 0 < hits < len(gold["helpers"]["aabb"])


def test_candidate_mask_follows_the_reference_gates():
    sp = SubmapAlignParams(submap_descriptor="stacked_frame_descriptors", submap_descriptor_thresh=0.8)
    dist = np.array([[1.0, 50.0], [5.0, 5.0]])
    sim = np.array([[0.9, 0.9], [0.79, 0.8]])
    mask = prefilter.candidate_mask(sp, 2, 2, dist=dist, skip_distance=20.0, similarity=sim)
    assert mask.tolist() == [[True, False], [False, True]]       # `>` and `<` are strict in the reference
    assert prefilter.candidate_mask(SubmapAlignParams(), 2, 2, dist=dist, similarity=sim).all()
    assert prefilter.nearby_mask(sp, dist).tolist() == [[True, False], [True, True]]   # < 2 * submap_radius


def test_wire_formats_equal_reference_writer(gold, tmp_path):
    """align.g2o / align.json byte-for-byte against the reference's save_submap_align_results
    (roman/align/results.py:149-197) run on the same result matrices."""
    from roman_b200.align.results import SubmapMeta, loop_closures, write_g2o, write_lc_json
    from roman_b200.align.submap_align import SubmapPairResults
    w = gold["wire_formats"]
    n = np.asarray(w["n_assoc"])
    S0, S1 = n.shape
    meta = []
    for run in range(2):
        traj = np.asarray(w["traj_times"][run])
        meta.append([SubmapMeta(time=t, pose_gravity_aligned=np.asarray(w["pose_gravity_aligned"][run][k]),
                                pose_flu=np.asarray(w["pose_flu"][run][k]), pose_index=int(np.argmin(np.abs(traj - t))))
                     for k, t in enumerate(w["times"][run])])
    res = SubmapPairResults(T_ij_hat_mat=np.asarray(w["T_hat"]), clipper_num_associations=n,
                            clipper_percent_associations=n / 10, associated_objs_mat=[[None] * S1 for _ in range(S0)],
                            status_mat=np.zeros((S0, S1), dtype=np.int32))
    edges = []
    for out in w["outputs"]:
        lcs = loop_closures(res, meta[0], meta[1], lc_association_thresh=out["lc_association_thresh"],
                            single_robot_lc=out["single_robot_lc"],
                            single_robot_lc_time_thresh=out["single_robot_lc_time_thresh"])
        g2o, js = tmp_path / "a.g2o", tmp_path / "a.json"
        write_g2o(str(g2o), lcs, g2o_t_std=out["g2o_t_std"], g2o_r_std=out["g2o_r_std"])
        write_lc_json(str(js), lcs, robot_names=out["robot_names"])
        assert g2o.read_text() == out["g2o"]
        assert js.read_text() == out["json"]
        edges.append(len(lcs))
    assert edges == [7, 2]


def test_pair_error_metrics_equal_reference_loop(gold):
    """The evaluation matrices of the pair loop (relative pose, yaw difference, rotation / translation error incl. the
    reference's sentinel values) recomputed from the reference's own T_ij_hat."""
    from roman_b200.align import prefilter
    from roman_b200.align.submap_align import pair_error_metrics
    for case in gold["pair_loop"]:
        sc = golden_inputs.pair_loop_scenario(case["name"])
        poses = [np.asarray(p) for p in case["poses"]]
        That = np.array([[np.full((4, 4), np.nan) if T is None else np.asarray(T) for T in row] for row in case["T_ij_hat_mat"]])
        dist = prefilter.submap_distances(poses[0][:, :3, 3], poses[1][:, :3, 3])
        nearby = prefilter.nearby_mask(SubmapAlignParams(**sc["sm_kwargs"]), dist)
        attempted = ~(dist > sc["skip_distance"])
        T_ij, ang, dd, yaw = pair_error_metrics(That, poses[0], poses[1], nearby, attempted)
        assert np.abs(T_ij - np.asarray(case["T_ij_mat"])).max() < 1e-12
        for got, key in ((ang, "clipper_angle_mat"), (dd, "clipper_dist_mat"), (yaw, "submap_yaw_diff_mat")):
            want = np.array([[np.nan if v is None else v for v in row] for row in case[key]], dtype=np.float64)
            assert np.array_equal(np.isnan(got), np.isnan(want)), (case["name"], key)
            assert np.nanmax(np.abs(got - want) / np.maximum(1.0, np.abs(want))) < 1e-9, (case["name"], key)


def test_submap_json_equals_reference_writer(gold, tmp_path):
    """`<robot>.sm.json` byte-for-byte against the reference writer (roman/align/results.py:199-246)."""
    from roman_b200.align.results import write_submaps_json
    w = gold["wire_formats"]
    want = w["outputs"][0]["sm_json"]
    assert want is not None
    segs = [dict(id=s["id"], centroid_odom=np.mean(np.asarray(s["points"]), axis=0), volume=s["volume"],
                 linearity=s["linearity"], planarity=s["planarity"], scattering=s["scattering"],
                 first_seen=s["first_seen"], last_seen=s["last_seen"]) for s in w["map_segments"]]
    subs = [dict(time=t, pose_gravity_aligned=np.asarray(w["pose_gravity_aligned"][0][k]),
                 segment_ids=[s["id"] for s in w["map_segments"][:k + 1]]) for k, t in enumerate(w["times"][0])]
    path = tmp_path / "r0.sm.json"
    write_submaps_json(str(path), w["outputs"][0]["robot_names"][0], segs, subs)
    assert path.read_text() == want


def test_g2o_fusion_equals_reference_tool(gold, tmp_path):
    """create_config + g2o_file_fusion byte-for-byte against roman/offline_rpgo/g2o_file_fusion.py on the same input
    files (two odometry files, one multi-robot and two single-robot loop-closure files), with and without the
    association threshold."""
    from roman_b200.align import g2o_fusion as gf
    g = gold["g2o_fusion"]
    for rel, text in g["files"].items():
        dst = tmp_path / rel
        dst.parent.mkdir(parents=True, exist_ok=True)
        dst.write_text(text)
    assert gf.create_config(["r1", "r2"], "ODOM", "ALIGN", "align") == g["config"]
    cfg = gf.create_config(["r1", "r2"], str(tmp_path / "odom"), str(tmp_path / "align"), "align")
    for key, want in g["fused"].items():
        out = tmp_path / f"fused_{key}.g2o"
        gf.g2o_file_fusion(cfg, str(out), thresh=None if key == "None" else int(key))
        assert out.read_text() == want, key
    assert g["fused"]["None"].count("\n") > g["fused"]["5"].count("\n") > 0
    assert gf.gtsam_symbol("a", 7) == (97 << 56) + 7


@pytest.mark.parametrize("name", list(golden_inputs.MAP_CUT_CASES))
def test_submap_cutting_equals_reference_function(gold, name):
    """cut_submaps against roman.map.map.submaps_from_roman_map run on the same map: which submaps exist, their poses,
    which objects they hold IN ORDER, the objects' centres in the submap frame, the submap descriptors."""
    from roman_b200.align.objects import SimpleObject
    from roman_b200.align.submaps import MapData, cut_submaps

    class Seg(SimpleObject):
        pass
    sc = golden_inputs.map_scenario()
    segs = []
    for g in sc["segments"]:
        o = Seg(center=g["center"].copy(), volume=g["volume"], linearity=g["linearity"], planarity=g["planarity"],
                scattering=g["scattering"], extent=g["extent"], semantic_descriptor=g["semantic_descriptor"], id=g["id"])
        o.first_seen, o.last_seen = g["first_seen"], g["last_seen"]
        segs.append(o)
    traj0 = sc["trajectory"].copy()
    subs = cut_submaps(MapData(segments=segs, trajectory=sc["trajectory"], times=sc["times"], descriptors=sc["descriptors"]),
                       SubmapAlignParams(**golden_inputs.MAP_CUT_CASES[name]))
    want = gold["submap_cutting"][name]
    assert [sm.id for sm in subs] == [w["id"] for w in want]
    for sm, w in zip(subs, want):
        assert sm.time == w["time"]
        assert np.abs(sm.pose_gravity_aligned - np.asarray(w["pose_gravity_aligned"])).max() < 1e-12
        assert [s.id for s in sm.segments] == w["segment_ids"], (name, sm.id)
        assert np.abs(np.array([np.asarray(s.center).reshape(-1) for s in sm.segments]) - np.asarray(w["centers"])).max() < 1e-9
        if w["descriptor"] is not None:
            assert np.abs(sm.descriptor - np.asarray(w["descriptor"])).max() < 1e-12
        elif w["descriptor_frames"] is not None:
            assert np.array_equal(sm.descriptor, sc["descriptors"][w["descriptor_frames"]])
        else:
            assert sm.descriptor is None
    assert np.array_equal(traj0, sc["trajectory"]) and np.array_equal(segs[0].center, sc["segments"][0]["center"])   # inputs untouched
