"""
GPU parity against tests/golden/reference_py_vectors.json: the registration classes of roman_b200.align
(CUDA underneath, through the C ABI) must return what the reference's own Python classes returned when
they were run by tests/golden/make_reference_py_golden.py — same start vectors, same inputs.
Bars: association lists identical including order; poses within 1e-9 relative Frobenius of the
reference's numpy-LAPACK result (BASELINE's gate is 1e-4); scores within 1e-9.
"""
import json
import os
import sys

import numpy as np
import pytest

from oracle import clipper_np as cn
from roman_b200 import synth
from roman_b200.align import (InsufficientAssociationsException, ObjectRegistration, SimpleObject, SubmapAlignParams,
                              align_submap_pairs)
from roman_b200.align.objects import objects_from_rows

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import golden_inputs  # noqa: E402


@pytest.fixture(scope="module")
def gold():
    with open(os.path.join(HERE, "golden", "reference_py_vectors.json")) as f:
        return json.load(f)


def rel(T, Tref):
    return float(np.linalg.norm(np.asarray(T) - np.asarray(Tref)) / np.linalg.norm(np.asarray(Tref)))


def test_register_and_t_align_equal_reference_classes(engine, gold):
    for case in gold["registration"]:
        sem = case["semantics_dim"]
        p = synth.make_pair(case["seed"], case["n1"], case["n2"], method=case["method"], semantics_dim=sem)
        m1, m2 = objects_from_rows(p.D1, case["method"], sem), objects_from_rows(p.D2, case["method"], sem)
        reg = SubmapAlignParams(method=case["method"], semantics_dim=sem).get_object_registration()
        u0 = np.random.default_rng(case["u0_seed"]).uniform(0.0, 1.0, case["n1"] * case["n2"])
        assoc = reg.register(m1, m2, u0=u0)
        assert np.asarray(assoc).tolist() == case["associations"], case["method"]
        assert rel(reg.T_align(m1, m2, assoc), case["T"]) < 1e-9, case["method"]
        # the batched entry gives the same answers
        a_b, T_b, st = reg.register_batch([(m1, m2)], u0=[u0])
        assert st.tolist() == [0] and a_b[0].tolist() == case["associations"] and rel(T_b[0], case["T"]) < 1e-9


def test_mno_clipper_equals_reference_loop(engine, gold):
    for case in gold["mno_clipper"]:
        p = synth.make_pair(case["seed"], case["n1"], case["n2"], method=case["method"])
        m1, m2 = objects_from_rows(p.D1, case["method"], 0), objects_from_rows(p.D2, case["method"], 0)
        reg = SubmapAlignParams(method=case["method"]).get_object_registration()
        rng = np.random.default_rng(case["u0_seed"])
        u0s = [rng.uniform(0.0, 1.0, case["n1"] * case["n2"]) for _ in range(case["num_solutions"])]
        sols = reg.mno_clipper(m1, m2, num_solutions=case["num_solutions"], u0s=u0s)
        assert len(sols) == len(case["solutions"])
        for (a, score), want in zip(sols, case["solutions"]):
            assert np.asarray(a).tolist() == want["associations"]
            assert abs(score - want["score"]) < 1e-9


def test_clipper_prune_equals_reference_class(engine, gold):
    for case in gold["clipper_prune"]:
        sem = case["semantics_dim"]
        p = synth.make_pair(case["seed"], case["n1"], case["n2"], method="roman", semantics_dim=sem,
                            desc_common=case["desc_common"])
        m1, m2 = objects_from_rows(p.D1, "roman", sem), objects_from_rows(p.D2, "roman", sem)
        reg = SubmapAlignParams(method="clipper+prune", semantics_dim=sem, **case["params"]).get_object_registration()
        u0 = np.random.default_rng(case["u0_seed"]).uniform(0.0, 1.0, len(case["putative"]))
        assert case["exception"] is None
        assert np.asarray(reg.register(m1, m2, u0=u0)).tolist() == case["associations"]


def test_t_align_kernel_equals_numpy_lapack_path(engine, gold):
    for case in gold["t_align"]:
        reg = ObjectRegistration(dim=case["dim"])
        m1 = [SimpleObject(center=np.asarray(c)) for c in case["pts1"]]
        m2 = [SimpleObject(center=np.asarray(c)) for c in case["pts2"]]
        corr = np.asarray(case["correspondences"])
        if case["exception"]:
            with pytest.raises(InsufficientAssociationsException):
                reg.T_align(m1, m2, corr)
            continue
        T = reg.T_align(m1, m2, corr)
        assert T.shape == (case["dim"] + 1,) * 2
        assert rel(T, case["T"]) < 1e-9, case["name"]


def test_similarity_kernel_equals_reference_similarity(engine, gold):
    sims = gold["helpers"]["similarity"]
    n_seeded = len(golden_inputs.SIMILARITY_SHAPES)
    for case in sims:
        d1, d2 = (golden_inputs.similarity_inputs(case["case"]) if "case" in case
                  else (np.asarray(case["desc1"]), np.asarray(case["desc2"])))
        got = engine.submap_similarity([d1], [d2]).cpu().numpy()
        assert got.shape == (1, 1) and abs(got[0, 0] - case["similarity"]) < 1e-12
    assert len(sims) == n_seeded + 2
    # ragged table, every pair at once, against the numpy restatement
    rng = np.random.default_rng(7)
    f0 = [rng.normal(size=(int(n), 96)) for n in rng.integers(1, 9, 6)]
    f1 = [rng.normal(size=(int(n), 96)) for n in rng.integers(1, 9, 9)]
    f0[2][0] = 0.0
    f1[4] = f0[1][:1] * 3.0                       # an exact match: similarity 1
    got = engine.submap_similarity(f0, f1).cpu().numpy()
    want = np.array([[cn.submap_similarity(a, b) for b in f1] for a in f0])
    assert np.abs(got - want).max() < 1e-12 and abs(got[1, 4] - 1.0) < 1e-12
    with pytest.raises(ValueError):
        engine.submap_similarity([np.zeros((0, 96))], f1)


def test_all_pairs_search_applies_the_prefilters(engine):
    sem, S = 32, 3
    sp = SubmapAlignParams(method="roman", semantics_dim=sem, submap_descriptor="stacked_frame_descriptors",
                           submap_descriptor_thresh=0.8)
    pairs = [synth.make_pair(900 + k, 12, 14, method="roman", semantics_dim=sem) for k in range(S)]
    sub0 = [objects_from_rows(p.D1, "roman", sem) for p in pairs]
    sub1 = [objects_from_rows(p.D2, "roman", sem) for p in pairs]
    rng = np.random.default_rng(5)
    base = [rng.normal(size=(4, 48)) for _ in range(S)]
    desc0 = base
    desc1 = [b[:2] + 0.01 * rng.normal(size=(2, 48)) for b in base]          # submap k of run 1 looks like submap k of run 0
    pos0 = np.array([[0.0, 0, 0], [10.0, 0, 0], [20.0, 0, 0]])
    pos1 = pos0 + np.array([1.0, 0, 0])
    pos1[2] += 500.0                                                          # beyond skip_distance from everything
    u0 = {(k, k): pairs[k].u0 for k in range(S)}
    res = align_submap_pairs(sp, sub0, sub1, u0=u0, positions0=pos0, positions1=pos1, skip_distance=100.0,
                             descriptors0=desc0, descriptors1=desc1)
    want_sim = np.array([[cn.submap_similarity(a, b) for b in desc1] for a in desc0])
    assert np.abs(res.similarity_mat - want_sim)[:, :2].max() < 1e-12 and np.isnan(res.similarity_mat[:, 2]).all()
    assert (res.status_mat[:, 2] == 3).all()                                  # too far: never registered
    assert res.status_mat[0, 1] == 4 and res.status_mat[1, 0] == 4            # dissimilar
    assert res.pairs.tolist() == [[0, 0], [1, 1]]
    assert (res.clipper_num_associations[res.status_mat >= 3] == 0).all()
    assert np.isnan(res.T_ij_hat_mat[0, 2]).all()
    ref = align_submap_pairs(SubmapAlignParams(method="roman", semantics_dim=sem), sub0, sub1, u0=u0,
                             pair_mask=np.eye(S, dtype=bool) & (np.arange(S) < 2)[:, None])
    for k in range(2):
        assert res.status_mat[k, k] == ref.status_mat[k, k] == 0
        assert res.associated_objs_mat[k][k].tolist() == ref.associated_objs_mat[k][k].tolist()
        assert rel(res.T_ij_hat_mat[k, k], pairs[k].T_gt) < 0.1


def test_multistart_equals_best_of_independent_solves(engine, coracle):
    p = synth.make_pair(77, 20, 24, method="gravity")
    m1, m2 = objects_from_rows(p.D1, "gravity", 0), objects_from_rows(p.D2, "gravity", 0)
    reg = SubmapAlignParams(method="gravity").get_object_registration()
    rng = np.random.default_rng(3)
    u0s = [rng.uniform(0.0, 1.0, 480) for _ in range(5)]
    ms = reg.register_multistart(m1, m2, u0s)
    singles = [reg.register(m1, m2, u0=u) for u in u0s]
    assert ms["status"] == 0 and ms["associations"].tolist() == np.asarray(singles[ms["best"]]).tolist()
    ip = cn.InvariantParams(kind=0, sigma=0.4, epsilon=0.6, mindist=0.2, gravity_guided=True, gravity_unc_ang_rad=0.0872665)
    want = [coracle.align_pair(ip, cn.SolverParams(), p.D1, p.D2, u)["score"] for u in u0s]
    assert np.abs(ms["scores"] - np.array(want)).max() < 1e-9 and ms["best"] == int(np.argmax(want))


@pytest.mark.parametrize("name", golden_inputs.PAIR_LOOP_SCENARIOS)
def test_pair_loop_equals_reference_submap_align(engine, gold, name):
    """align_submap_pairs against the result matrices of the reference's own submap_align() pair loop
    (roman/align/submap_align.py:93-200) run on the same submaps with the same start vectors."""
    want = next(c for c in gold["pair_loop"] if c["name"] == name)
    sc = golden_inputs.pair_loop_scenario(name)
    runs = []
    for run in sc["runs"]:
        subs = []
        for sm in run:
            objs = objects_from_rows(sm["rows"], sc["method"], sc["sem"])
            for o, sid in zip(objs, sm["ids"]):
                o.id = sid
            subs.append(objs)
        runs.append(subs)
    pos = [np.array([sm["pose_flu"][:3, 3] for sm in run]) for run in sc["runs"]]
    desc = [[sm["descriptor"] for sm in run] for run in sc["runs"]]
    has_desc = desc[0][0] is not None
    rng = np.random.default_rng(want["u0_seed"])
    sizes = []

    def u0(i, j, m):
        sizes.append(m)
        return rng.uniform(0.0, 1.0, m)
    res = align_submap_pairs(SubmapAlignParams(**sc["sm_kwargs"]), runs[0], runs[1], u0=u0, positions0=pos[0],
                             positions1=pos[1], skip_distance=sc["skip_distance"],
                             descriptors0=desc[0] if has_desc else None, descriptors1=desc[1] if has_desc else None,
                             poses0=np.array([sm["pose_flu"] for sm in sc["runs"][0]]),
                             poses1=np.array([sm["pose_flu"] for sm in sc["runs"][1]]))
    assert sizes == want["u0_sizes"]                       # same pairs reach the solver, in the same order
    assert res.clipper_num_associations.tolist() == want["clipper_num_associations"]
    if want["similarity_mat"] is not None:
        ws = np.asarray(want["similarity_mat"], dtype=np.float64)
        assert np.array_equal(np.isnan(res.similarity_mat), np.isnan(ws))
        assert np.nanmax(np.abs(res.similarity_mat - ws)) < 1e-12
    for row_got, row_want in zip(res.robots_nearby_mat, want["robots_nearby_mat"]):
        for g, w in zip(row_got, row_want):
            assert (np.isnan(g) and w is None) or abs(g - w) < 1e-12
    S0, S1 = res.status_mat.shape
    n_pose = 0
    for i in range(S0):
        for j in range(S1):
            assert res.associated_objs_mat[i][j].tolist() == want["associated_objs_mat"][i][j], (i, j)
            Tw = want["T_ij_hat_mat"][i][j]
            if Tw is None:
                assert np.isnan(res.T_ij_hat_mat[i, j]).all(), (i, j)
            else:
                assert rel(res.T_ij_hat_mat[i, j], Tw) < 1e-9, (i, j)
                n_pose += 1
    assert n_pose >= 1
    # evaluation matrices of the loop (relative pose, yaw difference, registration errors with the reference's sentinels)
    assert np.abs(res.T_ij_mat - np.asarray(want["T_ij_mat"])).max() < 1e-12
    for got, key in ((res.clipper_angle_mat, "clipper_angle_mat"), (res.clipper_dist_mat, "clipper_dist_mat"),
                     (res.submap_yaw_diff_mat, "submap_yaw_diff_mat")):
        w = np.array([[np.nan if v is None else v for v in row] for row in want[key]], dtype=np.float64)
        assert np.array_equal(np.isnan(got), np.isnan(w)), key
        assert np.nanmax(np.abs(got - w) / np.maximum(1.0, np.abs(w))) < 1e-6, key
