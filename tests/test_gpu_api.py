"""
GPU tests through the reference-facing Python surface (clipperpy-compatible shim and the
roman.align mirror classes), against the reference KATs, the committed oracle vectors, and — at
BASELINE.json's full sizes — size-independent properties.
"""
import json
import os

import numpy as np
import pytest

from oracle import clipper_np as cn
from roman_b200 import synth
from tests.helpers import oracle_invariant, params_for_method, rel_fro, same_set

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
T_TOL = 1e-4  # relative Frobenius (BASELINE.json north_star)


def _objects(D, method, sem):
    from roman_b200.align.objects import objects_from_rows
    return objects_from_rows(D, method, sem)


def test_reference_kat_through_clipperpy(engine):
    """affinity_test.cpp:55-107 and clipper_test.cpp:55-66 through the clipperpy-compatible module."""
    from roman_b200 import clipperpy
    K = json.load(open(os.path.join(HERE, "golden", "reference_kats.json")))
    model = np.array(K["scene"]["model"], dtype=float)
    th = np.pi / 8
    R = np.array([[np.cos(th), -np.sin(th), 0], [np.sin(th), np.cos(th), 0], [0, 0, 1.0]])
    data = ((model - np.array(K["scene"]["translation"], dtype=float)) @ R)[:3]
    ip = clipperpy.invariants.EuclideanDistanceParams()
    assert (ip.sigma, ip.epsilon, ip.mindist) == (0.01, 0.06, 0.0)
    c = clipperpy.CLIPPER(clipperpy.invariants.EuclideanDistance(ip), clipperpy.Params())
    c.score_pairwise_consistency(model.T, data.T, np.zeros((0, 2), dtype=np.int32))  # empty A => all-to-all
    A = c.get_initial_associations()
    assert A.shape == (12, 2) and A.dtype == np.int32 and all(tuple(A[i * 3 + j]) == (i, j) for i in range(4) for j in range(3))
    M, C = c.get_affinity_matrix(), c.get_constraint_matrix()
    assert np.array_equal(M, np.array(K["affinity_12x12"], dtype=float)) and np.array_equal(M, C)
    for u0 in (np.ones(12), np.random.default_rng(0).uniform(0, 1, 12)):
        c.solve(u0)
        Ain = c.get_selected_associations()
        assert Ain.shape == (3, 2) and all(a == b for a, b in Ain)
        s = c.get_solution()
        assert abs(s.score - 3.0) < 1e-6 and len(s.u) == 12 and np.array_equal(s.u0, u0)
    c.solve()  # random start, as the reference default: still a clique of the KAT graph
    assert all(M[a, b] == 1 for a in c.get_solution().nodes for b in c.get_solution().nodes)


def test_roman_invariant_through_score_pairwise_consistency(engine, coracle):
    """clipperpy.CLIPPER.score_pairwise_consistency with a ROMAN invariant: pairwise_similarity alone, no single scores
    and no fusion (clipper.cpp:59 calls invariant->operator(); abstract.h:99-102) — what the base
    ObjectRegistration.get_MCA computes with a ROMAN invariant."""
    import copy
    from roman_b200 import clipperpy
    ip_o, _ = params_for_method("roman", semantics_dim=16)
    ip_o = copy.copy(ip_o)
    ip_o.kind = 2
    p = synth.make_pair(91, 11, 13, method="roman", semantics_dim=16, desc_common=1.0)
    rp = clipperpy.invariants.ROMANParams()
    rp.point_dim, rp.ratio_feature_dim, rp.cos_feature_dim = 3, 4, 16
    rp.sigma, rp.epsilon, rp.mindist = ip_o.sigma, ip_o.epsilon, ip_o.mindist
    rp.ratio_epsilon = np.zeros(4)
    rp.cosine_min, rp.cosine_max = ip_o.cosine_min, ip_o.cosine_max
    rp.gravity_guided, rp.gravity_unc_ang_rad = True, ip_o.gravity_unc_ang_rad
    c = clipperpy.CLIPPER(clipperpy.invariants.ROMAN(rp), clipperpy.Params())
    c.score_pairwise_consistency(p.D1.T, p.D2.T, clipperpy.utils.create_all_to_all(11, 13))
    Mo, _, A, _ = coracle.score(ip_o, 1e-4, p.D1, p.D2)
    Mfull = Mo + Mo.T + np.eye(len(A))
    Mg = c.get_affinity_matrix()
    assert np.array_equal(Mg != 0, Mfull != 0) and np.abs(Mg - Mfull).max() <= 1e-12
    c.solve(p.u0)
    ref = coracle.solve(cn.SolverParams(), Mo, p.u0)
    assert c.get_solution().nodes == ref["nodes"]


def test_oracle_vectors_through_registration_classes(engine):
    """Committed fixtures (tests/golden/oracle_vectors.json): ObjectRegistration.register / T_align."""
    from roman_b200.align import SubmapAlignParams
    fx = json.load(open(os.path.join(HERE, "golden", "oracle_vectors.json")))
    for c in fx["cases"]:
        D = c["semantics_dim"]
        p = synth.make_pair(c["seed"], c["n1"], c["n2"], method=c["method"], semantics_dim=D if D else 768,
                            desc_common=c["desc_common"])
        reg = SubmapAlignParams(method=c["method"], semantics_dim=D if D else 768).get_object_registration()
        m1, m2 = _objects(p.D1, c["method"], D), _objects(p.D2, c["method"], D)
        assoc = reg.register(m1, m2, u0=p.u0)
        assert assoc.dtype == np.int32 and assoc.tolist() == c["sel"], c["method"]
        if c["status"] == 0:
            T = reg.T_align(m1, m2, assoc)
            assert T.shape == (4, 4) and rel_fro(T, np.array(c["T"])) <= T_TOL
        else:
            from roman_b200.align import InsufficientAssociationsException
            with pytest.raises(InsufficientAssociationsException):
                reg.T_align(m1, m2, assoc)


def test_register_batch_equals_per_pair_and_oracle(engine, coracle):
    from roman_b200.align import SubmapAlignParams
    reg = SubmapAlignParams(method="roman", semantics_dim=40).get_object_registration()
    ip_o, _ = params_for_method("roman", semantics_dim=40)
    shapes = [(20, 22), (5, 40), (31, 17), (0, 9), (40, 40), (2, 2)]
    pairs_rows = [synth.make_pair(70 + i, max(a, 1), max(b, 1), method="roman", semantics_dim=40, desc_common=1.0)
                  for i, (a, b) in enumerate(shapes)]
    pairs, u0s = [], []
    for (a, b), p in zip(shapes, pairs_rows):
        m1 = _objects(p.D1[:a], "roman", 40)
        m2 = _objects(p.D2[:b], "roman", 40)
        pairs.append((m1, m2))
        u0s.append(np.random.default_rng(a * 100 + b).uniform(0, 1, a * b))
    assoc, T, status = reg.register_batch(pairs, u0s)
    assert T.shape == (len(pairs), 4, 4)
    for b, ((a, bb), p) in enumerate(zip(shapes, pairs_rows)):
        if a == 0 or bb == 0:
            assert len(assoc[b]) == 0 and status[b] == 1 and np.isnan(T[b]).all()
            continue
        ref = coracle.align_pair(ip_o, cn.SolverParams(), p.D1[:a], p.D2[:bb], u0s[b])
        assert np.array_equal(assoc[b], ref["sel"])
        assert status[b] == ref["status"]
        single = reg.register(pairs[b][0], pairs[b][1], u0=u0s[b])
        assert np.array_equal(single, assoc[b])
        if ref["status"] == 0:
            Tn = cn.T_align(p.D1[ref["sel"][:, 0], :3], p.D2[ref["sel"][:, 1], :3])
            assert rel_fro(T[b], Tn) <= T_TOL


def test_mno_clipper_matches_oracle_chain(engine, coracle):
    """mno_clipper (object_registration.py:57-86): second solution after zeroing the first clique in M."""
    from roman_b200.align import SubmapAlignParams
    reg = SubmapAlignParams(method="gravity").get_object_registration()
    ip_o, _ = params_for_method("gravity")
    p = synth.make_pair(21, 18, 20, method="gravity")
    m1, m2 = _objects(p.D1, "gravity", 0), _objects(p.D2, "gravity", 0)
    rng = np.random.default_rng(1)
    u0s = [rng.uniform(0, 1, 360), rng.uniform(0, 1, 360)]
    sols = reg.mno_clipper(m1, m2, num_solutions=2, u0s=u0s)
    sp = cn.SolverParams()
    Mo, _, A, _ = coracle.score(ip_o, sp.affinityeps, p.D1, p.D2)
    Mfull = Mo + Mo.T
    Cup = np.triu((Mfull != 0).astype(float), 1)
    r1 = coracle.solve(sp, np.triu(Mfull, 1), u0s[0], Cupper=Cup)
    assert sols[0][0].tolist() == A[r1["nodes"]].tolist()
    M2 = Mfull.copy()
    M2[np.ix_(r1["nodes"], r1["nodes"])] = 0.0
    r2 = coracle.solve(sp, np.triu(M2, 1), u0s[1], Cupper=Cup)
    assert sols[1][0].tolist() == A[r2["nodes"]].tolist()
    assert sols[0][0].tolist() != sols[1][0].tolist()  # zeroing the first clique's affinities moves the solver on


def test_dist_reg_with_pruning_matches_oracle(engine, coracle):
    from roman_b200.align import DistRegWithPruning, GravityConstraintError
    p = synth.make_pair(33, 26, 30, method="roman", semantics_dim=24, desc_common=1.0)
    m1, m2 = _objects(p.D1, "roman", 24), _objects(p.D2, "roman", 24)
    reg = DistRegWithPruning(sigma=0.4, epsilon=0.6, mindist=0.2, shape_epsilon=0.2, cos_min=0.6, use_gravity=True)
    A = reg.putative_associations(m1, m2)
    assert 0 < len(A) < 26 * 30
    u0 = np.random.default_rng(8).uniform(0, 1, len(A))
    got = reg.register(m1, m2, u0=u0)
    ip = oracle_invariant(kind=1, sigma=0.4, epsilon=0.6, mindist=0.2)
    ref = coracle.align_pair(ip, cn.SolverParams(), p.D1[:, :3].copy(), p.D2[:, :3].copy(), u0, A=A)
    assert np.array_equal(got, ref["sel"])
    # a pose with large roll is rejected by the gravity gate
    bad = [type(o)(**{**o.__dict__}) for o in m2]
    Rx = np.array([[1, 0, 0], [0, 0, -1], [0, 1, 0.0]])
    for o in bad:
        o.center = Rx @ o.center
    with pytest.raises(GravityConstraintError):
        DistRegWithPruning(sigma=0.4, epsilon=0.6, mindist=0.2, cos_min=0.6, use_gravity=True).register(m1, bad, u0=None)


def test_T_align_dim2_and_insufficient(engine):
    from roman_b200.align import ROMANParams, ROMANRegistration, SimpleObject, InsufficientAssociationsException
    rng = np.random.default_rng(12)
    P2 = rng.uniform(-5, 5, (6, 2))
    th = 0.7
    R = np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]])
    P1 = P2 @ R.T + np.array([1.5, -2.0])
    reg = ROMANRegistration(ROMANParams(point_dim=2))
    m1 = [SimpleObject(center=np.r_[p, 0.0]) for p in P1]
    m2 = [SimpleObject(center=np.r_[p, 0.0]) for p in P2]
    corr = np.stack([np.arange(6), np.arange(6)], axis=1)
    T = reg.T_align(m1, m2, corr)
    Tn = cn.T_align(P1, P2)
    assert T.shape == (3, 3) and rel_fro(T, Tn) <= T_TOL
    with pytest.raises(InsufficientAssociationsException):
        reg.T_align(m1, m2, corr[:1])
    with pytest.raises(InsufficientAssociationsException):
        reg.T_align([], m2, corr)


# ---------------------------------------------------------------- full-size properties (BASELINE configs)

def _planted_ok(assoc, n_in):
    return all(a == b and a < n_in for a, b in assoc)


def test_metric_config_m2000_properties(engine, coracle):
    """BASELINE metric config: n1=40, n2=50 (m=2000), roman, D=768 — round trip through the planted
    transform, idempotence, symmetry of M, and bit-exact inliers vs the oracle on a sample."""
    ip_o, ip_g = params_for_method("roman")
    B = 24
    D1, D2, u0, pairs = synth.make_batch(500, B, 40, 50, method="roman", semantics_dim=768)
    engine.set_invariant_params(ip_g)
    engine.plan(D1.shape[2], [40] * B, [50] * B)
    r1 = engine.align_batch(D1.reshape(-1), D2.reshape(-1), u0.reshape(-1))
    T1, st1, n1 = r1.T.cpu().numpy().copy(), r1.status.cpu().numpy().copy(), r1.n_sel.cpu().numpy().copy()
    a1 = [r1.associations(b) for b in range(B)]
    for b, p in enumerate(pairs):
        assert st1[b] == 0 and n1[b] >= 10
        assert _planted_ok(a1[b], p.n_inliers)                       # only planted matches are selected
        assert np.linalg.norm(T1[b] - p.T_gt) < 0.3                  # and the planted SE(3) comes back (5 cm noise)
        R = T1[b][:3, :3]
        assert np.allclose(R @ R.T, np.eye(3), atol=1e-9) and abs(np.linalg.det(R) - 1) < 1e-9
    Mg, Cg = engine.get_dense(3)
    Mg = Mg.cpu().numpy()
    assert np.array_equal(Mg, Mg.T) and np.array_equal(np.diag(Mg), np.ones(2000)) and Mg.min() >= 0 and Mg.max() <= 1
    r2 = engine.align_batch(D1.reshape(-1), D2.reshape(-1), u0.reshape(-1))   # idempotence / run-to-run determinism
    assert np.array_equal(r2.T.cpu().numpy(), T1, equal_nan=True)
    assert all(np.array_equal(r2.associations(b), a1[b]) for b in range(B))
    for b in (0, 7, 23):                                            # oracle on a sample (about 0.1 s each)
        ref = coracle.align_pair(ip_o, cn.SolverParams(), pairs[b].D1, pairs[b].D2, pairs[b].u0)
        assert np.array_equal(a1[b], ref["sel"])
        Tn = cn.T_align(pairs[b].D1[ref["sel"][:, 0], :3], pairs[b].D2[ref["sel"][:, 1], :3])
        assert rel_fro(T1[b], Tn) <= T_TOL


def test_cfg3_batch256_m1024(engine):
    """BASELINE cfg3: 256 pairs, n1=n2=32 (m=1024), roman, D=768."""
    _, ip_g = params_for_method("roman")
    B = 256
    D1, D2, u0, pairs = synth.make_batch(2000, B, 32, 32, method="roman", semantics_dim=768)
    engine.set_invariant_params(ip_g)
    engine.plan(D1.shape[2], [32] * B, [32] * B)
    r = engine.align_batch(D1.reshape(-1), D2.reshape(-1), u0.reshape(-1))
    st, T = r.status.cpu().numpy(), r.T.cpu().numpy()
    ok = 0
    for b, p in enumerate(pairs):
        a = r.associations(b)
        assert _planted_ok(a, p.n_inliers)
        if st[b] == 0:
            ok += np.linalg.norm(T[b] - p.T_gt) < 0.5
    assert ok >= 0.95 * B


def test_cfg4_single_dense_problem_m4032(engine, coracle):
    """BASELINE cfg4: one problem with 63 x 64 = 4032 correspondences, 768-d (solver stress; the
    1-slot shared-memory configuration of K2), checked against the oracle."""
    ip_o, ip_g = params_for_method("roman")
    p = synth.make_pair(4, 63, 64, method="roman", semantics_dim=768, desc_common=0.8)
    engine.set_invariant_params(ip_g)
    engine.plan(p.D1.shape[1], [63], [64])
    r = engine.align_batch(p.D1.reshape(-1), p.D2.reshape(-1), p.u0)
    ref = coracle.align_pair(ip_o, cn.SolverParams(), p.D1, p.D2, p.u0)
    assert np.array_equal(r.associations(0), ref["sel"])
    Tn = cn.T_align(p.D1[ref["sel"][:, 0], :3], p.D2[ref["sel"][:, 1], :3])
    assert rel_fro(r.T[0].cpu().numpy(), Tn) <= T_TOL
    assert abs(float(r.score[0]) - ref["score"]) <= 1e-9


def test_geometry_only_dense_graph_m2000(engine, coracle):
    """Method 'clipper' (no single scores): ~4 % dense affinity, ~40-75 solver passes at m=2000."""
    ip_o, ip_g = params_for_method("clipper")
    p = synth.make_pair(6, 40, 50, method="clipper")
    engine.set_invariant_params(ip_g)
    engine.plan(3, [40], [50])
    r = engine.align_batch(p.D1.reshape(-1), p.D2.reshape(-1), p.u0)
    ref = coracle.align_pair(ip_o, cn.SolverParams(), p.D1, p.D2, p.u0)
    assert np.array_equal(r.associations(0), ref["sel"])
    assert abs(float(r.score[0]) - ref["score"]) <= 1e-9
    assert int(r.n_passes[0]) >= 20


def _world_submaps(seed, S0=3, S1=4, n=18, sem=24):
    """Run 0: S0 random submaps; run 1: transformed, noisy, partially overlapping copies + one unrelated."""
    from roman_b200.align.objects import objects_from_rows
    rng = np.random.default_rng(seed)
    run0, run1, gt = [], [], {}
    for i in range(S0):
        p = synth.make_pair(seed * 10 + i, n + i, n + 2, method="roman", semantics_dim=sem, desc_common=1.0)
        run0.append(p)
    for j in range(S1):
        if j < S0:
            run1.append(("match", run0[j]))
        else:
            run1.append(("rand", synth.make_pair(seed * 10 + 50 + j, n, n + 3, method="roman", semantics_dim=sem, desc_common=1.0)))
    sub0 = [objects_from_rows(p.D1, "roman", sem) for p in run0]
    sub1 = [objects_from_rows(q.D2, "roman", sem) for _, q in run1]
    rows0 = [p.D1 for p in run0]
    rows1 = [q.D2 for _, q in run1]
    return sub0, sub1, rows0, rows1, run0


def test_all_pairs_batched_equals_pair_loop(engine, coracle):
    """align_submap_pairs (shared tables, one launch sequence) == the reference-style per-pair loop:
    register + T_align + upside-down reject + roll/pitch removal (submap_align.py:153-184)."""
    from roman_b200.align import SubmapAlignParams, align_submap_pairs, InsufficientAssociationsException
    from roman_b200.align.submap_align import rm_roll_pitch, roll_pitch
    sem = 24
    sub0, sub1, rows0, rows1, run0 = _world_submaps(3, sem=sem)
    smp = SubmapAlignParams(method="roman", semantics_dim=sem)
    rng = np.random.default_rng(0)
    u0 = {(i, j): rng.uniform(0, 1, len(sub0[i]) * len(sub1[j])) for i in range(len(sub0)) for j in range(len(sub1))}
    mask = np.ones((len(sub0), len(sub1)), dtype=bool)
    mask[1, 3] = False
    res = align_submap_pairs(smp, sub0, sub1, pair_mask=mask, u0=u0)
    ip_o, _ = params_for_method("roman", semantics_dim=sem)
    reg = smp.get_object_registration()
    n_ok = 0
    for i in range(len(sub0)):
        for j in range(len(sub1)):
            if not mask[i, j]:
                assert res.status_mat[i, j] == -1 and np.isnan(res.clipper_num_associations[i, j])
                continue
            ref = coracle.align_pair(ip_o, cn.SolverParams(), rows0[i], rows1[j], u0[(i, j)])
            assoc = reg.register(sub0[i], sub1[j], u0=u0[(i, j)])
            assert np.array_equal(assoc, ref["sel"])
            try:
                T = reg.T_align(sub0[i], sub1[j], assoc)
                roll, pitch = roll_pitch(T)
                if abs(roll) > np.pi / 2 or abs(pitch) > np.pi / 2:
                    raise ValueError
                T = rm_roll_pitch(T)
                assert res.status_mat[i, j] == 0
                assert np.array_equal(res.associated_objs_mat[i][j], assoc)
                assert res.clipper_num_associations[i, j] == len(assoc)
                assert rel_fro(res.T_ij_hat_mat[i, j], T) <= T_TOL
                n_ok += 1
            except (InsufficientAssociationsException, ValueError):
                assert res.status_mat[i, j] in (1, 2)
                assert np.isnan(res.T_ij_hat_mat[i, j]).all() and len(res.associated_objs_mat[i][j]) == 0
    assert n_ok >= 3
    # planted matches come back as yaw-only transforms close to ground truth
    for i in range(3):
        if res.status_mat[i, i] == 0:
            assert np.linalg.norm(res.T_ij_hat_mat[i, i] - run0[i].T_gt) < 0.5


def test_all_pairs_single_robot_common_ids_removed(engine):
    """single_robot_lc: objects whose id occurs in both submaps are dropped before registering the pair."""
    from roman_b200.align import SubmapAlignParams, align_submap_pairs
    sem = 16
    sub0, sub1, rows0, rows1, run0 = _world_submaps(5, S0=2, S1=2, sem=sem)
    for k, o in enumerate(sub0[0]):
        o.id = k
    for k, o in enumerate(sub1[0]):
        o.id = 1000 + k
    sub1[0][2].id = 4      # object 4 of submap (0) and object 2 of submap (0') are "the same segment"
    smp = SubmapAlignParams(method="roman", semantics_dim=sem, single_robot_lc=True)
    u0 = {(0, 0): np.random.default_rng(1).uniform(0, 1, (len(sub0[0]) - 1) * (len(sub1[0]) - 1))}
    res = align_submap_pairs(smp, sub0[:1], sub1[:1], u0=u0)
    a = res.associated_objs_mat[0][0]   # indexes the submaps AFTER the removal, like the reference loop
    reg = smp.get_object_registration()
    m1 = [o for o in sub0[0] if o.id != 4]
    m2 = [o for o in sub1[0] if o.id != 4]
    direct = reg.register(m1, m2, u0=u0[(0, 0)])
    assert len(a) > 0 and np.array_equal(direct, a)
    back1 = np.array([k for k, o in enumerate(sub0[0]) if o.id != 4])
    back2 = np.array([k for k, o in enumerate(sub1[0]) if o.id != 4])
    assert 4 not in back1[a[:, 0]] and 2 not in back2[a[:, 1]]


def test_register_pairs_distributed_single_rank(engine):
    """roman_b200.dist with world size 1 (no process group): same records as register_batch, in pair order."""
    from roman_b200 import dist as rd
    from roman_b200.align import SubmapAlignParams
    reg = SubmapAlignParams(method="roman", semantics_dim=16).get_object_registration()
    pairs, u0s = [], []
    for s in range(5):
        p = synth.make_pair(80 + s, 12 + s, 15, method="roman", semantics_dim=16, desc_common=1.0)
        pairs.append((_objects(p.D1, "roman", 16), _objects(p.D2, "roman", 16)))
        u0s.append(p.u0)
    out = rd.register_pairs_distributed(reg, pairs, u0s, max_k=32)
    assoc, T, status = reg.register_batch(pairs, u0s)
    assert out["status"].tolist() == status.tolist()
    assert out["n_inliers"].tolist() == [len(a) for a in assoc]
    for g in range(5):
        assert np.array_equal(out["inliers"][g], assoc[g])
        assert np.array_equal(out["T"][g], T[g], equal_nan=True)
