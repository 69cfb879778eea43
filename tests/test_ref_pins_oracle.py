"""
CPU tests: oracle/_ref — the REFERENCE's own C++ (dependencies/clipper: clipper.cpp, utils.cpp, dsd.cpp,
invariants/roman.cpp, invariants/euclidean_distance.cpp, compiled unchanged against the Eigen stand-in of
oracle/eigen_standin; recipe `make -C oracle ref`) — against
  (1) the reference's own known-answer tests (tests/golden/reference_kats.json), and
  (2) the C and numpy restatements, on randomised parameter sweeps: single scores incl. NaN, pairwise scores in every
      branch, the three fusions, top-k tie order, k2ij, and whole problems (affinity build -> solve -> selected
      associations) for a fixed u0.
With (2) the restatements' ROMAN invariant, solver iterates and tie order are pinned by the reference's code itself.
"""
import json
import os

import numpy as np
import pytest

from oracle import clipper_np as cn
from oracle import ref as oref
from roman_b200 import synth
from tests.helpers import oracle_invariant, params_for_method

HERE = os.path.dirname(os.path.abspath(__file__))
KATS = json.load(open(os.path.join(HERE, "golden", "reference_kats.json")))

pytestmark = pytest.mark.skipif(not oref.available(), reason="oracle/_ref not built (needs /root/reference at build time)")


@pytest.fixture(scope="module")
def R():
    return oref.RefClipper()


def kat_scene():
    model = np.array(KATS["scene"]["model"], dtype=float)
    th = np.pi / 8
    Rm = np.array([[np.cos(th), -np.sin(th), 0], [np.sin(th), np.cos(th), 0], [0, 0, 1.0]])
    t = np.array(KATS["scene"]["translation"], dtype=float)
    return model, ((model - t) @ Rm)[:3]


def test_ref_reproduces_the_reference_kats(R):
    """affinity_test.cpp:55-107 (12x12 literal, M == C), clipper_test.cpp:55-66 (3 inliers (i, i)), dsd_test.cpp."""
    model, data = kat_scene()
    ip = oracle_invariant(kind=1, **KATS["scene"]["euclidean_params"])
    sp = cn.SolverParams()
    M, C = R.score(ip, sp, model, data)
    lit = np.array(KATS["affinity_12x12"], dtype=float)
    assert np.array_equal(M, lit) and np.array_equal(C, lit)
    for u0 in (np.ones(12), np.random.default_rng(0).uniform(0, 1, 12)):
        r = R.align_pair(ip, sp, model, data, u0, want_T=False)
        assert len(r["sel"]) == KATS["clipper_test"]["n_inliers"] and all(a == b for a, b in r["sel"])
    W = np.array(KATS["dsd_20x20"])
    assert sorted(R.dsd(W)) == sorted(KATS["dsd_nodes"])


def random_invariant(rng, D):
    r = int(rng.integers(0, 5))
    kw = dict(kind=0, point_dim=3, ratio_feature_dim=r, cos_feature_dim=D if rng.uniform() < 0.7 else 0,
              fusion_method=int(rng.integers(0, 3)), sigma=float(rng.uniform(0.1, 0.8)), epsilon=float(rng.uniform(0.2, 1.2)),
              mindist=float(rng.choice([0.0, 0.2, 1.0])), distance_weight=float(rng.choice([1.0, 2.0, 0.5])),
              ratio_weight=float(rng.choice([1.0, 2.0])), cosine_weight=float(rng.choice([1.0, 0.5])),
              ratio_epsilon=rng.uniform(0.0, 0.4, r), cosine_min=float(rng.uniform(0.0, 0.6)),
              cosine_max=float(rng.uniform(0.65, 0.95)), gravity_guided=bool(rng.integers(0, 2)),
              gravity_unc_ang_rad=float(rng.choice([0.0, 0.0872665, 0.3])), drift_aware=bool(rng.integers(0, 2)),
              drift_scale_sigma=bool(rng.integers(0, 2)), drift_scale=float(rng.uniform(0.05, 0.3)))
    if not kw["gravity_guided"]:
        kw["gravity_unc_ang_rad"] = 0.0
    return oracle_invariant(**kw)


def test_scalar_functions_equal_the_restatements(R, coracle):
    """single_similarity (incl. 0/0 ratios -> NaN, cosine at the window edges), pairwise_similarity in the gravity /
    plain / drift-aware / mindist branches, pairwise_single_fusion: bit for bit (same libm, same operation order)."""
    import ctypes as C
    from oracle import to_c_invariant
    rng = np.random.default_rng(11)
    L = coracle.lib
    L.orc_single_similarity.restype = C.c_double
    L.orc_pairwise_similarity.restype = C.c_double
    L.orc_pairwise_single_fusion.restype = C.c_double
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    for trial in range(400):
        D = int(rng.choice([8, 32]))
        ip = random_invariant(rng, D)
        d = 3 + ip.ratio_feature_dim + ip.cos_feature_dim
        rows = [np.ascontiguousarray(np.concatenate([rng.uniform(-3, 3, 3), np.abs(rng.normal(1, 0.5, ip.ratio_feature_dim)),
                                                     rng.normal(size=ip.cos_feature_dim) + 0.8])) for _ in range(4)]
        if ip.ratio_feature_dim and trial % 7 == 0:
            rows[0][3] = rows[1][3] = 0.0            # 0/0 -> NaN ratio (roman.cpp:147-150)
        if ip.cos_feature_dim and trial % 5 == 0:
            rows[1][3 + ip.ratio_feature_dim:] = rows[0][3 + ip.ratio_feature_dim:]   # cosine exactly 1
        cip = to_c_invariant(ip)
        s_ref = R.single_similarity(ip, rows[0], rows[1])
        s_c = float(L.orc_single_similarity(C.byref(cip), dp(rows[0]), dp(rows[1])))
        s_np = float(cn.single_similarity(ip, rows[0][None, :], rows[1][None, :])[0])
        assert (np.isnan(s_ref) and np.isnan(s_c) and np.isnan(s_np)) or (s_ref == s_c and abs(s_ref - s_np) < 1e-14), (trial, s_ref, s_c, s_np)
        p_ref = R.pairwise_similarity(ip, *rows)
        cu, su = np.cos(ip.gravity_unc_ang_rad), np.sin(ip.gravity_unc_ang_rad)
        p_c = float(L.orc_pairwise_similarity(C.byref(cip), C.c_double(cu), C.c_double(su), *[dp(x) for x in rows]))
        assert p_ref == p_c, (trial, p_ref, p_c)
        assert abs(p_ref - float(cn._pairwise_roman(ip, *[r[None, :] for r in rows])[0])) < 1e-14
        a, b, c = float(rng.uniform(0, 1)), float(rng.uniform(0, 1)), float(rng.uniform(0, 1))
        f_ref = R.fusion(ip, a, b, c)
        assert f_ref == float(L.orc_pairwise_single_fusion(C.byref(cip), C.c_double(a), C.c_double(b), C.c_double(c)))
    # EuclideanDistance (strict <, whole-datum norms)
    for trial in range(100):
        ip = oracle_invariant(kind=1, sigma=float(rng.uniform(0.01, 0.5)), epsilon=float(rng.uniform(0.05, 1.0)),
                              mindist=float(rng.choice([0.0, 0.3])))
        rows = [rng.uniform(-2, 2, 3) for _ in range(4)]
        assert abs(R.pairwise_similarity(ip, *rows) - float(cn._pairwise_euclid(ip, *[r[None, :] for r in rows])[0])) < 1e-15


def k2ij_model(k, n):
    """position k of the row-major enumeration of the strict upper triangle (what utils.cpp:87-97 computes in closed form)"""
    i = 0
    while k >= n - 1 - i:
        k -= n - 1 - i
        i += 1
    return i, i + 1 + k


def test_k_largest_and_k2ij_equal_the_restatements(R, coracle):
    rng = np.random.default_rng(3)
    for trial in range(400):
        n = int(rng.integers(1, 40))
        x = rng.integers(0, 4, n).astype(float) * (rng.uniform() < 0.8) + (rng.uniform(size=n) < 0.1)
        k = int(rng.integers(0, n + 1))
        assert R.k_largest(x, k) == coracle.k_largest(x, k) == cn.k_largest(x, k), (x, k)
    for n in (2, 3, 7, 50, 2000):
        ks = range(n * (n - 1) // 2) if n <= 50 else rng.integers(0, n * (n - 1) // 2, 500)
        for k in ks:
            assert R.k2ij(int(k), n) == k2ij_model(int(k), n)


@pytest.mark.parametrize("method,n1,n2,D,dc", [("clipper", 14, 17, 0, 0.0), ("gravity", 22, 19, 0, 0.0),
                                               ("roman", 9, 13, 16, 1.0), ("roman", 25, 30, 64, 1.0),
                                               ("roman", 40, 50, 768, 0.0), ("sevg", 12, 11, 8, 1.0),
                                               ("pcavolgrav", 20, 20, 0, 0.0), ("spv", 10, 7, 0, 0.0)])
def test_whole_problem_equals_the_c_restatement(R, coracle, method, n1, n2, D, dc):
    """Affinity build (pattern and values), solver iterates for a fixed u0 (u, F, ifinal), ordered inlier nodes and the
    selected associations: the reference's code vs oracle/clipper_oracle.c."""
    ip, _ = params_for_method(method, semantics_dim=D if D else 768)
    sp = cn.SolverParams()
    for seed in range(3):
        p = synth.make_pair(70 + seed, n1, n2, method=method, semantics_dim=D if D else 768, desc_common=dc)
        Mr, Cr = R.score(ip, sp, p.D1, p.D2)
        Mo, _, A, _ = coracle.score(ip, sp.affinityeps, p.D1, p.D2)
        Mo_full = Mo + Mo.T + np.eye(n1 * n2)
        assert np.array_equal(Mr, Mo_full), "affinity matrix differs from the reference's"
        assert np.array_equal(Cr, (Mo_full != 0).astype(float))
        r = R.align_pair(ip, sp, p.D1, p.D2, p.u0)
        o = coracle.solve(sp, Mo, p.u0)
        assert r["nodes"] == o["nodes"]
        assert np.array_equal(r["sel"], A[o["nodes"]])
        assert abs(r["score"] - o["score"]) <= 1e-12 and r["ifinal"] == o["ifinal"]
        assert np.abs(r["u"] - o["u"]).max() <= 1e-12
        full = coracle.align_pair(ip, sp, p.D1, p.D2, p.u0)
        assert np.array_equal(full["sel"], r["sel"]) and full["status"] == r["status"]


def test_randomised_parameter_sweep_whole_problems(R, coracle):
    """Random invariant and solver parameters (all fusions, drift-aware, 2 roundings, no rescale, tight iteration
    limits), explicit association lists included."""
    rng = np.random.default_rng(2024)
    for trial in range(40):
        D = int(rng.choice([8, 24]))
        ip = random_invariant(rng, D)
        d = 3 + ip.ratio_feature_dim + ip.cos_feature_dim
        n1, n2 = int(rng.integers(4, 16)), int(rng.integers(4, 16))
        p = synth.make_pair(500 + trial, n1, n2, method="clipper")
        def rows(c, n):
            return np.ascontiguousarray(np.concatenate([c, np.abs(rng.normal(1, 0.3, (n, ip.ratio_feature_dim))),
                                                        rng.normal(size=(n, ip.cos_feature_dim)) + 1.0], axis=1))
        D1, D2 = rows(p.centres1, n1), rows(p.centres2, n2)
        assert D1.shape[1] == d
        sp = cn.SolverParams()
        sp.rounding = int(rng.choice([0, 2]))
        sp.rescale_u0 = bool(rng.integers(0, 2))
        if trial % 4 == 0:
            sp.maxiniters, sp.maxoliters, sp.maxlsiters = 3, 2, 2
        A = None
        if trial % 3 == 0:
            A = cn.create_all_to_all(n1, n2)
            A = np.ascontiguousarray(A[rng.uniform(size=len(A)) < 0.6])
            if len(A) < 3:
                continue
        m = n1 * n2 if A is None else len(A)
        u0 = rng.uniform(0, 1, m)
        r = R.align_pair(ip, sp, D1, D2, u0, A=A, want_T=False)
        o = coracle.align_pair(ip, sp, D1, D2, u0, A=A)
        assert np.array_equal(r["sel"], o["sel"]), trial
        assert abs(r["score"] - o["score"]) <= 1e-12


def test_set_matrix_data_path_equals_the_restatement(R, coracle):
    """CLIPPER::setMatrixData + solve (the mno_clipper path), incl. constraint-only entries (M zeroed, C kept)."""
    rng = np.random.default_rng(9)
    sp = cn.SolverParams()
    for trial in range(10):
        n = int(rng.integers(6, 30))
        U = np.triu((rng.uniform(size=(n, n)) < 0.4) * rng.uniform(0.1, 1.0, (n, n)), 1)
        Cu = (U != 0).astype(float)
        kill = np.triu(rng.uniform(size=(n, n)) < 0.15, 1)
        Mu = np.where(kill, 0.0, U)            # some affinities zeroed, their constraints stay active
        u0 = rng.uniform(0, 1, n)
        r = R.solve_dense(sp, Mu + Mu.T + np.eye(n), u0, Cu + Cu.T + np.eye(n))
        o = coracle.solve(sp, Mu, u0, Cupper=Cu)
        assert r["nodes"] == o["nodes"] and abs(r["score"] - o["score"]) <= 1e-12
        assert np.abs(r["u"] - o["u"]).max() <= 1e-12


def test_roman_invariant_through_score_pairwise_consistency(R, coracle):
    """CLIPPER::scorePairwiseConsistency with a ROMAN invariant evaluates pairwise_similarity alone (abstract.h:99-102):
    no single scores, no fusion — the base ObjectRegistration._clipper_score_all_to_all / get_MCA path.  kind = 2 of
    the restatements against the reference's own code."""
    import copy
    ip, _ = params_for_method("roman", semantics_dim=16)
    sp = cn.SolverParams()
    for seed in range(3):
        p = synth.make_pair(90 + seed, 11, 13, method="roman", semantics_dim=16, desc_common=1.0)
        Mr, Cr = R.score(ip, sp, p.D1, p.D2, pairwise_only=True)
        ip2 = copy.copy(ip)
        ip2.kind = 2
        Mo, _, A, _ = coracle.score(ip2, sp.affinityeps, p.D1, p.D2)
        Mn, _, _ = cn.score(ip2, sp.affinityeps, p.D1, p.D2)
        assert np.array_equal(Mr, Mo + Mo.T + np.eye(len(A)))
        assert np.abs(Mo - Mn).max() < 1e-14
        Mfull, _ = R.score(ip, sp, p.D1, p.D2)          # and it is NOT what the pairwise-and-single entry computes
        assert not np.array_equal(Mr, Mfull)


def k1_screen(ip, ai, aj, bi, bj):
    """numpy restatement of K1's z-gate screen (rb_affinity.cu: make_zgate / zgate_pass): False <=> K1 skips the pair."""
    if not (ip.gravity_guided and not ip.drift_aware):
        return True
    K13 = 0.57735026919                                  # roman.cpp:12, truncated literal
    daz, dbz = ai[2] - aj[2], bi[2] - bj[2]
    eps_z = ip.epsilon
    if ip.gravity_unc_ang_rad > 0.0:
        eps_z = eps_z + abs(0.5 * (abs(daz) + abs(dbz)) * np.sin(ip.gravity_unc_ang_rad))
    return not (abs(daz - dbz) > K13 * eps_z)


def test_k1_z_gate_screen_is_implied_by_the_reference_pairwise_score(R):
    """K1 evaluates the reference's z test (roman.cpp:54-77) BEFORE everything else and skips a pair that fails it.  Against
    the reference's own pairwise_similarity: a skipped pair scores exactly 0 there, in either operand order, and near the
    gate's threshold the screen and the reference agree on which side a pair falls (same fp64 expressions)."""
    rng = np.random.default_rng(23)
    n_skipped = n_kept_nonzero = 0
    for trial in range(3000):
        ip = random_invariant(rng, 8)
        pts = [np.ascontiguousarray(np.concatenate([rng.uniform(-3, 3, 2), rng.uniform(-1.5, 1.5, 1)])) for _ in range(4)]
        if trial % 3 == 0 and ip.gravity_guided and not ip.drift_aware:
            # put the pair within a few ulp of the z threshold
            dz = 0.57735026919 * ip.epsilon
            if ip.gravity_unc_ang_rad == 0.0:
                pts[3][2] = pts[2][2] - ((pts[0][2] - pts[1][2]) - dz) + float(rng.integers(-3, 4)) * np.spacing(dz)
        rows = [np.ascontiguousarray(np.concatenate([p, np.ones(ip.ratio_feature_dim), np.ones(ip.cos_feature_dim)])) for p in pts]
        keep = k1_screen(ip, pts[0], pts[1], pts[2], pts[3])
        assert keep == k1_screen(ip, pts[1], pts[0], pts[3], pts[2])        # operand order does not matter
        p = R.pairwise_similarity(ip, *rows)
        if not keep:
            n_skipped += 1
            assert p == 0.0 and R.pairwise_similarity(ip, rows[1], rows[0], rows[3], rows[2]) == 0.0, (trial, p)
        elif p > 0.0:
            n_kept_nonzero += 1
    assert n_skipped > 200 and n_kept_nonzero > 100
