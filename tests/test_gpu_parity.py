"""
GPU parity tests proper: the CUDA path, called through the C ABI, against the CPU oracle on the
same seeded inputs.  Bars (BASELINE.json north_star): identical inlier index sets (bit-exact
integer), SE(3) within 1e-4 relative Frobenius; additionally identical M zero pattern and
|dM| <= 1e-12, |dF| <= 1e-9 (SURVEY §8d "Parity gates").
"""
import numpy as np
import pytest

from oracle import clipper_np as cn
from roman_b200 import synth
from tests.helpers import params_for_method, rel_fro, same_set

pytestmark = pytest.mark.gpu

T_TOL = 1e-4      # relative Frobenius, north_star
M_TOL = 1e-12     # absolute, SURVEY §8d
F_TOL = 1e-9


@pytest.mark.parametrize("method,n1,n2,D", [
    ("clipper", 30, 30, 0),      # BASELINE cfg1: two 30-object submaps, geometry only
    ("gravity", 20, 25, 0),
    ("pcavolgrav", 17, 23, 0),
    ("roman", 12, 14, 64),
    ("roman", 40, 40, 512),      # BASELINE cfg2 shape (synthetic stand-in)
    ("sevg", 9, 33, 32),
])
def test_affinity_and_solve_match_oracle(engine, coracle, method, n1, n2, D):
    ip_o, ip_g = params_for_method(method, semantics_dim=D if D else 768)
    sp = cn.SolverParams()
    for seed in range(3):
        p = synth.make_pair(100 + seed, n1, n2, method=method, semantics_dim=D if D else 768, desc_common=1.0)
        d = p.D1.shape[1]
        engine.set_invariant_params(ip_g)
        engine.plan(d, [n1], [n2])
        engine.affinity(p.D1.reshape(-1), p.D2.reshape(-1))
        Mg, Cg = engine.get_dense(0)
        Mg, Cg = Mg.cpu().numpy(), Cg.cpu().numpy()
        Mo, single, A, nnz = coracle.score(ip_o, sp.affinityeps, p.D1, p.D2)
        Mo_full = Mo + Mo.T + np.eye(n1 * n2)
        assert np.array_equal(Mg != 0, Mo_full != 0), "affinity zero pattern differs"
        assert np.array_equal(Cg, (Mo_full != 0).astype(float))
        assert np.abs(Mg - Mo_full).max() <= M_TOL
        sol = engine.solve(p.u0, want_u=True)
        ref = coracle.solve(sp, Mo, p.u0)
        k = int(sol["n_nodes"][0])
        nodes = sol["nodes"][:k].cpu().numpy().tolist()
        assert nodes == ref["nodes"], "inlier node list (ordered) differs"
        assert abs(float(sol["score"][0]) - ref["score"]) <= F_TOL
        assert np.abs(sol["u"].cpu().numpy() - ref["u"]).max() <= 1e-9
        assert int(sol["ifinal"][0]) == ref["ifinal"]


def test_align_batch_ragged_matches_oracle(engine, coracle):
    """A ragged batch through rb_align_batch: inlier sets bit-exact, T within 1e-4 rel. Frobenius."""
    ip_o, ip_g = params_for_method("roman", semantics_dim=48)
    sp = cn.SolverParams()
    shapes = [(10, 12), (25, 31), (40, 40), (7, 5), (33, 18), (40, 50)]
    pairs = [synth.make_pair(7 + i, a, b, method="roman", semantics_dim=48, desc_common=0.9)
             for i, (a, b) in enumerate(shapes)]
    d = pairs[0].D1.shape[1]
    engine.set_invariant_params(ip_g)
    engine.plan(d, [s[0] for s in shapes], [s[1] for s in shapes])
    D1 = np.concatenate([p.D1.reshape(-1) for p in pairs])
    D2 = np.concatenate([p.D2.reshape(-1) for p in pairs])
    u0 = np.concatenate([p.u0 for p in pairs])
    res = engine.align_batch(D1, D2, u0)
    T = res.T.cpu().numpy()
    st = res.status.cpu().numpy()
    for b, p in enumerate(pairs):
        ref = cn.align_pair(ip_o, sp, p.D1, p.D2, p.u0)
        cref = coracle.align_pair(ip_o, sp, p.D1, p.D2, p.u0)
        assert same_set(cref["sel"], ref["sel"])
        got = res.associations(b)
        assert np.array_equal(got, cref["sel"]), f"problem {b}: inlier associations differ"
        assert st[b] == ref["status"]
        if ref["status"] == 0:
            assert rel_fro(T[b], ref["T"]) <= T_TOL
        else:
            assert np.isnan(T[b]).all()


def test_host_entry_point_matches_device_path(engine):
    ip_o, ip_g = params_for_method("roman", semantics_dim=32)
    B, n1, n2 = 5, 16, 20
    D1, D2, u0, pairs = synth.make_batch(50, B, n1, n2, method="roman", semantics_dim=32, desc_common=1.0)
    engine.set_invariant_params(ip_g)
    engine.plan(D1.shape[2], [n1] * B, [n2] * B)
    dev = engine.align_batch(D1.reshape(-1), D2.reshape(-1), u0.reshape(-1))
    host = engine.align_batch_host(np.ascontiguousarray(D1), np.ascontiguousarray(D2), np.ascontiguousarray(u0))
    assert np.array_equal(host["n_sel"], dev.n_sel.cpu().numpy())
    for b in range(B):
        o = int(engine.offsets[b])
        assert np.array_equal(host["sel"][o:o + host["n_sel"][b]], dev.associations(b))
    assert np.array_equal(host["T"], dev.T.cpu().numpy(), equal_nan=True)


@pytest.mark.parametrize("cluster", [1, 2, 4, 8])
def test_solver_cluster_sizes_agree_with_oracle(engine, coracle, cluster):
    """K2 with 1/2/4/8 CTAs per problem: same ordered inliers as the oracle, F within 1e-9."""
    ip_o, ip_g = params_for_method("gravity")
    sp = cn.SolverParams()
    shapes = [(30, 30), (40, 50), (13, 9), (40, 40)]
    pairs = [synth.make_pair(300 + i, a, b, method="gravity") for i, (a, b) in enumerate(shapes)]
    engine.set_solver_cluster(cluster)
    try:
        engine.set_invariant_params(ip_g)
        engine.plan(3, [s[0] for s in shapes], [s[1] for s in shapes])
        assert engine.solver_cluster == cluster
        D1 = np.concatenate([p.D1.reshape(-1) for p in pairs])
        D2 = np.concatenate([p.D2.reshape(-1) for p in pairs])
        u0 = np.concatenate([p.u0 for p in pairs])
        engine.affinity(D1, D2)
        sol = engine.solve(u0, want_u=True)
        for b, p in enumerate(pairs):
            Mo, _, A, _ = coracle.score(ip_o, sp.affinityeps, p.D1, p.D2)
            ref = coracle.solve(sp, Mo, p.u0)
            o = int(engine.offsets[b])
            k = int(sol["n_nodes"][b])
            assert sol["nodes"][o:o + k].cpu().numpy().tolist() == ref["nodes"]
            assert abs(float(sol["score"][b]) - ref["score"]) <= F_TOL
            assert np.array_equal(sol["sel"][o:o + k].cpu().numpy(), A[ref["nodes"]])
    finally:
        engine.set_solver_cluster(0)


@pytest.mark.parametrize("n1,n2,cluster", [(70, 70, 2), (64, 80, 2), (64, 80, 4), (64, 80, 1), (70, 73, 8)])
def test_largest_problems_every_cluster_size(engine, coracle, n1, n2, cluster):
    """m = 4900 / 5110 / 5120 (the largest m the solver's shared-memory layout takes): a 2-CTA cluster's slice
    of the combined (M x, C x) vectors is 2560 entries — more than the 2304 the round-1 cluster_combine summed."""
    ip_o, ip_g = params_for_method("gravity")
    sp = cn.SolverParams()
    p = synth.make_pair(77 + n2, n1, n2, method="gravity")
    engine.set_solver_cluster(cluster)
    try:
        engine.set_invariant_params(ip_g)
        engine.plan(3, [n1], [n2])
        assert engine.solver_cluster == cluster
        engine.affinity(p.D1.reshape(-1), p.D2.reshape(-1))
        sol = engine.solve(p.u0, want_u=True)
        Mo, _, A, _ = coracle.score(ip_o, sp.affinityeps, p.D1, p.D2)
        ref = coracle.solve(sp, Mo, p.u0)
        k = int(sol["n_nodes"][0])
        assert sol["nodes"][:k].cpu().numpy().tolist() == ref["nodes"]
        assert abs(float(sol["score"][0]) - ref["score"]) <= F_TOL
        assert np.abs(sol["u"].cpu().numpy() - ref["u"]).max() <= 1e-9
    finally:
        engine.set_solver_cluster(0)


@pytest.mark.parametrize("cluster", [1, 2, 4])
def test_resident_csr_split_over_a_cluster(engine, coracle, cluster):
    """~9 thousand edges at m = 2000 (method 'gravity'): more than one SM's shared-memory CSR holds, so with a cluster
    every CTA builds the CSR of ITS tile-rows from K1's edge list and the partial products are combined each pass;
    with one CTA per problem the matrix keeps streaming.  Either way: the oracle's answer."""
    ip_o, ip_g = params_for_method("gravity")
    sp = cn.SolverParams()
    pairs = [synth.make_pair(11 + i, 40, 50, method="gravity") for i in range(3)]
    engine.set_solver_cluster(cluster)
    engine.set_solver_promote(2)
    try:
        engine.set_invariant_params(ip_g)
        engine.plan(3, [40] * 3, [50] * 3)
        engine.affinity(np.concatenate([p.D1.reshape(-1) for p in pairs]), np.concatenate([p.D2.reshape(-1) for p in pairs]))
        sol = engine.solve(np.concatenate([p.u0 for p in pairs]), want_u=True)
        kinds = engine.solver_resident().cpu().numpy()[:3]
        edges = engine.edge_counts().cpu().numpy()[:3]
        for b, p in enumerate(pairs):
            Mo, _, A, nnz = coracle.score(ip_o, sp.affinityeps, p.D1, p.D2)
            assert nnz == edges[b] and 7000 < nnz < 8 * 2016
            ref = coracle.solve(sp, Mo, p.u0)
            o, k = int(engine.offsets[b]), int(sol["n_nodes"][b])
            assert sol["nodes"][o:o + k].cpu().numpy().tolist() == ref["nodes"]
            assert abs(float(sol["score"][b]) - ref["score"]) <= F_TOL
            assert np.abs(sol["u"][o:o + 2000].cpu().numpy() - ref["u"]).max() <= 1e-9
        if cluster > 1:
            assert (kinds == (2 if engine.default_tile_skip else 4)).all(), kinds
    finally:
        engine.set_solver_cluster(0)
        engine.set_solver_promote(8)


def test_shared_matrix_multistart_plan(engine, coracle):
    """rb_plan_shared: K start vectors per submap pair on ONE stored affinity matrix.  Per start: bit-identical to the
    same problem planned on its own; storage: one matrix per pair; K1's live / edge counts reported for the copies are
    the owner's."""
    ip_o, ip_g = params_for_method("roman", semantics_dim=24)
    sp = cn.SolverParams()
    shapes = [(18, 22), (30, 25)]
    K = 4
    pairs = [synth.make_pair(40 + i, a, b, method="roman", semantics_dim=24, desc_common=1.0) for i, (a, b) in enumerate(shapes)]
    rng = np.random.default_rng(8)
    u0s = [[rng.uniform(0, 1, a * b) for _ in range(K)] for (a, b) in shapes]
    D1 = np.concatenate([p.D1 for p in pairs]); D2 = np.concatenate([p.D2 for p in pairs])
    o1 = np.cumsum([0] + [a for a, _ in shapes]); o2 = np.cumsum([0] + [b for _, b in shapes])
    n1 = [shapes[g][0] for g in range(2) for _ in range(K)]; n2 = [shapes[g][1] for g in range(2) for _ in range(K)]
    engine.set_invariant_params(ip_g)
    engine.plan(D1.shape[1], n1, n2, obj1_offsets=[o1[g] for g in range(2) for _ in range(K)],
                obj2_offsets=[o2[g] for g in range(2) for _ in range(K)], total_obj1=len(D1), total_obj2=len(D2),
                share=[g * K for g in range(2) for _ in range(K)])
    nt = lambda m: ((m + 31) // 32) * ((m + 31) // 32 + 1) // 2
    assert engine.packed_bytes == 8192 * sum(nt(a * b) for a, b in shapes)          # one matrix per pair, not K
    shared = engine.align_batch_host(np.ascontiguousarray(D1), np.ascontiguousarray(D2), np.concatenate([u for g in u0s for u in g]))
    Mg, _ = engine.get_dense(K + 1)                                                  # a copy reads its owner's matrix
    Mo, _, _, _ = coracle.score(ip_o, sp.affinityeps, pairs[1].D1, pairs[1].D2)
    assert np.abs(Mg.cpu().numpy() - (Mo + Mo.T + np.eye(len(Mo)))).max() <= M_TOL
    for g, p in enumerate(pairs):
        for k in range(K):
            b = g * K + k
            engine.plan(p.D1.shape[1], [shapes[g][0]], [shapes[g][1]])
            alone = engine.align_batch_host(np.ascontiguousarray(p.D1), np.ascontiguousarray(p.D2), u0s[g][k])
            o = sum(shapes[h][0] * shapes[h][1] * K for h in range(g)) + k * shapes[g][0] * shapes[g][1]
            n = int(shared["n_sel"][b])
            assert n == int(alone["n_sel"][0]) and np.array_equal(shared["sel"][o:o + n], alone["sel"][:n])
            assert shared["score"][b] == alone["score"][0] and np.array_equal(shared["T"][b], alone["T"][0], equal_nan=True)
            ref = coracle.align_pair(ip_o, sp, p.D1, p.D2, u0s[g][k])
            assert np.array_equal(shared["sel"][o:o + n], ref["sel"])


def test_explicit_associations_euclidean(engine, coracle):
    """clipper+prune style: EuclideanDistance invariant on an arbitrary (pruned) association list."""
    rng = np.random.default_rng(5)
    ip_o, ip_g = None, None
    from tests.helpers import oracle_invariant
    from roman_b200.engine import make_invariant_params
    kw = dict(kind=1, sigma=0.4, epsilon=0.6, mindist=0.2)
    ip_o, ip_g = oracle_invariant(**kw), make_invariant_params(**kw)
    sp = cn.SolverParams()
    probs = []
    for s in range(3):
        p = synth.make_pair(900 + s, 25, 28, method="clipper")
        A_all = cn.create_all_to_all(25, 28)
        keep = rng.uniform(size=len(A_all)) < 0.35
        keep[[i * 28 + i for i in range(12)]] = True  # keep the planted matches
        A = np.ascontiguousarray(A_all[keep])
        probs.append((p, A, rng.uniform(0, 1, len(A))))
    engine.set_invariant_params(ip_g)
    engine.plan(3, [25] * 3, [28] * 3, m=[len(A) for _, A, _ in probs])
    D1 = np.concatenate([p.D1.reshape(-1) for p, _, _ in probs])
    D2 = np.concatenate([p.D2.reshape(-1) for p, _, _ in probs])
    A_cat = np.concatenate([A for _, A, _ in probs]).astype(np.int32)
    u0 = np.concatenate([u for _, _, u in probs])
    res = engine.align_batch(D1, D2, u0, A=A_cat)
    for b, (p, A, u) in enumerate(probs):
        ref = coracle.align_pair(ip_o, sp, p.D1, p.D2, u, A=A)
        assert np.array_equal(res.associations(b), ref["sel"])
        refT = cn.T_align(p.D1[ref["sel"][:, 0], :3], p.D2[ref["sel"][:, 1], :3])
        assert rel_fro(res.T[b].cpu().numpy(), refT) <= T_TOL


def test_set_dense_get_dense_and_constraint_only_entries(engine, coracle):
    """set_matrix_data / get_affinity_matrix round trip, incl. C != pattern(M) as mno_clipper makes it."""
    rng = np.random.default_rng(11)
    m = 75
    W = np.triu(rng.uniform(0.2, 1.0, (m, m)) * (rng.uniform(size=(m, m)) < 0.15), 1)
    W[:8, :8] = np.triu(np.full((8, 8), 0.95), 1)  # planted clique
    M = W + W.T + np.eye(m)
    Cfull = (M != 0).astype(float)
    M2 = M.copy()
    idx = np.arange(8)
    M2[np.ix_(idx, idx)] = 0.0  # mno_clipper zeroes solved rows/cols of M but keeps C
    engine.plan(1, [m], [1], m=[m])
    engine.set_dense(0, M2, Cfull)
    Mg, Cg = engine.get_dense(0)
    M2_expect = M2.copy()
    np.fill_diagonal(M2_expect, 1.0)
    assert np.array_equal(Mg.cpu().numpy(), M2_expect)
    assert np.array_equal(Cg.cpu().numpy(), Cfull)
    u0 = rng.uniform(0, 1, m)
    sol = engine.solve(u0, want_u=True)
    sp = cn.SolverParams()
    ref = coracle.solve(sp, np.triu(M2, 1), u0, Cupper=np.triu(Cfull, 1))
    k = int(sol["n_nodes"][0])
    assert sol["nodes"][:k].cpu().numpy().tolist() == ref["nodes"]
    assert abs(float(sol["score"][0]) - ref["score"]) <= F_TOL


def test_topk_tie_semantics_match_heap(engine, coracle):
    """findIndicesOfkLargest (utils.cpp:33-55) tie order, exercised through exact ties in u:
    a graph made of disjoint identical cliques gives exactly equal u entries."""
    m = 40
    M = np.eye(m)
    for c0 in (0, 10, 20):  # three disjoint 4-cliques with identical weights
        idx = np.arange(c0, c0 + 4)
        M[np.ix_(idx, idx)] = 1.0
    engine.plan(1, [m], [1], m=[m])
    engine.set_dense(0, M, None)
    sp = cn.SolverParams()
    for u0 in (np.ones(m), np.linspace(1, 2, m) * 0 + 0.5):
        sol = engine.solve(u0, want_u=True)
        ref = coracle.solve(sp, np.triu(M, 1), u0)
        k = int(sol["n_nodes"][0])
        assert sol["nodes"][:k].cpu().numpy().tolist() == ref["nodes"]


def test_tile_skip_and_dense_formats_agree(engine, coracle):
    """rb_set_tile_skip: the stored matrices are bitwise identical in both formats; the solves differ only in
    summation order (the tile-skip format may run from a shared-memory CSR): same ordered inliers, same outer
    iteration count, u and F within the oracle tolerances."""
    ip_o, ip_g = params_for_method("roman", semantics_dim=32)
    shapes = [(40, 50), (33, 18), (40, 40), (9, 9)]
    pairs = [synth.make_pair(40 + i, a, b, method="roman", semantics_dim=32, desc_common=(0.0 if i % 2 else 1.0))
             for i, (a, b) in enumerate(shapes)]
    D1 = np.concatenate([p.D1.reshape(-1) for p in pairs])
    D2 = np.concatenate([p.D2.reshape(-1) for p in pairs])
    u0 = np.concatenate([p.u0 for p in pairs])
    engine.set_invariant_params(ip_g)
    outs = []
    try:
        for mode in (True, False):
            engine.set_tile_skip(mode)
            engine.plan(pairs[0].D1.shape[1], [s[0] for s in shapes], [s[1] for s in shapes])
            engine.affinity(D1, D2)
            dense = [tuple(t.cpu().numpy() for t in engine.get_dense(b)) for b in range(len(pairs))]
            sol = engine.solve(u0, want_u=True)
            outs.append((dense, {k: (v.cpu().numpy() if v is not None else None) for k, v in sol.items()}))
    finally:
        engine.set_tile_skip(engine.default_tile_skip)
    (d_a, s_a), (d_b, s_b) = outs
    for b in range(len(pairs)):
        assert np.array_equal(d_a[b][0], d_b[b][0]) and np.array_equal(d_a[b][1], d_b[b][1])
    for k in ("n_nodes", "ifinal", "n_passes"):
        assert np.array_equal(s_a[k], s_b[k]), k
    assert np.abs(s_a["score"] - s_b["score"]).max() <= F_TOL
    assert np.abs(s_a["u"] - s_b["u"]).max() <= 1e-9
    for b, p in enumerate(pairs):
        o, n = int(engine.offsets[b]), int(s_a["n_nodes"][b])
        assert np.array_equal(s_a["nodes"][o:o + n], s_b["nodes"][o:o + n])
        Mo, _, _, _ = coracle.score(ip_o, 1e-4, p.D1, p.D2)
        ref = coracle.solve(cn.SolverParams(), Mo, p.u0)
        assert s_a["nodes"][o:o + n].tolist() == ref["nodes"]


@pytest.mark.parametrize("variant", ["nonzero_rounding", "no_rescale", "tight_iters", "loose_tol"])
def test_solver_parameter_variants_match_oracle(engine, coracle, variant):
    """clipper::Params other than the defaults (clipper.h:27-60): Rounding::NONZERO, rescale_u0 = false,
    iteration caps that bite, different tolerances."""
    from roman_b200.engine import default_solver_params
    ip_o, ip_g = params_for_method("gravity")
    sp_o = cn.SolverParams()
    sp_g = default_solver_params()
    if variant == "nonzero_rounding":
        sp_o.rounding, sp_g.rounding = 0, 0
    elif variant == "no_rescale":
        sp_o.rescale_u0, sp_g.rescale_u0 = False, 0
    elif variant == "tight_iters":
        sp_o.maxiniters, sp_g.maxiniters = 3, 3
        sp_o.maxoliters, sp_g.maxoliters = 2, 2
        sp_o.maxlsiters, sp_g.maxlsiters = 2, 2
    else:
        sp_o.tol_u, sp_g.tol_u = 1e-5, 1e-5
        sp_o.tol_F, sp_g.tol_F = 1e-6, 1e-6
        sp_o.beta, sp_g.beta = 0.5, 0.5
    try:
        engine.set_invariant_params(ip_g)
        engine.set_solver_params(sp_g)
        for seed, (n1, n2) in enumerate([(20, 20), (35, 28)]):
            p = synth.make_pair(600 + seed, n1, n2, method="gravity")
            engine.plan(3, [n1], [n2])
            engine.affinity(p.D1.reshape(-1), p.D2.reshape(-1))
            sol = engine.solve(p.u0, want_u=True)
            Mo, _, A, _ = coracle.score(ip_o, sp_o.affinityeps, p.D1, p.D2)
            ref = coracle.solve(sp_o, Mo, p.u0)
            k = int(sol["n_nodes"][0])
            assert sol["nodes"][:k].cpu().numpy().tolist() == ref["nodes"]
            assert abs(float(sol["score"][0]) - ref["score"]) <= F_TOL
            assert int(sol["ifinal"][0]) == ref["ifinal"]
            assert np.abs(sol["u"].cpu().numpy() - ref["u"]).max() <= 1e-9
    finally:
        engine.set_solver_params(default_solver_params())


def test_error_paths_and_limits(engine):
    """Unsupported configurations fail loudly with the library's message; nothing falls back silently."""
    from roman_b200.engine import RomanB200Error, default_solver_params, make_invariant_params
    sp = default_solver_params()
    sp.rounding = 3  # not a Rounding value
    with pytest.raises(RomanB200Error, match="rounding"):
        engine.set_solver_params(sp)
    with pytest.raises(RomanB200Error, match="point_dim"):
        engine.set_invariant_params(make_invariant_params(point_dim=4))
    with pytest.raises(RomanB200Error, match="gravity"):
        engine.set_invariant_params(make_invariant_params(point_dim=2, gravity_guided=True))
    engine.set_invariant_params(make_invariant_params())
    with pytest.raises(RomanB200Error, match="shared-memory"):
        engine.plan(3, [100], [100])  # m = 10000 exceeds the solver's layout
    engine.plan(3, [4], [3])
    with pytest.raises(RomanB200Error, match="before rb_affinity"):
        engine.solve(np.ones(12))
    with pytest.raises(RomanB200Error, match="row length"):
        engine.set_invariant_params(make_invariant_params(cos_feature_dim=8))
        engine.plan(3, [4], [3])
        engine.affinity(np.zeros(12), np.zeros(9))
    engine.set_invariant_params(make_invariant_params())


def test_degenerate_inputs(engine, coracle):
    """Empty maps, single objects, coincident points (mindist), NaN single scores from zero-norm descriptors."""
    ip_o, ip_g = params_for_method("roman", semantics_dim=8)
    engine.set_invariant_params(ip_g)
    sp = cn.SolverParams()
    rng = np.random.default_rng(2)
    p = synth.make_pair(9, 6, 7, method="roman", semantics_dim=8, desc_common=1.0)
    D1, D2 = p.D1.copy(), p.D2.copy()
    D1[1, 7:] = 0.0           # zero-norm descriptor -> cosine 0/0 = NaN -> no edge (roman.cpp:129, clipper.cpp:104-124)
    D2[3, :3] = D2[2, :3]     # coincident objects -> mindist gate
    D1[4, 6] = 0.0            # zero volume -> ratio 0 with epsilon_shape 0: min/max = 0 -> score 0
    shapes = [(0, 3), (1, 1), (6, 7), (2, 0)]
    tabs1 = [D1[:a] for a, _ in shapes]
    tabs2 = [D2[:b] for _, b in shapes]
    engine.plan(D1.shape[1], [a for a, _ in shapes], [b for _, b in shapes])
    u0 = [rng.uniform(0, 1, a * b) for a, b in shapes]
    res = engine.align_batch(np.concatenate([t.reshape(-1) for t in tabs1]), np.concatenate([t.reshape(-1) for t in tabs2]),
                             np.concatenate(u0))
    st = res.status.cpu().numpy()
    for b, (a, bb) in enumerate(shapes):
        if a == 0 or bb == 0:
            assert st[b] == 1 and int(res.n_sel[b]) == 0
            continue
        ref = coracle.align_pair(ip_o, sp, tabs1[b], tabs2[b], u0[b])
        assert np.array_equal(res.associations(b), ref["sel"])
        assert st[b] == ref["status"]
    Mg, _ = engine.get_dense(2)
    Mo, single, _, _ = coracle.score(ip_o, sp.affinityeps, tabs1[2], tabs2[2])
    assert np.isnan(single).any()
    assert np.array_equal(Mg.cpu().numpy() != 0, (Mo + Mo.T + np.eye(42)) != 0)


def test_randomised_invariant_parameter_sweep(engine, coracle):
    """Seeded sweep over methods, sizes, descriptor lengths and every invariant parameter (per-feature shape
    thresholds, cosine window, mindist, gravity uncertainty): affinity and solve against the oracle."""
    from roman_b200.engine import make_invariant_params, method_invariant_kwargs
    from tests.helpers import oracle_invariant
    rng = np.random.default_rng(2024)
    methods = list(synth.METHOD_FLAGS)
    sp = cn.SolverParams()
    n_edges = 0
    for case in range(28):
        method = methods[case % len(methods)]
        n1, n2 = int(rng.integers(3, 23)), int(rng.integers(3, 23))
        D = int(rng.integers(5, 71))
        cmin = float(rng.uniform(0.2, 0.7))
        kw = method_invariant_kwargs(method, sigma=float(rng.uniform(0.05, 0.6)), epsilon=float(rng.uniform(0.1, 1.0)),
                                     mindist=float(rng.choice([0.0, 0.2, 1.5])), cosine_min=cmin,
                                     cosine_max=float(rng.uniform(cmin + 0.05, 1.0)), semantics_dim=D,
                                     gravity_unc_ang_rad=float(rng.choice([0.0, 0.05, 0.3])))
        kw["ratio_epsilon"] = rng.uniform(0.0, 0.6, kw["ratio_feature_dim"])
        ip_o, ip_g = oracle_invariant(**kw), make_invariant_params(**kw)
        p = synth.make_pair(5000 + case, n1, n2, method=method, semantics_dim=D, desc_common=float(rng.uniform(0.0, 1.5)),
                            pos_noise=float(rng.choice([0.02, 0.1])))
        engine.set_invariant_params(ip_g)
        engine.plan(p.D1.shape[1], [n1], [n2])
        engine.affinity(p.D1.reshape(-1), p.D2.reshape(-1))
        Mg = engine.get_dense(0)[0].cpu().numpy()
        Mo, _, _, nnz = coracle.score(ip_o, sp.affinityeps, p.D1, p.D2)
        Mo_full = Mo + Mo.T + np.eye(n1 * n2)
        assert np.array_equal(Mg != 0, Mo_full != 0), (case, method)
        assert np.abs(Mg - Mo_full).max() <= M_TOL, (case, method)
        sol = engine.solve(p.u0)
        ref = coracle.solve(sp, Mo, p.u0)
        k = int(sol["n_nodes"][0])
        assert sol["nodes"][:k].cpu().numpy().tolist() == ref["nodes"], (case, method)
        assert abs(float(sol["score"][0]) - ref["score"]) <= F_TOL
        n_edges += nnz
    assert n_edges > 200          # the sweep is not vacuous
    engine.set_invariant_params(make_invariant_params())


def test_exact_dsd_rounding_matches_reference_semantics(engine, coracle):
    """Rounding::DSD (clipper.cpp:364-371 + dsd.cpp): the support of u comes from the solver kernel, the exact densest
    subgraph from the host epilogue; checked against the exhaustive restatement and, for the KAT literal of
    clipper/test/dsd_test.cpp fed through set_matrix_data, against the reference's expected nodes."""
    import json, os
    from roman_b200 import clipperpy
    from roman_b200.engine import default_solver_params
    ip_o, ip_g = params_for_method("gravity")
    sp_o = cn.SolverParams(rounding=1)
    sp_g = default_solver_params()
    sp_g.rounding = 1
    engine.set_invariant_params(ip_g)
    engine.set_solver_params(sp_g)
    try:
        shapes = [(9, 11), (14, 12), (6, 20)]
        pairs = [synth.make_pair(300 + i, a, b, method="gravity") for i, (a, b) in enumerate(shapes)]
        engine.plan(3, [a for a, _ in shapes], [b for _, b in shapes])
        res = engine.align_batch(np.concatenate([p.D1.reshape(-1) for p in pairs]),
                                 np.concatenate([p.D2.reshape(-1) for p in pairs]), np.concatenate([p.u0 for p in pairs]))
        checked = 0
        for b, p in enumerate(pairs):
            Mo, _, A, _ = coracle.score(ip_o, sp_o.affinityeps, p.D1, p.D2)
            ref = cn.find_dense_clique(sp_o, Mo, None, p.u0)
            assert res.associations(b).tolist() == A[ref["nodes"]].tolist()
            checked += len(ref["nodes"]) >= 3
        assert checked >= 2
    finally:
        engine.set_solver_params(default_solver_params())
    # the reference's DSD literal through the clipperpy surface: solve() with DSD rounding returns a subset of the support
    # whose density equals that of the exact answer on the support
    kats = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_kats.json")))
    M = np.array(kats["dsd_20x20"])
    M = np.triu(M, 1) + np.triu(M, 1).T + np.eye(20)
    params = clipperpy.Params()
    params.rounding = clipperpy.Rounding.DSD
    c = clipperpy.CLIPPER(clipperpy.invariants.PairwiseInvariant(), params)
    c.set_matrix_data(M=M, C=(M != 0).astype(np.float64))
    c.solve(np.full(20, 0.5))
    sol = c.get_solution()
    support = np.nonzero(sol.u > 0)[0].tolist()
    assert list(sol.nodes) == cn.densest_subgraph(M, support)
    assert set(sol.nodes) <= set(kats["dsd_nodes"]) or set(kats["dsd_nodes"]) <= set(support)


def test_exact_dsd_rounding_on_a_broad_support(engine):
    """A support much larger than the answer: no constraints (C all ones), two solver iterations, a planted heavy
    cluster among light random edges.  The exact rounding must return the cluster; checked against the exhaustive
    restatement run on the numpy solver's support."""
    from roman_b200.engine import default_solver_params
    sp_g = default_solver_params()
    sp_g.rounding, sp_g.maxoliters, sp_g.maxiniters = 1, 2, 2
    sp_o = cn.SolverParams(rounding=1, maxoliters=2, maxiniters=2)
    engine.set_solver_params(sp_g)
    try:
        for seed in range(40, 44):
            rng = np.random.default_rng(seed)
            n, kc = 15, 5
            cl = np.sort(rng.choice(n, kc, replace=False))
            Mu = np.zeros((n, n))
            for a in range(n):
                for b in range(a + 1, n):
                    if a in cl and b in cl:
                        Mu[a, b] = rng.uniform(0.7, 1.0)
                    elif rng.uniform() < 0.2:
                        Mu[a, b] = rng.uniform(0.02, 0.2)
            u0 = rng.uniform(0, 1, n)
            ones = np.ones((n, n))
            engine.plan(1, [n], [1], m=[n])
            engine.set_dense(0, Mu + Mu.T + np.eye(n), ones)
            sol = engine.solve(u0)
            k = int(sol["n_nodes"][0])
            got = sol["nodes"][:k].cpu().numpy().tolist()
            ref = cn.find_dense_clique(sp_o, Mu, np.triu(ones, 1), u0)
            assert got == ref["nodes"] == cl.tolist(), seed
    finally:
        engine.set_solver_params(default_solver_params())


@pytest.mark.parametrize("cluster", [1, 2, 4])
def test_late_residency_promotion_in_the_dense_format(engine, coracle, cluster):
    """rb_set_solver_promote: after N streamed passes a still-iterating problem finishes from the shared-memory CSR (built
    from K1's edge list, or extracted from the tiles when there is none).  Same answers as the oracle; the resident /
    streamed-pass reports say what happened; problems whose edges do not fit keep streaming."""
    if engine.default_tile_skip:
        pytest.skip("promotion belongs to the dense format (the tile-skip format is resident from the start)")
    ip_o, ip_g = params_for_method("roman", semantics_dim=40)
    sp = cn.SolverParams()
    shapes = [(12, 14), (30, 31), (40, 50), (9, 40), (33, 18), (40, 40)]
    pairs = [synth.make_pair(70 + i, a, b, method="roman", semantics_dim=40, desc_common=0.9) for i, (a, b) in enumerate(shapes)]
    D1 = np.concatenate([p.D1.reshape(-1) for p in pairs]); D2 = np.concatenate([p.D2.reshape(-1) for p in pairs])
    u0 = np.concatenate([p.u0 for p in pairs])
    engine.set_invariant_params(ip_g)
    engine.set_solver_cluster(cluster)
    try:
        outs = {}
        for promote in (2, 0):
            engine.set_solver_promote(promote)
            engine.plan(pairs[0].D1.shape[1], [a for a, _ in shapes], [b for _, b in shapes])
            engine.affinity(D1, D2)
            sol = engine.solve(u0, want_u=True)
            outs[promote] = ({k: v.cpu().numpy() for k, v in sol.items() if v is not None},
                             engine.solver_resident().cpu().numpy()[:len(pairs)], engine.solver_streamed_passes().cpu().numpy())
        (s2, res2, str2), (s0, res0, str0) = outs[2], outs[0]
        assert (res0 == 0).all() and np.array_equal(str0, s0["n_passes"])
        assert np.array_equal(s2["n_passes"], s0["n_passes"]) and np.array_equal(s2["n_nodes"], s0["n_nodes"])
        long_running = s0["n_passes"] > 2
        # promoted from K1's edge list (code 4): exactly the 2 streamed passes before the promotion, no extraction pass
        assert long_running.any() and (res2[long_running] == 4).all() and (str2[long_running] == 2).all()
        assert np.abs(s2["u"] - s0["u"]).max() <= 1e-9 and np.abs(s2["score"] - s0["score"]).max() <= F_TOL
        for b, p in enumerate(pairs):
            Mo, _, _, _ = coracle.score(ip_o, sp.affinityeps, p.D1, p.D2)
            ref = coracle.solve(sp, Mo, p.u0)
            o, n = int(engine.offsets[b]), int(s2["n_nodes"][b])
            assert s2["nodes"][o:o + n].tolist() == ref["nodes"]
            assert np.abs(s2["u"][o:o + len(p.u0)] - ref["u"]).max() <= 1e-9
        # a matrix handed in by set_matrix_data has no edge list: the promotion extracts the non-zero words from the
        # tiles with one more streamed pass (code 3)
        b = int(np.argmax(s0["n_passes"]))
        p = pairs[b]
        Mo, _, _, _ = coracle.score(ip_o, sp.affinityeps, p.D1, p.D2)
        ref = coracle.solve(sp, Mo, p.u0)
        engine.set_solver_promote(2)
        engine.plan(p.D1.shape[1], [shapes[b][0]], [shapes[b][1]])
        engine.set_dense(0, Mo + Mo.T + np.eye(len(p.u0)))
        sol = engine.solve(p.u0, want_u=True)
        k = int(sol["n_nodes"][0])
        assert sol["nodes"][:k].cpu().numpy().tolist() == ref["nodes"]
        assert np.abs(sol["u"].cpu().numpy() - ref["u"]).max() <= 1e-9
        assert int(engine.solver_resident()[0]) == 3 and int(engine.solver_streamed_passes()[0]) == 3
        # a geometry-only graph at m = 2000 overflows the edge list and is far too dense for the shared-memory CSR:
        # the extraction fails, streaming goes on
        ip_o2, ip_g2 = params_for_method("clipper")
        p = synth.make_pair(3, 40, 50, method="clipper")
        engine.set_invariant_params(ip_g2)
        engine.set_solver_promote(3)
        engine.plan(3, [40], [50])
        engine.affinity(p.D1.reshape(-1), p.D2.reshape(-1))
        sol = engine.solve(p.u0)
        Mo, _, _, nnz = coracle.score(ip_o2, sp.affinityeps, p.D1, p.D2)
        ref = coracle.solve(sp, Mo, p.u0)
        k = int(sol["n_nodes"][0])
        assert nnz > 8 * 2016 and sol["nodes"][:k].cpu().numpy().tolist() == ref["nodes"]
        assert int(engine.solver_resident()[0]) == 0
        assert int(engine.solver_streamed_passes()[0]) == int(sol["n_passes"][0]) + 1      # + the failed extraction pass
    finally:
        engine.set_solver_promote(8)
        engine.set_solver_cluster(0)
        engine.plan(3, [1], [1])
