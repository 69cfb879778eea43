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
