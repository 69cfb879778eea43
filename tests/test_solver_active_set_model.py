"""
CPU test of the ARGUMENT behind K2's active-association lists (rb_solver.cu: build_active), on the numpy restatement
of the reference solver (oracle/clipper_np.find_dense_clique, clipper.cpp:243-394).

Claim: an association without any edge whose u_i is exactly zero at the top of an outer iteration stays exactly zero
for the rest of the solve and contributes exactly nothing to any quantity of the algorithm, so all vector phases may
skip it.  The model below is the restatement with that rule applied (vectors reduced to the active set once per outer
iteration); it must return the same nodes, ifinal, pass count and — up to the re-association of floating-point sums
that dropping zero terms causes — the same u and F.  The CUDA kernel itself is compared with the oracle in the GPU
tests; this file pins the reasoning on the CPU.
"""
import numpy as np
import pytest

from oracle import clipper_np as cn
from roman_b200 import synth
from tests.helpers import params_for_method


def solve_active_set(sp, Mupper, u0):
    Ms = np.triu(Mupper, 1)
    Ms = Ms + Ms.T
    Cs = (Ms != 0).astype(np.float64)
    has_edge = (Ms != 0).any(axis=1)
    n = Ms.shape[0]
    u0 = np.asarray(u0, dtype=np.float64)
    passes = 0
    u = Ms @ u0 + u0 if sp.rescale_u0 else u0.copy()
    passes += 1 if sp.rescale_u0 else 0
    u = u / np.sqrt(u @ u)
    Cbu = (u.sum() - Cs @ u) - u
    idx = (Cbu > sp.eps) & (u > sp.eps)
    passes += 1
    d = float(np.mean((Ms @ u + u)[idx] / Cbu[idx])) if idx.sum() > 0 else 0.0
    F, i = 0.0, 0
    dropped_ever = np.zeros(n, dtype=bool)
    while i < sp.maxoliters:
        if d >= 0.0 and u.sum() >= 0.0 and sp.eps >= 0.0:       # build_active's conditions
            act = np.nonzero(has_edge | (u != 0.0))[0]
        else:
            act = np.arange(n)
        dropped_ever |= ~np.isin(np.arange(n), act)
        A, Ca, ua = Ms[np.ix_(act, act)], Cs[np.ix_(act, act)], u[act]
        g = (((1 + d) * ua - d * ua.sum()) + A @ ua) + (Ca @ ua) * d
        passes += 1
        F = float(ua @ g)
        for _j in range(sp.maxiniters):
            alpha, Fn, dF = 1.0, 0.0, 0.0
            for _k in range(sp.maxlsiters):
                un = np.maximum(ua + alpha * g, 0.0)
                z = float(un @ un)
                if z > 0:
                    un = un / np.sqrt(z)
                gn = (((1 + d) * un - d * un.sum()) + A @ un) + (Ca @ un) * d
                passes += 1
                Fn = float(un @ gn)
                dF = Fn - F
                if dF < -sp.eps:
                    alpha *= sp.beta
                else:
                    break
            du = float(np.sqrt(((un - ua) ** 2).sum()))
            F, ua, g = Fn, un, gn
            if du < sp.tol_u or abs(dF) < sp.tol_F:
                break
        u = np.zeros(n)
        u[act] = ua
        Cbu = (u.sum() - Cs @ u) - u
        idx = (Cbu > sp.eps) & (u > sp.eps)
        passes += 1
        if idx.sum() > 0:
            d += float(np.mean(np.abs((Ms @ u + u)[idx] / Cbu[idx])))
        else:
            break
        i += 1
    omega = int(np.floor(abs(F) + 0.5) * (1 if F >= 0 else -1))
    return dict(nodes=cn.k_largest(u, omega), u=u, score=F, ifinal=i, passes=passes, dropped=int(dropped_ever.sum()))


@pytest.mark.parametrize("seed,n1,n2,dc", [(0, 12, 14, 1.0), (1, 16, 16, 1.0), (2, 10, 20, 0.6), (3, 14, 12, 0.0), (4, 18, 15, 1.0)])
def test_dropping_inactive_associations_changes_nothing(seed, n1, n2, dc):
    sem = 16
    ip, _ = params_for_method("roman", semantics_dim=sem)
    sp = cn.SolverParams()
    p = synth.make_pair(seed=seed, n1=n1, n2=n2, semantics_dim=sem, desc_common=dc)
    Mu, _, _ = cn.score(ip, sp.affinityeps, p.D1, p.D2)     # strict upper triangle, dense
    ref = cn.find_dense_clique(sp, Mu, None, p.u0)
    got = solve_active_set(sp, Mu, p.u0)
    assert got["nodes"] == ref["nodes"]
    assert got["ifinal"] == ref["ifinal"] and got["passes"] == ref["passes"]
    assert np.allclose(got["u"], ref["u"], rtol=0, atol=1e-12) and abs(got["score"] - ref["score"]) <= 1e-11
    # every association the model ever dropped ends at exactly zero in the reference's own iterate
    no_edge = ~(np.triu(Mu, 1) + np.triu(Mu, 1).T != 0).any(axis=1)
    assert (ref["u"][no_edge & (got["u"] == 0.0)] == 0.0).all()
    if dc > 0:
        assert got["dropped"] > 0      # the cases are not vacuous
