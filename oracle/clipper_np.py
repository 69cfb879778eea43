"""
clipper_np.py — independent numpy fp64 restatement of the roman.align / CLIPPER hot path.

TEST INFRASTRUCTURE ONLY (see oracle/clipper_oracle.c header): nothing under roman_b200/ may
import this module.  It exists to cross-validate the C restatement (a second, differently
structured reading of the same reference lines) and to check the CUDA path on small and
medium problems.  It keeps numpy's LAPACK SVD for T_align exactly as the reference does.

Reference lines followed (paths relative to /root/reference; clipper/ = dependencies/clipper/):
  single_similarity      clipper/src/invariants/roman.cpp:116-164
  pairwise_similarity    clipper/src/invariants/roman.cpp:17-85
  pairwise_single_fusion clipper/src/invariants/roman.cpp:87-114
  euclidean              clipper/src/invariants/euclidean_distance.cpp:13-31
  score_*                clipper/src/clipper.cpp:29-73, 77-136
  find_dense_clique      clipper/src/clipper.cpp:243-394
  k_largest              clipper/src/utils.cpp:33-55
  T_align                roman/align/object_registration.py:88-129
  rm_roll_pitch          roman/utils.py:128-130; upside-down gate roman/align/submap_align.py:167-170
  submap_similarity      roman/map/map.py:145-162
  aabb_intersects        roman/utils.py:160-169

Parity pin status: identical to clipper_oracle.c — pinned by the reference KATs for the
Euclidean affinity build and the end-to-end solve; T_align, submap_similarity and aabb_intersects
are pinned by outputs of the reference's own Python functions (tests/golden/reference_py_vectors.json,
made by tests/golden/make_reference_py_golden.py); PARITY UNPINNED for the ROMAN invariant, the
solver iterates and the top-k tie order (C++ code that cannot be built here; no golden vector
exists in the reference).
"""
from __future__ import annotations

import heapq
from dataclasses import dataclass, field

import numpy as np

SQRT_TWO_THIRDS = 0.81649658092  # roman.cpp:11 (truncated literal, verbatim)
SQRT_ONE_THIRD = 0.57735026919   # roman.cpp:12


@dataclass
class InvariantParams:
    """clipper/include/clipper/invariants/roman.h:27-53 (+ kind selector)."""
    kind: int = 0  # 0 ROMAN (pairwise + single), 1 EuclideanDistance, 2 ROMAN through scorePairwiseConsistency (pairwise only)
    point_dim: int = 3
    ratio_feature_dim: int = 0
    cos_feature_dim: int = 0
    fusion_method: int = 0
    sigma: float = 0.01
    epsilon: float = 0.06
    mindist: float = 0.0
    distance_weight: float = 1.0
    ratio_weight: float = 1.0
    cosine_weight: float = 1.0
    ratio_epsilon: np.ndarray = field(default_factory=lambda: np.zeros(0))
    cosine_min: float = 0.85
    cosine_max: float = 0.95
    gravity_guided: bool = False
    gravity_unc_ang_rad: float = 0.0
    drift_aware: bool = False
    drift_scale_sigma: bool = False
    drift_scale: float = 0.1


@dataclass
class SolverParams:
    """clipper/include/clipper/clipper.h:27-60."""
    tol_u: float = 1e-8
    tol_F: float = 1e-9
    tol_Fop: float = 1e-10
    maxiniters: int = 200
    maxoliters: int = 1000
    beta: float = 0.25
    maxlsiters: int = 99
    eps: float = 1e-9
    affinityeps: float = 1e-4
    rescale_u0: bool = True
    rounding: int = 2  # 0 NONZERO, 1 DSD, 2 DSD_HEU


def create_all_to_all(n1: int, n2: int) -> np.ndarray:
    """clipper/include/clipper/utils.h:61-71"""
    i, j = np.divmod(np.arange(n1 * n2, dtype=np.int64), n2)
    return np.stack([i, j], axis=1).astype(np.int32)


def single_similarity(p: InvariantParams, a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """Vectorised roman.cpp:116-164.  a, b: (m, d) rows of the associated objects."""
    m = a.shape[0]
    if p.cos_feature_dim == 0 and p.ratio_feature_dim == 0:
        return np.ones(m)
    zero = np.zeros(m, dtype=bool)
    cos_scaled = np.zeros(m)
    if p.cos_feature_dim > 0:
        o = p.point_dim + p.ratio_feature_dim
        fa, fb = a[:, o:o + p.cos_feature_dim], b[:, o:o + p.cos_feature_dim]
        c = np.einsum("ij,ij->i", fa, fb) / (np.linalg.norm(fa, axis=1) * np.linalg.norm(fb, axis=1))
        cos_scaled = np.where(c >= p.cosine_max, 1.0, (c - p.cosine_min) / (p.cosine_max - p.cosine_min))
        zero |= ~(c >= p.cosine_max) & (c <= p.cosine_min)
        if p.ratio_feature_dim == 0:
            return np.where(zero, 0.0, cos_scaled)
    ra, rb = a[:, p.point_dim:p.point_dim + p.ratio_feature_dim], b[:, p.point_dim:p.point_dim + p.ratio_feature_dim]
    with np.errstate(divide="ignore", invalid="ignore"):
        r = np.where(ra < rb, ra / rb, rb / ra)
        below = (r < np.asarray(p.ratio_epsilon, dtype=np.float64)[None, :]).any(axis=1)
        ratio_score = np.power(np.prod(r, axis=1), 1.0 / p.ratio_feature_dim)
        if p.cos_feature_dim == 0:
            return np.where(below, 0.0, ratio_score)
        both = np.power(np.power(ratio_score, p.ratio_weight) * np.power(cos_scaled, p.cosine_weight),
                        1.0 / (p.ratio_weight + p.cosine_weight))
    return np.where(zero | below, 0.0, both)


def _pairwise_roman(p: InvariantParams, ai, aj, bi, bj) -> np.ndarray:
    """Vectorised roman.cpp:17-85 over broadcastable (..., d) arrays."""
    pd = p.point_dim
    l1 = np.sqrt(((ai[..., :pd] - aj[..., :pd]) ** 2).sum(-1))
    l2 = np.sqrt(((bi[..., :pd] - bj[..., :pd]) ** 2).sum(-1))
    dead = np.zeros(l1.shape, dtype=bool)
    if p.mindist > 0:
        dead |= (l1 < p.mindist) | (l2 < p.mindist)
    eps = np.maximum(p.epsilon, p.epsilon * p.drift_scale * 0.5 * (l1 + l2)) if p.drift_aware else p.epsilon
    sig = np.maximum(p.sigma, p.sigma * p.drift_scale * 0.5 * (l1 + l2)) if (p.drift_aware and p.drift_scale_sigma) else p.sigma
    if p.gravity_guided:
        xy1 = np.sqrt(((ai[..., :2] - aj[..., :2]) ** 2).sum(-1))
        xy2 = np.sqrt(((bi[..., :2] - bj[..., :2]) ** 2).sum(-1))
        z1 = ai[..., 2] - aj[..., 2]
        z2 = bi[..., 2] - bj[..., 2]
        c_xy = np.abs(xy1 - xy2)
        c_z = np.abs(z1 - z2)
        s_xy = s_z = sig
        e_xy = e_z = eps
        if p.gravity_unc_ang_rad > 0.0:
            cu, su = np.cos(p.gravity_unc_ang_rad), np.sin(p.gravity_unc_ang_rad)
            xym = 0.5 * (xy1 + xy2)
            zm = 0.5 * (np.abs(z1) + np.abs(z2))
            s_xy = sig + np.abs(xym * cu - xym)
            s_z = sig + np.abs(zm * su)
            e_xy = eps + np.abs(xym * cu - xym)
            e_z = eps + np.abs(zm * su)
        dead |= (c_xy > SQRT_TWO_THIRDS * e_xy) | (c_z > SQRT_ONE_THIRD * e_z)
        val = np.exp(-0.5 * (c_xy * c_xy / (2.0 / 3.0 * s_xy * s_xy) + c_z * c_z / (s_z * s_z / 3.0)))
    else:
        c = np.abs(l1 - l2)
        dead |= c > eps
        val = np.exp(-0.5 * c * c / (sig * sig))
    return np.where(dead, 0.0, val)


def _pairwise_euclid(p: InvariantParams, ai, aj, bi, bj) -> np.ndarray:
    """Vectorised euclidean_distance.cpp:13-31 (whole-datum norms, strict c < epsilon)."""
    l1 = np.sqrt(((ai - aj) ** 2).sum(-1))
    l2 = np.sqrt(((bi - bj) ** 2).sum(-1))
    dead = np.zeros(l1.shape, dtype=bool)
    if p.mindist > 0:
        dead |= (l1 < p.mindist) | (l2 < p.mindist)
    c = np.abs(l1 - l2)
    val = np.where(c < p.epsilon, np.exp(-0.5 * c * c / (p.sigma * p.sigma)), 0.0)
    return np.where(dead, 0.0, val)


def _fusion(p: InvariantParams, pair, si, sj):
    """roman.cpp:87-114"""
    if p.ratio_feature_dim > 0 or p.cos_feature_dim > 0:
        if p.fusion_method == 0:
            return np.power(np.power(pair, p.distance_weight) * si * sj, 1.0 / (p.distance_weight + 2.0))
        if p.fusion_method == 1:
            return (p.distance_weight * pair + si + sj) / (p.distance_weight + 2.0)
        if p.fusion_method == 2:
            return pair * si * sj
        return np.zeros_like(pair)
    return pair


def score(p: InvariantParams, affinityeps: float, D1: np.ndarray, D2: np.ndarray, A: np.ndarray | None = None,
          row_chunk: int = 256):
    """
    clipper.cpp:29-73 / 77-136.  D1 (n1, d), D2 (n2, d) object-major; A (m, 2) or None.
    Returns (Mupper, single, A): Mupper is the m x m dense matrix holding the strict upper
    triangle (zeros elsewhere) — i.e. the dense scratch before sparseView().
    """
    D1 = np.asarray(D1, dtype=np.float64)
    D2 = np.asarray(D2, dtype=np.float64)
    if A is None or np.asarray(A).size == 0:
        A = create_all_to_all(D1.shape[0], D2.shape[0])
    A = np.asarray(A)
    m = A.shape[0]
    a = D1[A[:, 0]]
    b = D2[A[:, 1]]
    single = single_similarity(p, a, b) if p.kind == 0 else np.ones(m)
    M = np.zeros((m, m))
    for r0 in range(0, m, row_chunk):
        r1 = min(m, r0 + row_chunk)
        ai, bi = a[r0:r1, None, :], b[r0:r1, None, :]
        aj, bj = a[None, :, :], b[None, :, :]
        with np.errstate(invalid="ignore", divide="ignore"):
            if p.kind == 0:
                pw = _pairwise_roman(p, ai, aj, bi, bj)
                sc = _fusion(p, pw, single[r0:r1, None], single[None, :])
                gate = ~((single[r0:r1, None] < affinityeps) | (single[None, :] < affinityeps))
            elif p.kind == 2:  # ROMAN::operator() = pairwise_similarity (abstract.h:99-102), clipper.cpp:59
                sc = _pairwise_roman(p, ai, aj, bi, bj)
                gate = np.ones(sc.shape, dtype=bool)
            else:
                sc = _pairwise_euclid(p, ai, aj, bi, bj)
                gate = np.ones(sc.shape, dtype=bool)
            distinct = (A[r0:r1, None, 0] != A[None, :, 0]) & (A[r0:r1, None, 1] != A[None, :, 1])
            upper = np.arange(r0, r1)[:, None] < np.arange(m)[None, :]
            keep = distinct & gate & upper & (sc > affinityeps)
        M[r0:r1] = np.where(keep, sc, 0.0)
    return M, single, A


def k_largest(x: np.ndarray, k: int) -> list[int]:
    """utils.cpp:33-55 — heap of (value, index) pairs, strict-less replacement."""
    if k < 1:
        return []
    k = min(k, len(x))
    q: list[tuple[float, int]] = []
    for i, v in enumerate(x):
        v = float(v)
        if len(q) < k:
            heapq.heappush(q, (v, i))
        elif q[0][0] < v:
            heapq.heapreplace(q, (v, i))
    out = [0] * k
    for i in range(k):
        out[k - i - 1] = heapq.heappop(q)[1]
    return out


def find_dense_clique(sp: SolverParams, Mupper: np.ndarray, Cupper: np.ndarray | None, u0: np.ndarray):
    """
    clipper.cpp:243-394 on dense symmetric matrices (strict upper triangles given).
    Returns dict(nodes, u, score, ifinal, passes).
    """
    Ms = np.triu(Mupper, 1)
    Ms = Ms + Ms.T
    if Cupper is None:
        Cs = (Ms != 0).astype(np.float64)
    else:
        Cs = np.triu(Cupper, 1)
        Cs = Cs + Cs.T
    n = Ms.shape[0]
    u0 = np.asarray(u0, dtype=np.float64)
    passes = 0
    if sp.rescale_u0:
        u = Ms @ u0 + u0
        passes += 1
    else:
        u = u0.copy()
    u = u / np.sqrt(u @ u)

    def cbu_idx(u):
        Cbu = (u.sum() - Cs @ u) - u
        return Cbu, (Cbu > sp.eps) & (u > sp.eps)

    d = 0.0
    Cbu, idx = cbu_idx(u)
    passes += 1
    if idx.sum() > 0:
        Mu = Ms @ u + u
        d = float(np.mean(Mu[idx] / Cbu[idx]))

    F = 0.0
    i = 0
    while i < sp.maxoliters:
        g = (((1 + d) * u - d * u.sum()) + Ms @ u) + (Cs @ u) * d
        passes += 1
        F = float(u @ g)
        for _j in range(sp.maxiniters):
            alpha = 1.0
            Fn = 0.0
            dF = 0.0
            for _k in range(sp.maxlsiters):
                un = np.maximum(u + alpha * g, 0.0)
                z = float(un @ un)
                if z > 0:
                    un = un / np.sqrt(z)
                gn = (((1 + d) * un - d * un.sum()) + Ms @ un) + (Cs @ un) * d
                passes += 1
                Fn = float(un @ gn)
                dF = Fn - F
                if dF < -sp.eps:
                    alpha *= sp.beta
                else:
                    break
            du = float(np.sqrt(((un - u) ** 2).sum()))
            F, u, g = Fn, un, gn
            if du < sp.tol_u or abs(dF) < sp.tol_F:
                break
        Cbu, idx = cbu_idx(u)
        passes += 1
        if idx.sum() > 0:
            Mu = Ms @ u + u
            d += float(np.mean(np.abs(Mu[idx] / Cbu[idx])))
        else:
            break
        i += 1
    if sp.rounding == 0:
        nodes = [int(t) for t in np.nonzero(u > 0.0)[0]]
    elif sp.rounding == 2:
        # std::round: half away from zero
        omega = int(np.floor(abs(F) + 0.5) * (1 if F >= 0 else -1))
        nodes = k_largest(u, omega)
    else:  # DSD (clipper.cpp:364-371): exact densest subgraph of the graph induced by the support of u
        support = [int(t) for t in np.nonzero(u > 0.0)[0]]
        if len(support) > 16:
            raise NotImplementedError("the exhaustive DSD restatement handles supports of at most 16 nodes")
        nodes = densest_subgraph(Mupper, support) if support else []
    return dict(nodes=nodes, u=u, score=F, ifinal=i, passes=passes)


def T_align(pts1: np.ndarray, pts2: np.ndarray) -> np.ndarray:
    """object_registration.py:110-129 on already-gathered matched points (k, dim), unit weights."""
    pts1 = np.asarray(pts1, dtype=np.float64)
    pts2 = np.asarray(pts2, dtype=np.float64)
    weights = np.ones((pts1.shape[0], 1))
    mean1 = (np.sum(pts1 * weights, axis=0) / np.sum(weights)).reshape(-1)
    mean2 = (np.sum(pts2 * weights, axis=0) / np.sum(weights)).reshape(-1)
    p1 = pts1 - mean1
    p2 = pts2 - mean2
    H = p1.T @ (p2 * weights)
    U, s, Vh = np.linalg.svd(H)
    R = U @ Vh
    if np.allclose(np.linalg.det(R), -1.0):
        Vh_prime = Vh.copy()
        Vh_prime[-1, :] *= -1.0
        R = U @ Vh_prime
    t = mean1.reshape((-1, 1)) - R @ mean2.reshape((-1, 1))
    return np.concatenate([np.concatenate([R, t], axis=1), np.hstack([np.zeros((1, R.shape[0])), [[1]]])], axis=0)


def align_pair(ip: InvariantParams, sp: SolverParams, D1, D2, u0, A=None, dim: int = 3):
    """register + T_align for one pair (object_registration.py:22-29, 88-129)."""
    M, single, A = score(ip, sp.affinityeps, D1, D2, A)
    sol = find_dense_clique(sp, M, None, u0)
    sel = A[sol["nodes"]] if len(sol["nodes"]) else np.zeros((0, 2), dtype=np.int32)
    if len(sel) < dim:
        T = np.full((dim + 1, dim + 1), np.nan)
        status = 1
    else:
        T = T_align(np.asarray(D1)[sel[:, 0], :dim], np.asarray(D2)[sel[:, 1], :dim])
        status = 0
    return dict(M=M, single=single, A=A, sel=sel, T=T, status=status, **sol)


# ---------------------------------------------------------------- pair-loop prefilters
def submap_similarity(desc1: np.ndarray, desc2: np.ndarray) -> float:
    """Submap.similarity (roman/map/map.py:145-162): max over frame pairs of the cosine of their
    descriptors; a pair whose norm product is within 1e-9 of zero scores 0.  1-d inputs (mean
    descriptors, :149-152) are the one-frame case of the same formula."""
    d1 = np.atleast_2d(np.asarray(desc1, dtype=np.float64))
    d2 = np.atleast_2d(np.asarray(desc2, dtype=np.float64))
    dots = np.einsum("ad,bd->ab", d1, d2)
    norm_prod = np.sqrt(np.einsum("ad,ad->a", d1, d1))[:, None] * np.sqrt(np.einsum("bd,bd->b", d2, d2))[None, :]
    with np.errstate(divide="ignore", invalid="ignore"):
        sims = dots / norm_prod
    sims[np.abs(norm_prod) <= 1e-9] = 0.0
    return float(np.max(sims))


def aabb_intersects(p1: np.ndarray, p2: np.ndarray) -> bool:
    """roman/utils.py:160-169: do the axis-aligned bounding boxes of two point sets overlap (closed intervals)."""
    lo1, hi1 = np.min(p1, axis=0), np.max(p1, axis=0)
    lo2, hi2 = np.min(p2, axis=0), np.max(p2, axis=0)
    return bool(np.all(lo1[:3] <= hi2[:3]) and np.all(hi1[:3] >= lo2[:3]))


# ---------------------------------------------------------------- Rounding::DSD (exact densest subgraph)
def densest_subgraph(M: np.ndarray, S=None) -> list[int]:
    """clipper::dsd::solve(A, S) (clipper/src/dsd.cpp:265-314, densest_subgraph :157-262) by EXHAUSTIVE search —
    usable for |S| <= ~16.  Goldberg's network for density g has, for the source side {s} + T, the cut value
    const + 2 (g |T| - w(T)); so "the minimum cut is non-trivial" <=> max_T (w(T) - g |T|) > 0, and the set the
    reference keeps (vertices reachable from s in the residual graph = the minimal source side) is the
    intersection of all maximisers.  The binary search uses the reference's constants: m = |S|^2 - |S|,
    interval [0, m // 2], stop when n (n - 1) (U - L) < 1 with n = all nodes of the problem.
    M: (n, n) weights, upper triangle authoritative, diagonal ignored."""
    M = np.asarray(M, dtype=np.float64)
    n = M.shape[0]
    S = list(range(n)) if S is None or len(S) == 0 else [int(v) for v in S]
    k = len(S)
    if k == 1:
        return S
    Wu = np.triu(M, 1)
    W = (Wu + Wu.T)[np.ix_(S, S)]
    sizes = np.zeros(1 << k)
    wT = np.zeros(1 << k)
    for mask in range(1, 1 << k):
        low = (mask & -mask).bit_length() - 1
        rest = mask & (mask - 1)
        sizes[mask] = sizes[rest] + 1
        wT[mask] = wT[rest] + sum(W[low, j] for j in range(k) if rest >> j & 1)
    L, U = 0.0, float((k * k - k) // 2)
    best = None
    while n * (n - 1) * (U - L) >= 1:
        g = (U + L) / 2
        val = wT[1:] - g * sizes[1:]
        vmax = val.max()
        if vmax > 1e-12:
            L = g
            best = int(np.bitwise_and.reduce(np.nonzero(val >= vmax - 1e-12)[0] + 1))
        else:
            U = g
    if best is None:
        return []
    return sorted(S[j] for j in range(k) if best >> j & 1)
