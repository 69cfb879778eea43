"""
Synthetic submap-pair generator shared by the tests, bench.py and the CPU baseline.

Follows SURVEY.md §8(d) "Synthetic generator": gravity-aligned submap pairs whose shapes match
what `submaps_from_roman_map` produces (reference roman/map/map.py:244-357, <=40 objects in a
15 m radius, roman/params/submap_align_params.py:36,42) and whose per-object feature row is the
one `ROMANRegistration._object_to_clipper_list` packs (reference roman/align/roman_registration.py:98-108):
    [cx, cy, cz | linearity, planarity, scattering | volume | sorted extent(3) | descriptor(D)]
with only the groups the method enables.  The Kimera-Multi demo data is not available offline
(SURVEY §0.15), so every BASELINE.json config runs on these stand-ins.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

# method -> (gravity, volume, pca, extent, semantics); reference submap_align_params.py:98-108
METHOD_FLAGS = {
    "clipper":       (False, False, False, False, False),
    "gravity":       (True,  False, False, False, False),
    "pcavolgrav":    (True,  True,  True,  False, False),
    "extentvolgrav": (True,  True,  False, True,  False),
    "roman":         (True,  True,  True,  False, True),
    "sevg":          (True,  True,  False, True,  True),
    "spv":           (False, True,  True,  False, False),
    "semanticgrav":  (True,  False, False, False, True),
}


@dataclass
class SynthPair:
    D1: np.ndarray        # (n1, d) float64, object-major feature rows
    D2: np.ndarray        # (n2, d)
    u0: np.ndarray        # (n1*n2,) initial vector for the solver
    T_gt: np.ndarray      # (4,4) maps frame-2 points into frame 1
    n_inliers: int        # number of planted matches (objects 0..n_in-1 in both maps)
    centres1: np.ndarray
    centres2: np.ndarray


def _unit(v):
    return v / np.linalg.norm(v, axis=-1, keepdims=True)


def make_pair(seed: int, n1: int, n2: int, method: str = "roman", semantics_dim: int = 768,
              inlier_frac: float = 0.5, pos_noise: float = 0.05, desc_common: float = 0.0) -> SynthPair:
    """
    One synthetic pair per SURVEY §8(d).  `desc_common` > 0 adds a shared component to every
    descriptor (CLIP/DINO embeddings live in a cone, so unrelated objects have cosine well
    above 0); 0 reproduces the §8(d) generator verbatim (unmatched cosine ~ N(0, 1/sqrt(D))).
    """
    gravity, volume, pca, extent, semantics = METHOD_FLAGS[method]
    rng = np.random.default_rng(seed)

    def box(n):
        return np.stack([rng.uniform(-15, 15, n), rng.uniform(-15, 15, n), rng.uniform(-1, 3, n)], axis=1)

    c1 = box(n1)
    yaw = rng.uniform(0, 2 * np.pi)
    t = rng.uniform(-5, 5, 3)
    R = np.array([[np.cos(yaw), -np.sin(yaw), 0], [np.sin(yaw), np.cos(yaw), 0], [0, 0, 1.0]])
    n_in = int(np.floor(inlier_frac * min(n1, n2)))
    c2 = box(n2)
    c2[:n_in] = (c1[:n_in] - t) @ R + rng.normal(0, pos_noise, (n_in, 3))  # R^T (p1 - t) + noise

    def shape_feats(n):
        vol = np.exp(rng.uniform(np.log(0.05), np.log(5.0), n))
        e = np.sort(rng.uniform(0.02, 1.0, (n, 3)), axis=1)[:, ::-1]
        e = e / e.sum(axis=1, keepdims=True)
        lin = (e[:, 0] - e[:, 1]) / e[:, 0]
        plan = (e[:, 1] - e[:, 2]) / e[:, 0]
        scat = e[:, 2] / e[:, 0]
        ext = np.sort(np.cbrt(vol)[:, None] * rng.uniform(0.5, 2.0, (n, 3)), axis=1)
        return vol, np.stack([lin, plan, scat], axis=1), ext

    vol1, pca1, ext1 = shape_feats(n1)
    vol2, pca2, ext2 = shape_feats(n2)
    jitter = lambda shape: 1.0 + rng.normal(0, 0.05, shape)
    vol2[:n_in] = vol1[:n_in] * jitter(n_in)
    pca2[:n_in] = pca1[:n_in] * jitter((n_in, 3))
    ext2[:n_in] = np.sort(ext1[:n_in] * jitter((n_in, 3)), axis=1)

    cols1, cols2 = [c1], [c2]
    if pca:
        cols1.append(pca1); cols2.append(pca2)
    if volume:
        cols1.append(vol1[:, None]); cols2.append(vol2[:, None])
    if extent:
        cols1.append(ext1); cols2.append(ext2)
    if semantics:
        Dd = semantics_dim
        common = _unit(rng.normal(size=Dd)) * desc_common
        v1 = _unit(rng.normal(size=(n1, Dd)))
        v2 = _unit(rng.normal(size=(n2, Dd)))
        v2[:n_in] = _unit(v1[:n_in] + 0.3 * _unit(rng.normal(size=(n_in, Dd))))
        cols1.append(_unit(v1 + common)); cols2.append(_unit(v2 + common))
    D1 = np.ascontiguousarray(np.concatenate(cols1, axis=1))
    D2 = np.ascontiguousarray(np.concatenate(cols2, axis=1))
    u0 = rng.uniform(0, 1, n1 * n2)
    T = np.eye(4)
    T[:3, :3] = R
    T[:3, 3] = t
    return SynthPair(D1=D1, D2=D2, u0=u0, T_gt=T, n_inliers=n_in, centres1=c1, centres2=c2)


def make_batch(seed0: int, B: int, n1: int, n2: int, **kw):
    """B pairs with seeds seed0..seed0+B-1, stacked: D1 (B,n1,d), D2 (B,n2,d), u0 (B,m)."""
    pairs = [make_pair(seed0 + b, n1, n2, **kw) for b in range(B)]
    return (np.stack([p.D1 for p in pairs]), np.stack([p.D2 for p in pairs]),
            np.stack([p.u0 for p in pairs]), pairs)


@dataclass
class AllPairsWorld:
    """Two robots' submaps over the same street of objects (BASELINE cfg5): run r has S_r submaps, each a
    (n, d) feature-row table in that robot's own gravity-aligned odometry frame."""
    tables: list            # tables[r][s] -> (n_s, d) float64 rows
    centres: list           # centres[r] -> (S_r, 3) submap centre positions (robot r's frame)
    T_frames: list          # T_frames[r] -> (4, 4) world-from-odometry of robot r (yaw + translation only)
    street_x: list          # street_x[r] -> (S_r,) position of each submap centre along the street [m]


def make_all_pairs_world(seed: int, S0: int = 20, S1: int = 20, max_objects: int = 40, method: str = "roman",
                         semantics_dim: int = 768, desc_common: float = 1.0, spacing: float = 10.0,
                         radius: float = 15.0, pos_noise: float = 0.05, min_objects: int = 12) -> AllPairsWorld:
    """
    The all-pairs loop-closure search of roman/align/submap_align.py:93-200 on stand-in data: both robots drive down
    one street of objects (~200 m for 20 submaps at the reference's `submap_center_dist` 10 m, `submap_radius` 15 m,
    `submap_max_size` 40: roman/params/submap_align_params.py:36-43), each sees every object with probability 0.85,
    with 5 cm position noise, in its own yaw-rotated and shifted odometry frame.  Submap s holds the (at most
    `max_objects`) objects nearest to its centre within `radius` — so object counts are ragged (sparser stretches of
    street give smaller submaps) and submaps i, j of the two runs share objects iff they are within ~2 radii of each
    other along the street: a banded loop-closure matrix, everything else has no consistent clique.
    `desc_common` as in make_pair (1.0: descriptors live in a cone, like CLIP / DINO embeddings).
    """
    gravity, volume, pca, extent, semantics = METHOD_FLAGS[method]
    rng = np.random.default_rng(seed)
    S = (S0, S1)
    length = spacing * (max(S) + 1)
    # object density varies along the street (between ~0.6 and ~3 objects per metre): most submaps hit the size
    # cap, the ones in sparse stretches stay smaller
    n_world = int(3.2 * (length + 2 * radius))
    xs = np.sort(rng.uniform(-radius, length + radius, n_world))
    dens = 0.6 + 0.4 * np.sin(xs / 23.0 + rng.uniform(0, 6.28))
    xs = xs[rng.uniform(0, 1, n_world) < dens / dens.max()]
    n_world = len(xs)
    centres = np.stack([xs, rng.uniform(-9, 9, n_world), rng.uniform(-0.5, 2.5, n_world)], axis=1)
    vol = np.exp(rng.uniform(np.log(0.05), np.log(5.0), n_world))
    e = np.sort(rng.uniform(0.02, 1.0, (n_world, 3)), axis=1)[:, ::-1]
    pcaf = np.stack([(e[:, 0] - e[:, 1]) / e[:, 0], (e[:, 1] - e[:, 2]) / e[:, 0], e[:, 2] / e[:, 0]], axis=1)
    ext = np.sort(np.cbrt(vol)[:, None] * rng.uniform(0.5, 2.0, (n_world, 3)), axis=1)
    if semantics:
        common = _unit(rng.normal(size=semantics_dim)) * desc_common
        desc = _unit(rng.normal(size=(n_world, semantics_dim)))
    tables, cen, frames, sx = [], [], [], []
    for r in range(2):
        yaw = rng.uniform(0, 2 * np.pi)
        T = np.eye(4)
        T[:2, :2] = [[np.cos(yaw), -np.sin(yaw)], [np.sin(yaw), np.cos(yaw)]]
        T[:3, 3] = [rng.uniform(-20, 20), rng.uniform(-20, 20), rng.uniform(-0.5, 0.5)]
        Ti = np.linalg.inv(T)
        seen = rng.uniform(0, 1, n_world) < 0.85
        pos = centres @ Ti[:3, :3].T + Ti[:3, 3] + rng.normal(0, pos_noise, (n_world, 3))
        jit = lambda shape: 1.0 + rng.normal(0, 0.05, shape)
        cols = [pos]
        if pca:
            cols.append(pcaf * jit((n_world, 3)))
        if volume:
            cols.append((vol * jit(n_world))[:, None])
        if extent:
            cols.append(np.sort(ext * jit((n_world, 3)), axis=1))
        if semantics:
            cols.append(_unit(_unit(desc + 0.3 * _unit(rng.normal(size=(n_world, semantics_dim)))) + common))
        rows = np.concatenate(cols, axis=1)
        offs = rng.uniform(0.0, spacing) if r == 1 else 0.5 * spacing    # the two robots' submap centres interleave
        run_t, run_c, run_x = [], [], []
        for s in range(S[r]):
            cx = offs + spacing * s
            dist = np.hypot(centres[:, 0] - cx, centres[:, 1])
            idx = np.where(seen & (dist < radius))[0]
            idx = idx[np.argsort(dist[idx], kind="stable")][:max_objects]
            if len(idx) < min_objects:      # never hand out near-empty submaps: widen to the nearest objects
                idx = np.where(seen)[0]
                idx = idx[np.argsort(dist[idx], kind="stable")][:min_objects]
            idx = np.sort(idx)
            run_t.append(np.ascontiguousarray(rows[idx]))
            run_c.append(Ti[:3, :3] @ np.array([cx, 0.0, 0.0]) + Ti[:3, 3])
            run_x.append(cx)
        tables.append(run_t)
        cen.append(np.array(run_c))
        frames.append(T)
        sx.append(np.array(run_x))
    return AllPairsWorld(tables=tables, centres=cen, T_frames=frames, street_x=sx)
