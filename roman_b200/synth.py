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
