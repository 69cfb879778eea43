"""Minimal object carrier with the attributes the registration classes read from the reference's
roman.object.segment.Segment / SegmentMinimalData (centre, volume, linearity, planarity,
scattering, extent, semantic_descriptor; reference roman/object/segment.py:434-472)."""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np


@dataclass
class SimpleObject:
    center: np.ndarray
    volume: float = 1.0
    linearity: float = 0.0
    planarity: float = 0.0
    scattering: float = 0.0
    extent: np.ndarray = field(default_factory=lambda: np.ones(3))
    semantic_descriptor: np.ndarray | None = None
    id: int = 0

    def normalized_eigenvalues(self):
        return None


def objects_from_rows(D: np.ndarray, method: str, semantics_dim: int = 0):
    """Inverse of ROMANRegistration._object_to_clipper_list for rows made by roman_b200.synth."""
    from ..synth import METHOD_FLAGS
    gravity, volume, pca, extent, semantics = METHOD_FLAGS[method]
    out = []
    for k, row in enumerate(np.asarray(D)):
        o = 3
        ob = SimpleObject(center=row[:3].copy(), id=k)
        if pca:
            ob.linearity, ob.planarity, ob.scattering = row[o:o + 3]
            o += 3
        if volume:
            ob.volume = row[o]
            o += 1
        if extent:
            ob.extent = row[o:o + 3].copy()
            o += 3
        if semantics:
            ob.semantic_descriptor = row[o:o + semantics_dim].copy()
        out.append(ob)
    return out
