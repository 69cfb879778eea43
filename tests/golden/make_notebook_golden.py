#!/usr/bin/env python
"""
Fixtures from the reference's example notebook (clipper/examples/python/ex4_bunny.ipynb), executed here:
    python tests/golden/make_notebook_golden.py   ->  tests/golden/notebook_vectors.json

The notebook holds three pieces of the reference authors' own Python that touch this path:
  cell 2   generate_dataset(...)           the noisy-bunny-with-outliers association problem (examples/data/bun10k.ply)
  cell 13  class Custom(PairwiseInvariant) the Euclidean-distance invariant written in Python (c < eps ? exp(-c^2 / 2 sigma^2) : 0)
  cell 17  k2ij + score_pairwise_consistency  a pure-Python statement of CLIPPER::scorePairwiseConsistency
           (clipper/src/clipper.cpp:29-73) incl. the "same point on either side -> no edge" rule
They are exec'd verbatim from the .ipynb (open3d's PLY reader is replaced by a 10-line binary reader; `clipperpy` is
only needed for the base-class name).  Stored:
  "affinity": a 150-association problem with the notebook's M (non-zero strict-upper entries) — a second, independent
              pin of the Euclidean affinity build next to the 12x12 literal of affinity_test.cpp;
  "k2ij":     the notebook's k2ij for a few n;
  "bunny":    the notebook's headline problem (m = 1000, 90 % outlier associations, sigma 0.01) as arrays; its printed
              result there is "95 inliers ... precision 1.00, recall 0.95" for another random draw.
"""
import base64
import json
import os
import struct
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
NB = "/root/reference/dependencies/clipper/examples/python/ex4_bunny.ipynb"
PLY = "/root/reference/dependencies/clipper/examples/data/bun10k.ply"


def read_ply_xyz(path):
    raw = open(path, "rb").read()
    head, body = raw.split(b"end_header\n", 1)
    n = int([ln for ln in head.decode().splitlines() if ln.startswith("element vertex")][0].split()[-1])
    assert b"binary_little_endian" in head
    return np.array(struct.unpack("<" + "f" * 3 * n, body[:12 * n]), dtype=np.float64).reshape(n, 3)


def b64(a):
    a = np.ascontiguousarray(a)
    return dict(dtype=str(a.dtype), shape=list(a.shape), data=base64.b64encode(a.tobytes()).decode())


def main():
    cells = ["".join(c["source"]) for c in json.load(open(NB))["cells"]]
    o3d = types.SimpleNamespace(io=types.SimpleNamespace(read_point_cloud=lambda f: types.SimpleNamespace(points=read_ply_xyz(f))))
    clipperpy = types.SimpleNamespace(invariants=types.SimpleNamespace(PairwiseInvariant=type("PairwiseInvariant", (), {"__init__": lambda self: None})),
                                      utils=types.SimpleNamespace())
    ns = dict(np=np, o3d=o3d, clipperpy=clipperpy)
    exec(cells[2], ns)    # generate_dataset
    exec(cells[13], ns)   # Custom invariant
    exec(cells[17], ns)   # k2ij, score_pairwise_consistency
    from scipy.spatial.transform import Rotation

    def problem(seed, m, n1, n2o, outrat, sigma):
        np.random.seed(seed)
        T = np.eye(4)
        T[:3, :3] = Rotation.random(random_state=seed).as_matrix()
        T[:3, 3] = np.random.uniform(low=-5, high=5, size=(3,))
        D1, D2, Agt, A = ns["generate_dataset"](PLY, m, n1, n2o, outrat, sigma, T)
        return D1, D2, Agt.astype(np.int32), A, T

    # ---- affinity pin: the notebook's pure-Python scoring
    D1, D2, Agt, A, T = problem(1, 150, 120, 30, 0.8, 0.01)
    inv = ns["Custom"](0.015, 0.02)
    M, C = ns["score_pairwise_consistency"](inv, D1, D2, A.astype("int"))
    assert np.array_equal(C - np.eye(len(A)) > 0, M - np.eye(len(A)) > 0)   # C is the pattern of M
    iu = np.triu_indices(len(A), 1)
    nz = np.nonzero(M[iu])[0]
    affinity = dict(sigma=0.015, epsilon=0.02, D1=b64(D1.T), D2=b64(D2.T), A=b64(A),
                    upper_nonzero_ij=b64(np.stack([iu[0][nz], iu[1][nz]], axis=1).astype(np.int32)),
                    upper_nonzero_v=b64(M[iu][nz].astype(np.float64)))
    k2 = {str(n): [list(ns["k2ij"](k, n)) for k in range(n * (n - 1) // 2)] for n in (2, 3, 7, 12)}
    # ---- the notebook's headline problem
    D1, D2, Agt, A, T = problem(2, 1000, 1000, 250, 0.9, 0.01)
    bunny = dict(sigma=0.01, epsilon=0.02, D1=b64(D1.T.astype(np.float64)), D2=b64(D2.T), A=b64(A), Agt=b64(Agt), T_21=T.tolist(),
                 notebook_output="CLIPPER selected 95 inliers from 1000 putative associations (precision 1.00, recall 0.95)")
    doc = dict(source="mit-acl/clipper@0666d38 examples/python/ex4_bunny.ipynb cells 2, 13, 17 executed verbatim; "
                      "examples/data/bun10k.ply", affinity=affinity, k2ij=k2, bunny=bunny)
    path = os.path.join(HERE, "notebook_vectors.json")
    json.dump(doc, open(path, "w"))
    print("wrote", path, os.path.getsize(path), "bytes; affinity nnz", len(nz), "of", len(iu[0]))


if __name__ == "__main__":
    main()
