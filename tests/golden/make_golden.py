#!/usr/bin/env python
"""
Regenerates the committed fixtures in tests/golden/.  Run from the repo root, in the container
that has /root/reference mounted:   python tests/golden/make_golden.py

 1. reference_kats.json — the known-answer data held by the reference's own tests, transcribed
    from the test sources (numbers only):
      clipper/test/affinity_test.cpp:33-48,94-107   4-vs-3 point scene + the 12x12 affinity literal
      clipper/test/clipper_test.cpp:62-66           same scene: 3 inliers, each (i, i)
      clipper/test/dsd_test.cpp:15-36               20x20 weighted affinity literal, dense cluster {3,5,12,14,15}
 2. oracle_vectors.json — seeded synthetic cases (roman_b200.synth) with the outputs of the CPU
    oracle (oracle/clipper_oracle.c): inlier associations, T, score, nnz.  The reference itself cannot
    be built or imported here (no Eigen3 / clipperpy), so these vectors pin the CUDA path to the
    oracle, and the oracle is pinned to the reference by (1).
"""
import json
import os
import re
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/dependencies/clipper/test"


def literal_after(src: str, marker: str):
    i = src.index(marker)
    j = src.index(";", i)
    body = src[i + len(marker):j]
    return [float(x) for x in re.findall(r"[-+]?\d*\.\d+|[-+]?\d+\.?", body)]


def reference_kats():
    aff = open(os.path.join(REF, "affinity_test.cpp")).read()
    dsd = open(os.path.join(REF, "dsd_test.cpp")).read()
    Mtrue = np.array(literal_after(aff, "Mtrue <<")).reshape(12, 12)
    M20 = np.array(literal_after(dsd, "M <<")).reshape(20, 20)
    nodes = [int(x) for x in re.search(r"dsd_nodes = \{([^}]*)\}", dsd).group(1).split(",")]
    return {
        "source": "mit-acl/clipper@0666d38 test/affinity_test.cpp, test/clipper_test.cpp, test/dsd_test.cpp",
        "scene": {"model": [[0, 0, 0], [2, 0, 0], [0, 3, 0], [2, 2, 0]], "yaw": "pi/8", "translation": [5, 3, 0],
                  "data_is": "T^-1 * model, first 3 columns", "euclidean_params": {"sigma": 0.01, "epsilon": 0.06, "mindist": 0.0}},
        "affinity_12x12": Mtrue.astype(int).tolist(),
        "clipper_test": {"n_inliers": 3, "rule": "A(i,0) == A(i,1) for every inlier"},
        "dsd_20x20": M20.tolist(),
        "dsd_nodes": nodes,
    }


CASES = [  # (method, n1, n2, semantics_dim, desc_common, seed)
    ("clipper", 30, 30, 0, 0.0, 0),
    ("gravity", 16, 21, 0, 0.0, 1),
    ("pcavolgrav", 25, 25, 0, 0.0, 2),
    ("roman", 20, 24, 64, 1.0, 3),
    ("roman", 40, 40, 512, 0.0, 4),
    ("roman", 40, 50, 768, 0.0, 5),
    ("sevg", 12, 31, 32, 0.9, 6),
    ("semanticgrav", 18, 18, 48, 1.0, 7),
]


def oracle_vectors():
    import oracle
    from oracle import clipper_np as cn
    from roman_b200 import synth
    from roman_b200.engine import method_invariant_kwargs
    O = oracle.COracle()
    out = []
    for method, n1, n2, D, dc, seed in CASES:
        kw = method_invariant_kwargs(method, semantics_dim=D if D else 768)
        ip = cn.InvariantParams(**kw)
        p = synth.make_pair(seed, n1, n2, method=method, semantics_dim=D if D else 768, desc_common=dc)
        r = O.align_pair(ip, cn.SolverParams(), p.D1, p.D2, p.u0)
        Tn = cn.T_align(p.D1[r["sel"][:, 0], :3], p.D2[r["sel"][:, 1], :3]) if r["status"] == 0 else None
        out.append({"method": method, "n1": n1, "n2": n2, "semantics_dim": D, "desc_common": dc, "seed": seed,
                    "sel": r["sel"].tolist(), "status": r["status"], "score": r["score"], "nnz": r["nnz"],
                    "T": None if Tn is None else Tn.tolist()})
    return {"generator": "tests/golden/make_golden.py (oracle/clipper_oracle.c + numpy SVD for T)", "cases": out}


if __name__ == "__main__":
    if os.path.isdir(REF):
        with open(os.path.join(HERE, "reference_kats.json"), "w") as f:
            json.dump(reference_kats(), f, indent=1)
    else:
        print("reference tree not mounted: reference_kats.json left untouched")
    with open(os.path.join(HERE, "oracle_vectors.json"), "w") as f:
        json.dump(oracle_vectors(), f, indent=1)
    print("golden fixtures written")
