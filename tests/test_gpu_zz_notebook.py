"""
GPU counterpart of tests/test_notebook_golden.py: the CUDA path on the fixtures made by executing the reference authors'
own Python from clipper/examples/python/ex4_bunny.ipynb — their pure-Python affinity scoring and the notebook's headline
bunny problem (1000 putative associations, 90 % outliers).
"""
import json
import os

import numpy as np
import pytest

from oracle import clipper_np as cn
from roman_b200.engine import make_invariant_params
from tests.helpers import oracle_invariant, rel_fro
from tests.test_notebook_golden import arr

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def nb():
    with open(os.path.join(HERE, "golden", "notebook_vectors.json")) as f:
        return json.load(f)


def test_affinity_kernel_equals_the_notebooks_pure_python_scoring(engine, nb):
    a = nb["affinity"]
    D1, D2, A = arr(a["D1"]), arr(a["D2"]), arr(a["A"])
    ij, v = arr(a["upper_nonzero_ij"]), arr(a["upper_nonzero_v"])
    m = len(A)
    want = np.zeros((m, m))
    want[ij[:, 0], ij[:, 1]] = v
    want = want + want.T + np.eye(m)
    try:
        engine.set_invariant_params(make_invariant_params(kind=1, sigma=a["sigma"], epsilon=a["epsilon"], mindist=0.0))
        engine.plan(3, [len(D1)], [len(D2)], m=[m])
        engine.affinity(D1.reshape(-1), D2.reshape(-1), A=A.astype(np.int32))
        Mg, Cg = engine.get_dense(0)
        Mg, Cg = Mg.cpu().numpy(), Cg.cpu().numpy()
    finally:
        engine.set_invariant_params(make_invariant_params())
    assert np.array_equal(Mg != 0, want != 0) and np.array_equal(Cg, (want != 0).astype(float))
    assert np.abs(Mg - want).max() <= 1e-12


def test_bunny_problem_on_the_gpu(engine, coracle, nb):
    b = nb["bunny"]
    D1, D2, A, Agt = arr(b["D1"]), arr(b["D2"]), arr(b["A"]), arr(b["Agt"])
    kw = dict(kind=1, sigma=b["sigma"], epsilon=b["epsilon"], mindist=0.0)
    u0 = np.random.default_rng(0).uniform(0, 1, len(A))
    try:
        engine.set_invariant_params(make_invariant_params(**kw))
        engine.plan(3, [len(D1)], [len(D2)], m=[len(A)])
        res = engine.align_batch(D1.reshape(-1), D2.reshape(-1), u0, A=A.astype(np.int32))
        sel = res.associations(0)
        T = res.T[0].cpu().numpy()
    finally:
        engine.set_invariant_params(make_invariant_params())
    ref = coracle.align_pair(oracle_invariant(**kw), cn.SolverParams(), D1, D2, u0, A=A)
    assert np.array_equal(sel, ref["sel"])                        # same ordered inliers as the CPU restatement
    assert rel_fro(T, ref["T"]) <= 1e-9
    truth = {tuple(r) for r in Agt.tolist()}
    hits = sum(tuple(r) in truth for r in sel.tolist())
    assert hits / len(sel) >= 0.97 and hits / len(Agt) >= 0.9     # notebook: precision 1.00, recall 0.95
