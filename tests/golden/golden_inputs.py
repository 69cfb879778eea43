"""Seeded input constructors shared by make_reference_py_golden.py and the tests (so that large
inputs need not be stored in the JSON fixtures)."""
import numpy as np

SIMILARITY_SHAPES = [(1, 1, 16), (3, 5, 32), (7, 2, 768), (12, 9, 768), (40, 33, 768)]


def similarity_inputs(case: int):
    """Stacked frame descriptors (N1, D), (N2, D) of two submaps; cases with N1 > 2 contain a zero-norm
    row (similarity forced to 0 for it, reference roman/map/map.py:160-161) and one strongly similar pair."""
    a, b, D = SIMILARITY_SHAPES[case]
    rng = np.random.default_rng(500 + case)
    d1, d2 = rng.normal(size=(a, D)), rng.normal(size=(b, D))
    if a > 2:
        d1[1] = 0.0
        d2[0] = 0.9 * d1[0] + 0.1 * d2[0]
    return d1, d2
