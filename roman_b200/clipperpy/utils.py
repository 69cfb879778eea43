"""clipperpy.utils (py_clipper.cpp:148-160)."""
import math

import numpy as np


def create_all_to_all(n1: int, n2: int) -> np.ndarray:
    """utils::createAllToAll (clipper/include/clipper/utils.h:61-71): A[i*n2 + j] = (i, j), int32."""
    i, j = np.divmod(np.arange(int(n1) * int(n2), dtype=np.int64), max(int(n2), 1))
    return np.stack([i, j], axis=1).astype(np.int32)


def k2ij(k: int, n: int):
    """utils::k2ij (clipper/src/utils.cpp:87-97): flat index of the strict upper triangle -> (i, j)."""
    k += 1
    l = n * (n - 1) // 2 - k
    o = int(math.floor((math.sqrt(1 + 8 * l) - 1) / 2.0))
    p = l - o * (o + 1) // 2
    i = n - (o + 1)
    j = n - p
    return (i - 1, j - 1)
