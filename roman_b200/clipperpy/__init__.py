"""
A `clipperpy`-compatible module backed by libroman_b200.so (CUDA, sm_100a).

It exports the names roman.align touches in the reference's pybind11 module
(clipper/bindings/python/py_clipper.cpp:177-332): invariants.{ROMANParams, ROMAN,
EuclideanDistanceParams, EuclideanDistance, PairwiseInvariant}, Params, Rounding, Solution,
CLIPPER, CLIPPERPairwiseAndSingle, utils.{create_all_to_all, k2ij}, __version__.

To drop it in under an unmodified roman.align:   import roman_b200; roman_b200.install_as_clipperpy()
"""
from . import invariants, utils
from ._core import CLIPPER, CLIPPERPairwiseAndSingle, Params, Rounding, Solution, NONZERO, DSD, DSD_HEU

dsd = utils  # the reference wires the `dsd` submodule to pybind_utils (py_clipper.cpp:188-189)
__version__ = "0.2.4"  # interface version of the reference module this mirrors (clipper/CMakeLists.txt:2)

__all__ = ["invariants", "utils", "dsd", "CLIPPER", "CLIPPERPairwiseAndSingle", "Params", "Rounding", "Solution",
           "NONZERO", "DSD", "DSD_HEU", "__version__"]
