"""
roman_b200 — B200-native (sm_100a) CLIPPER submap registration behind the roman.align / clipperpy API.

Only the hot path of mit-acl/roman's `roman.align` is here: the pairwise affinity
build, the CLIPPER dense-clique solver and the Arun/SVD registration, as hand-written CUDA behind
the C ABI of include/roman_b200.h, plus the Python host layer that mirrors the reference's
`clipperpy` module and `roman.align` registration classes.
"""
__version__ = "0.1.0"
