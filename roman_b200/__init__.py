"""
roman_b200 — B200-native (sm_100a) CLIPPER submap registration behind the roman.align / clipperpy API.

Only the hot path of mit-acl/roman's `roman.align` is here: the pairwise affinity
build, the CLIPPER dense-clique solver and the Arun/SVD registration, as hand-written CUDA behind
the C ABI of include/roman_b200.h, plus the Python host layer that mirrors the reference's
`clipperpy` module and `roman.align` registration classes.
"""
__version__ = "0.1.0"


def set_default_tile_skip(on: bool):
    """Process-wide default packed-affinity format of new engines: False dense (default), True tile skip (fast path)."""
    from . import engine
    engine.set_default_tile_skip(on)


def install_as_clipperpy():
    """Register roman_b200.clipperpy under the module name `clipperpy`, so that an unmodified
    `roman.align` (which does `import clipperpy`) runs on the B200 path."""
    import sys
    from . import clipperpy as _cp
    sys.modules["clipperpy"] = _cp
    sys.modules["clipperpy.invariants"] = _cp.invariants
    sys.modules["clipperpy.utils"] = _cp.utils
    return _cp
