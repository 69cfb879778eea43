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


def install_as_roman():
    """Register this package's mirrors under the reference's module paths (`roman.align.submap_align`,
    `roman.params.submap_align_params`, `roman.align.results`, `roman.align.object_registration`, ...), together with
    `clipperpy`, so that a caller written against the reference — demo/loop_closures.py:3-4,19-28 — runs unchanged:

        from roman.params.submap_align_params import SubmapAlignInputOutput, SubmapAlignParams
        from roman.align.submap_align import submap_align
    """
    import importlib
    import sys
    import types
    from . import align
    mods = {n: importlib.import_module(f"{__name__}.align.{n}") for n in
            ("dist_reg_with_pruning", "object_registration", "params", "results", "roman_registration", "submap_align")}
    install_as_clipperpy()
    for name in ("roman", "roman.align", "roman.params"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__path__ = []
            sys.modules[name] = m
    table = {"roman.align.submap_align": mods["submap_align"], "roman.align.results": mods["results"],
             "roman.align.object_registration": mods["object_registration"],
             "roman.align.roman_registration": mods["roman_registration"],
             "roman.align.dist_reg_with_pruning": mods["dist_reg_with_pruning"],
             "roman.params.submap_align_params": mods["params"]}
    for name, mod in table.items():
        sys.modules[name] = mod
        setattr(sys.modules[name.rsplit(".", 1)[0]], name.rsplit(".", 1)[1], mod)
    sys.modules["roman"].align = sys.modules["roman.align"]
    sys.modules["roman"].params = sys.modules["roman.params"]
    return align
