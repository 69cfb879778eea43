"""
Host-side mirror of the reference's roman.align registration interface for the CLIPPER path
(same class names, argument meaning, return types and exceptions), plus the batched entry point
`register_batch` that the GPU path is built for.

  object_registration.py   <- roman/align/object_registration.py
  roman_registration.py    <- roman/align/roman_registration.py
  dist_reg_with_pruning.py <- roman/align/dist_reg_with_pruning.py
  params.py                <- roman/params/submap_align_params.py (SubmapAlignParams, SubmapAlignInputOutput)
  submap_align.py          <- roman/align/submap_align.py: submap_align(sm_params, sm_io) and its pair loop, batched (all pairs in one launch)
  submaps.py               <- submaps_from_roman_map / extract_submap_descriptors of roman/map/map.py:210-357 (cutting a map into submaps)
  prefilter.py             <- the tests made before a pair is registered (submap_align.py:96-149; map.py:145-162)
  results.py               <- roman/align/results.py: SubmapAlignResults, save_submap_align_results and the writers (align.g2o, align.json, <robot>.sm.json, timing / params / pickles)
  g2o_fusion.py            <- roman/offline_rpgo/g2o_file_fusion.py (odometry + loop-closure g2o files -> one pose graph file)
"""
from .object_registration import InsufficientAssociationsException, ObjectRegistration  # noqa: F401
from .roman_registration import FusionMethod, ROMANParams, ROMANRegistration  # noqa: F401
from .dist_reg_with_pruning import DistRegWithPruning, GravityConstraintError  # noqa: F401
from .params import SubmapAlignInputOutput, SubmapAlignParams  # noqa: F401
from .objects import SimpleObject  # noqa: F401
from .submap_align import align_submap_pairs, load_roman_map, submap_align, SubmapPairResults  # noqa: F401
from .results import (SubmapAlignResults, SubmapMeta, loop_closures, save_submap_align_results, write_g2o,  # noqa: F401
                      write_lc_json, write_submaps_json)
from .g2o_fusion import create_config, g2o_file_fusion  # noqa: F401
from .submaps import MapData, SubmapData, cut_submaps  # noqa: F401
