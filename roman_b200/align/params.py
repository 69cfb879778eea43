"""
SubmapAlignParams with the field names and defaults of the reference dataclass
(reference roman/params/submap_align_params.py:24-84) and its registration factory
`get_object_registration` (:86-150), here driven by the method table in roman_b200.synth.
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field, fields
from typing import Any, List, Optional

import numpy as np

import yaml

from ..synth import METHOD_FLAGS
from .dist_reg_with_pruning import DistRegWithPruning
from .roman_registration import FusionMethod, ROMANParams, ROMANRegistration

_METHOD_ALIASES = {"spvg": "roman", "roman_no_semantics": "pcavolgrav"}
_FUSION = {"geometric_mean": FusionMethod.GEOMETRIC_MEAN, "arithmetic_mean": FusionMethod.ARITHMETIC_MEAN,
           "product": FusionMethod.PRODUCT}


@dataclass
class SubmapAlignParams:
    # what is registered and how
    dim: int = 3                               # 2-D or 3-D object maps
    method: str = "roman"                      # see METHOD_FLAGS, plus 'clipper+prune'
    fusion_method: str = "geometric_mean"      # parsed, but never reaches the invariant
    # submap construction (consumed by the caller that cuts submaps, not by the registration)
    force_fill_submaps: bool = False
    submap_max_size: int = 40
    submap_overlap: int = 20
    submap_radius: Optional[float] = 15.0
    submap_center_dist: float = 10.0
    submap_center_time: float = 50.0
    submap_pruning_method: str = "distance"
    submap_descriptor: Optional[str] = None
    frame_descriptor_dist: Optional[float] = None
    submap_descriptor_thresh: float = 0.8
    # loop-closure post-filters
    single_robot_lc: bool = False
    single_robot_lc_time_thresh: float = 50.0
    force_rm_lc_roll_pitch: bool = True
    force_rm_upside_down: bool = True
    use_object_bottom_middle: bool = False
    # invariant parameters
    sigma: float = 0.4
    epsilon: float = 0.6
    mindist: float = 0.2
    epsilon_shape: float = 0.0
    ransac_iter: int = 1_000_000
    cosine_min: float = 0.5
    cosine_max: float = 0.7
    semantics_dim: int = 768
    gravity_unc_ang_rad: float = 0.0872665

    def __post_init__(self):
        if isinstance(self.submap_descriptor, str) and self.submap_descriptor.lower() == "none":
            self.submap_descriptor = None

    @classmethod
    def from_yaml(cls, yaml_file):
        with open(yaml_file, "r") as f:
            raw = yaml.full_load(f)
        known = {f.name for f in fields(cls)}
        unknown = set(raw) - known
        if unknown:
            raise TypeError(f"unknown SubmapAlignParams field(s): {sorted(unknown)}")
        return cls(**raw)

    def get_object_registration(self):
        self.method = _METHOD_ALIASES.get(self.method, self.method)
        if self.method == "clipper+prune":
            return DistRegWithPruning(sigma=self.sigma, epsilon=self.epsilon, mindist=self.mindist,
                                      shape_epsilon=self.epsilon_shape, cos_min=self.cosine_min, dim=self.dim,
                                      use_gravity=True)
        if self.method == "ransac":
            raise NotImplementedError("'ransac' (open3d RANSAC baseline) is outside the CLIPPER hot path and is "
                                      "unusable in the reference as well (SURVEY §0.10)")
        if self.method not in METHOD_FLAGS:
            raise AssertionError("Invalid method")
        gravity, volume, pca, extent, semantics = METHOD_FLAGS[self.method]
        rp = ROMANParams(point_dim=self.dim, sigma=self.sigma, epsilon=self.epsilon, mindist=self.mindist,
                         gravity=gravity, volume=volume, pca=pca, extent=extent,
                         semantics_dim=self.semantics_dim if semantics else 0,
                         gravity_unc_ang_rad=self.gravity_unc_ang_rad, cos_min=self.cosine_min,
                         cos_max=self.cosine_max, epsilon_shape=self.epsilon_shape)
        rp.fusion_method = _FUSION[self.fusion_method]
        return ROMANRegistration(rp)


@dataclass
class SubmapAlignInputOutput:
    """Input / output specification of `submap_align` — the field names, defaults and derived file names of the
    reference dataclass (reference roman/params/submap_align_params.py:153-198)."""
    inputs: List[Any]
    output_dir: str
    run_name: str
    input_type_pkl: bool = True
    input_type_json: bool = False
    input_gt_pose_yaml: List[Optional[str]] = field(default_factory=lambda: [None, None])
    robot_names: List[str] = field(default_factory=lambda: ["0", "1"])
    robot_env: Optional[str] = None
    lc_association_thresh: int = 4
    g2o_t_std: float = 0.5
    g2o_r_std: float = np.deg2rad(0.5)
    debug_show_maps: bool = False
    skip_distance: float = np.inf

    def _out(self, suffix: str) -> str:
        return os.path.join(self.output_dir, f"{self.run_name}{suffix}")

    @property
    def output_img(self):
        return self._out(".png")

    @property
    def output_matrix(self):
        return self._out(".matrix.pkl")

    @property
    def output_pkl(self):
        return self._out(".pkl")

    @property
    def output_timing(self):
        return self._out(".timing.txt")

    @property
    def output_params(self):
        return self._out(".params.txt")

    @property
    def output_g2o(self):
        return self._out(".g2o")

    @property
    def output_lc_json(self):
        return self._out(".json")

    @property
    def output_submaps(self):
        return [os.path.join(self.output_dir, f"{rn}.sm.json") for rn in self.robot_names]
