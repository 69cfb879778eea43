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


def _yaw_pose(yaw, xyz):
    T = np.eye(4)
    c, s = np.cos(yaw), np.sin(yaw)
    T[:2, :2] = [[c, -s], [s, c]]
    T[:3, 3] = xyz
    return T


PAIR_LOOP_SCENARIOS = ("roman_prefilters", "upside_down", "single_robot")


def pair_loop_scenario(name: str) -> dict:
    """Inputs of one run of the reference's pair loop (roman/align/submap_align.py:93-200): for each of the two
    robots a list of submaps, each a dict(rows, ids, pose_flu, time, descriptor); plus the SubmapAlignParams
    keywords and skip_distance.  rows are roman_b200.synth feature rows of `method`."""
    from roman_b200 import synth
    rng = np.random.default_rng({"roman_prefilters": 11, "upside_down": 12, "single_robot": 13}[name])
    if name == "roman_prefilters":
        sem, method = 64, "roman"
        ps = [synth.make_pair(700 + k, 14 + k, 16, method=method, semantics_dim=sem) for k in range(4)]
        base = [rng.normal(size=(4, 48)) for _ in range(3)]
        d1 = [base[0][:2] + 0.01 * rng.normal(size=(2, 48)), rng.normal(size=(3, 48)),
              base[2][:2] + 0.01 * rng.normal(size=(2, 48)), base[0][1:3] + 0.01 * rng.normal(size=(2, 48))]
        pos0 = [[0.0, 0, 0], [12.0, 3, 0], [25.0, -4, 0.5]]
        pos1 = [[2.0, 1, 0], [14.0, 2, 0], [400.0, 0, 0], [5.0, -3, 0.2]]
        runs = [[dict(rows=ps[k].D1, ids=list(range(len(ps[k].D1))), pose_flu=_yaw_pose(0.3 * k, pos0[k]), time=10.0 * k,
                      descriptor=base[k]) for k in range(3)],
                [dict(rows=ps[k].D2, ids=list(range(len(ps[k].D2))), pose_flu=_yaw_pose(-0.2 * k, pos1[k]), time=100.0 + 10 * k,
                      descriptor=d1[k]) for k in range(4)]]
        return dict(method=method, sem=sem, skip_distance=60.0, runs=runs,
                    sm_kwargs=dict(method=method, semantics_dim=sem, submap_descriptor="stacked_frame_descriptors",
                                   submap_descriptor_thresh=0.8))
    if name == "upside_down":
        method = "clipper"
        ps = [synth.make_pair(710 + k, 15, 17, method=method) for k in range(2)]
        flip = np.diag([1.0, -1.0, -1.0])                      # 180 degrees about x
        rows1 = [ps[0].D2 @ flip, ps[1].D2]
        runs = [[dict(rows=ps[k].D1, ids=list(range(15)), pose_flu=_yaw_pose(0.1, [3.0 * k, 0, 0]), time=5.0 * k,
                      descriptor=None) for k in range(2)],
                [dict(rows=rows1[k], ids=list(range(17)), pose_flu=_yaw_pose(0.4, [3.0 * k, 1, 0]), time=50.0 + k,
                      descriptor=None) for k in range(2)]]
        return dict(method=method, sem=0, skip_distance=np.inf, runs=runs, sm_kwargs=dict(method=method))
    if name == "single_robot":
        method = "gravity"
        ps = [synth.make_pair(720 + k, 16, 18, method=method) for k in range(2)]
        ids1 = [list(range(10, 28)), list(range(100, 118))]   # submap 0 of run 1 shares ids 10..15 with run 0's submaps
        runs = [[dict(rows=ps[k].D1, ids=list(range(16)), pose_flu=_yaw_pose(0.0, [2.0 * k, 0, 0]), time=5.0 * k,
                      descriptor=None) for k in range(2)],
                [dict(rows=ps[k].D2, ids=ids1[k], pose_flu=_yaw_pose(0.2, [2.0 * k, 1, 0]), time=300.0 + k,
                      descriptor=None) for k in range(2)]]
        return dict(method=method, sem=0, skip_distance=np.inf, runs=runs,
                    sm_kwargs=dict(method=method, single_robot_lc=True))
    raise KeyError(name)


MAP_CUT_CASES = {
    "default": dict(),
    "time_pruning_small": dict(submap_pruning_method="time", submap_max_size=6, submap_center_time=8.0),
    "no_radius": dict(submap_radius=None, submap_max_size=12, submap_center_dist=14.0),
    "force_fill": dict(force_fill_submaps=True, submap_max_size=10, submap_overlap=4),
    "mean_semantic": dict(submap_descriptor="mean_semantic", submap_max_size=8),
    "mean_frame": dict(submap_descriptor="mean_frame_descriptor"),
    "stacked": dict(submap_descriptor="stacked_frame_descriptors", submap_radius=8.0),
    "stacked_thinned": dict(submap_descriptor="stacked_frame_descriptors", frame_descriptor_dist=3.0),
}


def map_scenario() -> dict:
    """A robot's object map for the submap-cutting fixtures: 70 trajectory poses along a bending path (with some roll and
    pitch), 55 objects scattered around it with sighting intervals, 8-d semantic and frame descriptors."""
    from scipy.spatial.transform import Rotation as Rot
    rng = np.random.default_rng(900)
    n = 70
    s = np.linspace(0.0, 1.0, n)
    pos = np.stack([60.0 * s, 12.0 * np.sin(3.0 * s), 0.3 * np.cos(5.0 * s)], axis=1)
    traj = np.tile(np.eye(4), (n, 1, 1))
    for k in range(n):
        yaw = 0.6 * np.cos(3.0 * s[k])
        traj[k, :3, :3] = Rot.from_euler("ZYX", [yaw, 0.05 * np.sin(9 * s[k]), 0.04 * np.cos(7 * s[k])]).as_matrix()
        traj[k, :3, 3] = pos[k]
    times = 100.0 + 1.7 * np.arange(n) + rng.uniform(0, 0.3, n)
    segs = []
    for j in range(55):
        k = int(rng.integers(0, n))
        c = pos[k] + rng.normal(0, 5.0, 3) * np.array([1.0, 1.0, 0.2])
        first = times[k] - rng.uniform(0, 6)
        segs.append(dict(id=j, center=c, volume=float(rng.uniform(0.1, 3)), linearity=float(rng.uniform(0, 1)),
                         planarity=float(rng.uniform(0, 1)), scattering=float(rng.uniform(0, 1)),
                         extent=rng.uniform(0.2, 2, 3), semantic_descriptor=rng.normal(size=8),
                         first_seen=float(first), last_seen=float(first + rng.uniform(1, 15))))
    return dict(trajectory=traj, times=times, segments=segs, descriptors=rng.normal(size=(n, 8)))
