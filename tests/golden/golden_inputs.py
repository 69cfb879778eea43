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
