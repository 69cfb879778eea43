"""
CPU tests of the host side: the C-ABI library loads and exports every symbol the header declares
(no compute calls without a GPU), the clipperpy-compatible surface exists, the registration
factory mirrors the reference's method table, the generator is deterministic, and the multi-rank
driver logic (round-robin partition + one all-gather of fixed-size records) works at world_size 2
over gloo.
"""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from roman_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "roman_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(rb_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 20
    lib = _lib.load()  # resolves every symbol in _lib.SIGNATURES
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    for name in declared:
        assert hasattr(lib, name)
    assert b"sm_100a" in lib.rb_version()


def test_struct_layouts_match_defaults():
    import ctypes as C
    from roman_b200 import _lib
    lib = _lib.load()
    ip, sp = _lib.RbInvariantParams(), _lib.RbSolverParams()
    lib.rb_default_invariant_params(C.byref(ip))
    lib.rb_default_solver_params(C.byref(sp))
    # reference defaults: roman.h:27-53, clipper.h:27-60
    assert (ip.point_dim, ip.sigma, ip.epsilon, ip.cosine_min, ip.cosine_max, ip.drift_scale) == (3, 0.01, 0.06, 0.85, 0.95, 0.1)
    assert (sp.tol_u, sp.tol_F, sp.maxiniters, sp.maxoliters, sp.beta, sp.maxlsiters) == (1e-8, 1e-9, 200, 1000, 0.25, 99)
    assert (sp.eps, sp.affinityeps, sp.rescale_u0, sp.rounding) == (1e-9, 1e-4, 1, 2)


def test_engine_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from roman_b200.engine import Engine, RomanB200Error
    with pytest.raises(RomanB200Error):
        Engine()


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under roman_b200/ may reference it."""
    for dp, _, files in os.walk(os.path.join(ROOT, "roman_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), os.path.join(dp, f)
                assert "clipper_oracle" not in src and "liboracle" not in src, os.path.join(dp, f)


def test_clipperpy_surface():
    """Names roman.align touches in the reference module."""
    import roman_b200
    cp = roman_b200.install_as_clipperpy()
    import clipperpy
    assert clipperpy is cp
    for name in ["CLIPPER", "CLIPPERPairwiseAndSingle", "Params", "Rounding", "Solution", "invariants", "utils", "__version__"]:
        assert hasattr(cp, name)
    for name in ["ROMANParams", "ROMAN", "EuclideanDistanceParams", "EuclideanDistance", "PairwiseInvariant"]:
        assert hasattr(cp.invariants, name)
    assert (cp.invariants.ROMAN.GEOMETRIC_MEAN, cp.invariants.ROMAN.ARITHMETIC_MEAN, cp.invariants.ROMAN.PRODUCT) == (0, 1, 2)
    assert cp.Rounding.DSD_HEU == cp.DSD_HEU == 2
    p = cp.invariants.ROMANParams()
    for attr in ["point_dim", "ratio_feature_dim", "cos_feature_dim", "fusion_method", "sigma", "epsilon", "mindist",
                 "distance_weight", "ratio_weight", "cosine_weight", "ratio_epsilon", "cosine_min", "cosine_max",
                 "gravity_guided", "gravity_unc_ang_rad", "drift_aware", "drift_scale_sigma", "drift_scale"]:
        assert hasattr(p, attr)
    q = cp.Params()
    for attr in ["tol_u", "tol_F", "tol_Fop", "maxiniters", "maxoliters", "beta", "maxlsiters", "eps", "affinityeps",
                 "rescale_u0", "rounding"]:
        assert hasattr(q, attr)
    A = cp.utils.create_all_to_all(4, 3)
    assert A.dtype == np.int32 and A.shape == (12, 2) and tuple(A[7]) == (2, 1)
    # k2ij enumerates the strict upper triangle row by row (utils.cpp:87-97)
    n = 7
    assert [cp.utils.k2ij(k, n) for k in range(n * (n - 1) // 2)] == [(i, j) for i in range(n) for j in range(i + 1, n)]


def test_registration_factory_method_table():
    from roman_b200.align import SubmapAlignParams, ROMANRegistration, DistRegWithPruning
    expect = {  # method: (ratio_feature_dim, cos_feature_dim, gravity) — reference submap_align_params.py:98-108
        "clipper": (0, 0, False), "gravity": (0, 0, True), "pcavolgrav": (4, 0, True), "extentvolgrav": (4, 0, True),
        "roman": (4, 768, True), "sevg": (4, 768, True), "spv": (4, 0, False), "semanticgrav": (0, 768, True),
        "spvg": (4, 768, True), "roman_no_semantics": (4, 0, True)}
    for method, (r, c, g) in expect.items():
        reg = SubmapAlignParams(method=method).get_object_registration()
        assert isinstance(reg, ROMANRegistration)
        ip = reg.iparams
        assert (ip.ratio_feature_dim, ip.cos_feature_dim, bool(ip.gravity_guided)) == (r, c, g)
        assert (ip.sigma, ip.epsilon, ip.mindist, ip.cosine_min, ip.cosine_max) == (0.4, 0.6, 0.2, 0.5, 0.7)
        assert ip.gravity_unc_ang_rad == (0.0872665 if g else 0.0)
    assert isinstance(SubmapAlignParams(method="clipper+prune").get_object_registration(), DistRegWithPruning)
    with pytest.raises(AssertionError):
        SubmapAlignParams(method="nope").get_object_registration()


def test_object_row_packing_matches_reference_order():
    """[centre | lin, plan, scat | volume | sorted extent | descriptor] (roman_registration.py:98-108)"""
    from roman_b200.align import ROMANParams, ROMANRegistration, SimpleObject
    ob = SimpleObject(center=np.array([1.0, 2.0, 3.0]), volume=7.0, linearity=0.1, planarity=0.2, scattering=0.3,
                      extent=np.array([3.0, 1.0, 2.0]), semantic_descriptor=np.array([9.0, 8.0]))
    reg = ROMANRegistration(ROMANParams(volume=True, pca=True, extent=True, semantics_dim=2))
    assert reg._object_to_clipper_list(ob) == [1, 2, 3, 0.1, 0.2, 0.3, 7, 1, 2, 3, 9, 8]
    assert np.array_equal(reg.rows([ob, ob]), np.array([reg._object_to_clipper_list(ob)] * 2))  # vectorised packing
    table, counts = reg.pack([[ob], [], [ob, ob]])
    assert table.shape == (3, 12) and counts.tolist() == [1, 0, 2]
    reg = ROMANRegistration(ROMANParams(volume=True))
    assert reg._object_to_clipper_list(ob) == [1, 2, 3, 7]
    assert np.array_equal(reg.rows([ob]), [[1, 2, 3, 7]])
    assert reg.register([], [ob]).shape == (1, 0)  # np.array([[]]) for an empty map (object_registration.py:23-24)


def test_prune_mask_equals_reference_row_deletion():
    """DistRegWithPruning.putative_associations == np.delete on the all-to-all list (dist_reg_with_pruning.py:60-79)."""
    from roman_b200.align import DistRegWithPruning, SimpleObject
    from roman_b200.clipperpy.utils import create_all_to_all
    rng = np.random.default_rng(4)
    mk = lambda n: [SimpleObject(center=rng.normal(size=3), volume=rng.uniform(0.1, 2), linearity=rng.uniform(0.1, 1),
                                 planarity=rng.uniform(0.1, 1), scattering=rng.uniform(0.1, 1),
                                 semantic_descriptor=(lambda v: v / np.linalg.norm(v))(rng.normal(size=6) + 1.5))
                    for _ in range(n)]
    m1, m2 = mk(9), mk(11)
    reg = DistRegWithPruning(sigma=0.4, epsilon=0.6, shape_epsilon=0.3, cos_min=0.6)
    A_all = create_all_to_all(9, 11)
    d1 = np.array([o.semantic_descriptor for o in m1]); d2 = np.array([o.semantic_descriptor for o in m2])
    A_put = np.delete(A_all, ((d1 @ d2.T) < 0.6)[A_all[:, 0], A_all[:, 1]], axis=0)
    s = lambda objs, idx: np.array([[o.volume, o.linearity, o.planarity, o.scattering] for o in objs])[idx]
    comb = np.stack([s(m1, A_put[:, 0]), s(m2, A_put[:, 1])], axis=2)
    A_put = np.delete(A_put, np.any((comb.min(axis=2) / comb.max(axis=2)) < 0.3, axis=1), axis=0)
    assert np.array_equal(reg.putative_associations(m1, m2), A_put)


def test_synth_is_deterministic_and_shaped():
    from roman_b200 import synth
    a, b = synth.make_pair(5, 40, 50), synth.make_pair(5, 40, 50)
    assert np.array_equal(a.D1, b.D1) and np.array_equal(a.u0, b.u0)
    assert a.D1.shape == (40, 775) and a.D2.shape == (50, 775) and a.u0.shape == (2000,)
    assert a.n_inliers == 20
    assert np.allclose(np.linalg.norm(a.D1[:, 7:], axis=1), 1.0)
    got = (a.centres2[:20] @ a.T_gt[:3, :3].T) + a.T_gt[:3, 3]
    assert np.abs(got - a.centres1[:20]).max() < 0.3  # planted matches, 5 cm noise


def test_shard_roundtrip_single_process():
    from roman_b200 import dist as rd
    for n, world in [(10, 1), (10, 3), (7, 8), (0, 2)]:
        idx = [rd.shard_indices(n, world, r) for r in range(world)]
        assert sorted(np.concatenate(idx).tolist()) == list(range(n))
        assert max((len(i) for i in idx), default=0) <= rd.shard_size(n, world)


WORKER = r'''
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from roman_b200 import dist as rd
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
n = 11
mine = rd.shard_indices(n, world, rank)
# fake per-pair results that encode the global index, so the reassembly can be checked exactly
status = (mine % 2).astype(np.int32)
assoc = [np.stack([np.arange(g % 5 + 1), np.arange(g % 5 + 1) + g], axis=1).astype(np.int32) for g in mine]
T = np.tile(np.eye(4), (len(mine), 1, 1)); T[:, 0, 3] = mine
slots = rd.shard_size(n, world)
rec = rd.pack_records(status, [len(a) for a in assoc], mine * 0.5, T, slots)
inl = rd.pack_inliers(assoc, slots, 8)
out = rd.unshard(*((lambda r, i: (r, n, i))(*rd.all_gather_records(rec, inl))))
assert out["status"].tolist() == [g % 2 for g in range(n)]
assert np.allclose(out["score"], np.arange(n) * 0.5)
assert np.allclose(out["T"][:, 0, 3], np.arange(n))
for g in range(n):
    assert out["inliers"][g].tolist() == [[k, k + g] for k in range(g % 5 + 1)], (g, out["inliers"][g])
dist.barrier()
if rank == 0:
    print("GLOO_OK")
dist.destroy_process_group()
'''


def test_all_gather_of_records_world_size_2_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    port = 29500 + (os.getpid() % 2000)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", str(port), str(script), ROOT]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env={**os.environ, "OMP_NUM_THREADS": "1"})
    assert r.returncode == 0, r.stderr[-2000:]
    assert "GLOO_OK" in r.stdout


def test_loop_closure_wire_formats(tmp_path):
    """align.g2o / align.json (reference roman/align/results.py:149-197): record selection, pose chain, text layout."""
    import json
    from scipy.spatial.transform import Rotation as Rot
    from roman_b200.align.results import SubmapMeta, loop_closures, write_g2o, write_lc_json
    from roman_b200.align.submap_align import SubmapPairResults
    rng = np.random.default_rng(0)

    def rand_T():
        T = np.eye(4)
        T[:3, :3] = Rot.from_rotvec(rng.normal(size=3)).as_matrix()
        T[:3, 3] = rng.normal(size=3) * 5
        return T
    S0, S1 = 2, 3
    meta0 = [SubmapMeta(time=100.25 + 30 * i, pose_gravity_aligned=rand_T(), pose_flu=rand_T(), pose_index=10 * i) for i in range(S0)]
    meta1 = [SubmapMeta(time=120.5 + 30 * j, pose_gravity_aligned=rand_T(), pose_flu=rand_T(), pose_index=7 * j + 1) for j in range(S1)]
    T = np.stack([np.stack([rand_T() for _ in range(S1)]) for _ in range(S0)])
    n = np.array([[5.0, 3.0, np.nan], [4.0, 9.0, 0.0]])
    res = SubmapPairResults(T_ij_hat_mat=T, clipper_num_associations=n, clipper_percent_associations=n / 10,
                            associated_objs_mat=[[None] * S1 for _ in range(S0)], status_mat=np.zeros((S0, S1), dtype=np.int32))
    lcs = loop_closures(res, meta0, meta1, lc_association_thresh=4)
    assert [(c["i"], c["j"]) for c in lcs] == [(0, 0), (1, 0), (1, 1)]
    c = lcs[2]
    expect = np.linalg.inv(meta0[1].pose_flu) @ meta0[1].pose_gravity_aligned @ T[1, 1] @ np.linalg.inv(meta1[1].pose_gravity_aligned) @ meta1[1].pose_flu
    assert np.allclose(c["T_pi_pj"], expect)
    assert np.allclose(Rot.from_quat(c["q"]).as_matrix(), expect[:3, :3]) and np.allclose(c["t"], expect[:3, 3])
    # single-robot runs drop pairs that are close in time: (0,0) 20.25 s, (1,0) 9.75 s, (1,1) 20.25 s apart
    assert loop_closures(res, meta0, meta1, single_robot_lc=True, single_robot_lc_time_thresh=25.0) == []
    assert [(c["i"], c["j"]) for c in loop_closures(res, meta0, meta1, single_robot_lc=True, single_robot_lc_time_thresh=15.0)] == [(0, 0), (1, 1)]
    g2o, js = tmp_path / "align.g2o", tmp_path / "align.json"
    write_g2o(str(g2o), lcs)
    write_lc_json(str(js), lcs, robot_names=["sparkal1", "sparkal2"])
    lines = g2o.read_text().split("\n")
    assert lines[0] == "# LC: 5" and lines[2] == "# LC: 4" and lines[6] == ""
    head, tpart, qpart, *info = lines[1].split("\t")
    assert head == "EDGE_SE3:QUAT a0 b1 "
    assert np.allclose([float(x) for x in tpart.split()], lcs[0]["t"]) and np.allclose([float(x) for x in qpart.split()], lcs[0]["q"])
    assert [len(x.split()) for x in info[:6]] == [6, 5, 4, 3, 2, 1]
    assert float(info[0].split()[0]) == 1 / 0.5 ** 2 and abs(float(info[5].split()[0]) - 1 / np.deg2rad(0.5) ** 2) < 1e-6
    rec = json.loads(js.read_text())
    assert len(rec) == 3 and rec[0]["rotation_convention"] == "xyzw" and rec[0]["names"] == ["sparkal1", "sparkal2"]
    assert rec[0]["seconds"] == [100, 120] and rec[0]["nanoseconds"] == [250000000, 500000000]


def test_shim_covers_every_clipperpy_use_of_the_reference():
    """tests/golden/clipperpy_surface.json is an AST scan of the reference's Python sources for every use of the
    compiled `clipperpy` module; roman_b200.clipperpy must offer each of them (zero-change drop-in)."""
    import inspect
    import json
    from roman_b200 import clipperpy
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "clipperpy_surface.json")) as f:
        surf = json.load(f)
    assert len(surf["chains"]) >= 11 and len(surf["methods"]) >= 8
    for chain in surf["chains"]:
        obj = clipperpy
        for part in chain.split(".")[1:]:
            assert hasattr(obj, part), f"roman_b200.clipperpy lacks {chain}"
            obj = getattr(obj, part)
    for cls in (clipperpy.CLIPPER, clipperpy.CLIPPERPairwiseAndSingle):
        for m in surf["methods"]:
            if m == "score_pairwise_and_single_consistency" and cls is clipperpy.CLIPPER:
                continue
            assert callable(getattr(cls, m, None)), f"{cls.__name__}.{m} missing"
    rp = clipperpy.invariants.ROMANParams()
    for a in surf["params"]:
        assert hasattr(rp, a), f"ROMANParams.{a} missing"
    ep = clipperpy.invariants.EuclideanDistanceParams()
    assert all(hasattr(ep, a) for a in ("sigma", "epsilon", "mindist"))
    sol = clipperpy.Solution() if not inspect.signature(clipperpy.Solution).parameters else None
    fields = getattr(clipperpy.Solution, "__dataclass_fields__", None) or (vars(sol) if sol is not None else {})
    for a in surf["solution"]:
        assert a in fields or hasattr(clipperpy.Solution, a), f"Solution.{a} missing"
    # solve() takes the start vector as an optional argument, like the binding (py_clipper.cpp:289-297)
    assert "u0" in inspect.signature(clipperpy.CLIPPER.solve).parameters


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the reference algorithm timed on the host cores) needs no GPU, prints exactly one JSON
    line on stdout and carries the keys the driver reads."""
    import json
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "alignments/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["steps"] == 1 and d["n_gpus"] == 1
    from oracle import ref as oref      # "reference" when oracle/_ref (the reference's own sources) was built, else the port
    assert d["cpu_baseline"]["kind"] == ("reference" if oref.available() else "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert "workload" in d["config"]


def test_object_registration_glue_with_a_recording_clipper():
    """The Python orchestration of ObjectRegistration (register / get_MCA / mno_clipper / T_align preconditions) driven
    with a recording stand-in for the clipper object, so that it is exercised without a GPU."""
    from roman_b200.align import InsufficientAssociationsException, ObjectRegistration, SimpleObject
    from roman_b200.align import object_registration as oreg
    calls = []

    class FakeClipper:
        def score_pairwise_consistency(self, D1, D2, A):
            calls.append(("score", D1.shape, D2.shape, A.shape))
            self.A = A

        def solve(self, u0=None):
            calls.append(("solve", None if u0 is None else len(u0)))

        def get_selected_associations(self):
            return self.A[[0, 4]]

        def get_affinity_matrix(self):
            return np.eye(len(self.A))

        def get_constraint_matrix(self):
            return np.eye(len(self.A))

    class Reg(ObjectRegistration):
        def _setup_clipper(self):
            return FakeClipper()

        def _object_to_clipper_list(self, o):
            return np.asarray(o.center).reshape(-1)[:self.dim].tolist()

    reg = Reg(dim=3)
    m1 = [SimpleObject(center=np.array([k, 0.0, 1.0])) for k in range(2)]
    m2 = [SimpleObject(center=np.array([0.0, k, 2.0])) for k in range(3)]
    sel = reg.register(m1, m2, u0=np.ones(6))
    assert calls == [("score", (3, 2), (3, 3), (6, 2)), ("solve", 6)]          # (d, n) matrices, all-to-all list
    assert sel.tolist() == [[0, 0], [1, 1]]
    assert reg.register([], m2).shape == (1, 0)                                # the reference's np.array([[]])
    M, C, A = reg.get_MCA(m1, m2)
    assert M.shape == (6, 6) and A.tolist() == [[i, j] for i in range(2) for j in range(3)]
    with pytest.raises(InsufficientAssociationsException) as ex:
        reg.T_align(m1, [], None)
    assert str(ex.value) == "Insufficient associations. Map 1 length: 2. Map 2 length: 0. Associations: None"
    with pytest.raises(InsufficientAssociationsException) as ex:
        reg.T_align(m1, m2, np.array([[0, 0], [1, 1]]))
    assert ex.value.n_associations == 2 and str(ex.value).endswith("Associations: 2")


def test_euler_helpers_match_scipy_zyx():
    """zyx_euler / roll_pitch / rm_roll_pitch against scipy's Rotation.as_euler('ZYX') — the call the reference makes
    (roman/align/dist_reg_with_pruning.py:41, roman/utils.py:128-130) — on random rotations incl. large angles."""
    from scipy.spatial.transform import Rotation as Rot
    from roman_b200.align.dist_reg_with_pruning import zyx_euler
    from roman_b200.align.submap_align import rm_roll_pitch, roll_pitch
    R = Rot.random(200, random_state=7).as_matrix()
    T = np.tile(np.eye(4), (200, 1, 1))
    T[:, :3, :3] = R
    T[:, :3, 3] = np.random.default_rng(1).normal(size=(200, 3))
    ypr = Rot.from_matrix(R).as_euler("ZYX")
    roll, pitch = roll_pitch(T)
    assert np.abs(roll - ypr[:, 2]).max() < 1e-12 and np.abs(pitch - ypr[:, 1]).max() < 1e-12
    for k in range(0, 200, 17):
        assert np.allclose(zyx_euler(R[k]), ypr[k], atol=1e-12)
    yaw_only = rm_roll_pitch(T)
    want = Rot.from_euler("z", ypr[:, 0].reshape(-1, 1)).as_matrix()
    assert np.abs(yaw_only[:, :3, :3] - want).max() < 1e-12 and np.array_equal(yaw_only[:, :3, 3], T[:, :3, 3])


def test_example_script_dry_run():
    """examples/all_pairs_loop_closures.py up to (not including) the GPU registration: map synthesis, submap cutting."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "examples", "all_pairs_loop_closures.py"), "--dry-run"],
                       capture_output=True, text=True, timeout=300, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "submap pairs" in r.stdout and "dry run" in r.stdout


def test_bench_workloads_are_consistent():
    """bench.py's three configurations (BASELINE.json configs[1..4] stand-ins): shapes, shared tables, longest-first
    order of the all-pairs search, per-rank worlds, and the per-problem views the CPU arm and the parity check use."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    import bench
    wl = bench.build_workload("metric", 0, 1, batch=6)
    assert wl.B == 6 and not wl.indexed and (wl.m == 2000).all() and wl.D1.shape == (6 * 40, bench.D_ROW)
    D1, D2, u0 = wl.pair(3)
    assert D1.shape == (40, bench.D_ROW) and D2.shape == (50, bench.D_ROW) and u0.shape == (2000,)
    w5 = bench.build_workload("cfg5", 0, 4)
    assert w5.B == 400 and w5.indexed and w5.m.max() <= 1600 and (np.diff(w5.m) <= 0).all()      # longest problem first
    assert (w5.obj1 + w5.n1 <= len(w5.D1)).all() and (w5.obj2 + w5.n2 <= len(w5.D2)).all()
    D1, D2, u0 = w5.pair(17)
    assert len(u0) == len(D1) * len(D2) == w5.m[17]
    other = bench.build_workload("cfg5", 1, 4)                 # another rank searches another robot pair
    assert other.B == 400 and not np.array_equal(other.m, w5.m)
    t, why = bench.traffic_for("no-such-key")
    assert t is None and "no committed ncu capture" in why
