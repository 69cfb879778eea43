"""
Engine — Python host object over one rb_handle: owns the batch plan, holds PyTorch device tensors
and calls the CUDA path through the ctypes C ABI on torch's current stream.

PyTorch is plumbing only here (device memory, streams); all arithmetic of the hot path runs in
libroman_b200.so.  There is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib
from ._lib import RbInvariantParams, RbSolverParams


class RomanB200Error(RuntimeError):
    pass


# Packed-affinity format new Engines start in: False = dense (every tile written and streamed, the library default),
# True = tile skip (the sparse fast path, see rb_set_tile_skip).  set_default_tile_skip() changes it process-wide.
DEFAULT_TILE_SKIP = False


def set_default_tile_skip(on: bool):
    global DEFAULT_TILE_SKIP
    DEFAULT_TILE_SKIP = bool(on)


def default_invariant_params() -> RbInvariantParams:
    p = RbInvariantParams()
    _lib.lib().rb_default_invariant_params(C.byref(p))
    return p


def default_solver_params() -> RbSolverParams:
    p = RbSolverParams()
    _lib.lib().rb_default_solver_params(C.byref(p))
    return p


@dataclass
class BatchResult:
    """Device-resident results of one batch (all torch tensors on the engine's device)."""
    sel: torch.Tensor        # (sum m, 2) int32; rows [off[b], off[b]+n_sel[b]) are problem b's inliers
    n_sel: torch.Tensor      # (B,) int32
    T: torch.Tensor          # (B, 4, 4) float64
    status: torch.Tensor     # (B,) int32  (0 ok, 1 insufficient associations)
    score: torch.Tensor      # (B,) float64
    n_passes: torch.Tensor   # (B,) int32
    offsets: np.ndarray      # (B+1,) int64 host association offsets

    def associations(self, b: int) -> np.ndarray:
        n = int(self.n_sel[b])
        o = int(self.offsets[b])
        return self.sel[o:o + n].cpu().numpy()


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


class Engine:
    def __init__(self, device: int | torch.device | None = None):
        if not torch.cuda.is_available():
            raise RomanB200Error("roman_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        if device is None:
            device = torch.cuda.current_device()
        self.device = torch.device("cuda", device if isinstance(device, int) else device.index or 0)
        self._lib = _lib.lib()
        h = C.c_void_p()
        rc = self._lib.rb_create(C.byref(h), self.device.index)
        if rc != 0:
            raise RomanB200Error(f"rb_create failed ({rc})")
        self._h = h
        if DEFAULT_TILE_SKIP:
            self._lib.rb_set_tile_skip(self._h, 1)
        self.B = 0
        self.d = 0
        self.offsets = np.zeros(1, dtype=np.int64)
        self.obj1_offsets = np.zeros(1, dtype=np.int64)
        self.obj2_offsets = np.zeros(1, dtype=np.int64)
        self.all_to_all = True

    # ------------------------------------------------------------------ lifetime / errors
    def close(self):
        if getattr(self, "_h", None):
            self._lib.rb_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int):
        if rc != 0:
            msg = self._lib.rb_last_error(self._h)
            raise RomanB200Error(f"libroman_b200 error {rc}: {msg.decode() if msg else ''}")

    @property
    def sm_count(self) -> int:
        return int(self._lib.rb_device_sm_count(self._h))

    # ------------------------------------------------------------------ configuration
    def set_invariant_params(self, p: RbInvariantParams):
        self._check(self._lib.rb_set_invariant_params(self._h, C.byref(p)))

    def set_solver_params(self, p: RbSolverParams):
        self._check(self._lib.rb_set_solver_params(self._h, C.byref(p)))

    def set_solver_cluster(self, ctas_per_problem: int):
        """CTAs (SMs) per problem in the solver: 0 auto, 1, 2, 4, 8; applies at the next plan()."""
        self._check(self._lib.rb_set_solver_cluster(self._h, int(ctas_per_problem)))

    @property
    def solver_cluster(self) -> int:
        return int(self._lib.rb_get_solver_cluster(self._h))

    def set_tile_skip(self, on: bool):
        """Packed-affinity format: False (default) = dense, True = tile skip (stores/streams only tiles holding a non-zero,
        resident solves).
        Changing it invalidates the plan: call plan() again."""
        self._check(self._lib.rb_set_tile_skip(self._h, int(on)))

    def stored_tiles(self) -> torch.Tensor:
        """(B,) int32: 32x32 tiles actually stored per problem (tile-skip format)."""
        out = torch.zeros(max(self.B, 1), dtype=torch.int32, device=self.device)
        self._check(self._lib.rb_stored_tiles(self._h, _ptr(out), self._stream()))
        return out

    def live_counts(self) -> torch.Tensor:
        """(B,) int32: associations per problem that can carry an edge (single score not below affinityeps)."""
        out = torch.zeros(max(self.B, 1), dtype=torch.int32, device=self.device)
        self._check(self._lib.rb_live_counts(self._h, _ptr(out), self._stream()))
        return out[:self.B]

    def host_copy_bytes(self):
        """(h2d, d2h) bytes moved by the last align_batch_host call."""
        return int(self._lib.rb_host_copy_bytes(self._h, 0)), int(self._lib.rb_host_copy_bytes(self._h, 1))

    def edge_counts(self) -> torch.Tensor:
        """(B,) int32: edges K1 listed per problem (tile-skip format)."""
        out = torch.zeros(max(self.B, 1), dtype=torch.int32, device=self.device)
        self._check(self._lib.rb_edge_counts(self._h, _ptr(out), self._stream()))
        return out

    def solver_resident(self) -> torch.Tensor:
        """(B,) int32: 0 tiles streamed every pass, 1 resident (CSR from one pass over the tiles), 2 resident (CSR from K1's edge list)."""
        out = torch.zeros(max(self.B, 1), dtype=torch.int32, device=self.device)
        self._check(self._lib.rb_solver_resident(self._h, _ptr(out), self._stream()))
        return out

    def set_profiling(self, on: bool):
        self._check(self._lib.rb_set_profiling(self._h, int(on)))

    def solver_streamed_passes(self) -> torch.Tensor:
        """(B,) int32: passes of the last solve that read the packed matrix from HBM (rb_solver_streamed_passes)."""
        out = torch.zeros(max(self.B, 1), dtype=torch.int32, device=self.device)
        self._check(self._lib.rb_solver_streamed_passes(self._h, _ptr(out), self._stream()))
        return out[:self.B]

    def set_solver_promote(self, passes: int):
        """Dense format: streamed passes after which a still-iterating problem is made resident (0 = never; default 8)."""
        self._check(self._lib.rb_set_solver_promote(self._h, int(passes)))

    def stage_ms(self) -> dict:
        ms = (C.c_float * 4)()
        self._check(self._lib.rb_get_stage_ms(self._h, ms))
        return dict(prep=ms[0], affinity=ms[1], solve=ms[2], arun=ms[3])

    def plan(self, d: int, n1: Sequence[int], n2: Sequence[int], m: Optional[Sequence[int]] = None,
             obj1_offsets: Optional[Sequence[int]] = None, obj2_offsets: Optional[Sequence[int]] = None,
             total_obj1: Optional[int] = None, total_obj2: Optional[int] = None, share: Optional[Sequence[int]] = None):
        """Describe the batch shapes (host side).  m=None: all-to-all associations.
        obj*_offsets: problems reference row ranges of shared object tables (rb_plan_indexed).
        share: share[b] = index of the problem whose affinity matrix problem b uses (rb_plan_shared; multi-start)."""
        n1a = np.ascontiguousarray(n1, dtype=np.int32)
        n2a = np.ascontiguousarray(n2, dtype=np.int32)
        B = len(n1a)
        assert len(n2a) == B
        ma = None if m is None else np.ascontiguousarray(m, dtype=np.int32)
        if obj1_offsets is None:
            self._check(self._lib.rb_plan(self._h, B, int(d), n1a.ctypes.data, n2a.ctypes.data,
                                          None if ma is None else ma.ctypes.data))
        else:
            o1 = np.ascontiguousarray(obj1_offsets, dtype=np.int64)
            o2 = np.ascontiguousarray(obj2_offsets, dtype=np.int64)
            assert len(o1) == B and len(o2) == B
            if share is None:
                self._check(self._lib.rb_plan_indexed(self._h, B, int(d), n1a.ctypes.data, n2a.ctypes.data,
                                                      None if ma is None else ma.ctypes.data, o1.ctypes.data,
                                                      o2.ctypes.data, int(total_obj1), int(total_obj2)))
            else:
                sh = np.ascontiguousarray(share, dtype=np.int32)
                assert len(sh) == B
                self._check(self._lib.rb_plan_shared(self._h, B, int(d), n1a.ctypes.data, n2a.ctypes.data,
                                                     None if ma is None else ma.ctypes.data, o1.ctypes.data,
                                                     o2.ctypes.data, int(total_obj1), int(total_obj2), sh.ctypes.data))
        self.B, self.d, self.all_to_all = B, int(d), ma is None
        self.offsets = np.array([self._lib.rb_assoc_offset(self._h, b) for b in range(B + 1)], dtype=np.int64)
        self.obj1_offsets = np.array([self._lib.rb_obj1_offset(self._h, b) for b in range(B + 1)], dtype=np.int64)
        self.obj2_offsets = np.array([self._lib.rb_obj2_offset(self._h, b) for b in range(B + 1)], dtype=np.int64)
        self.total_obj1, self.total_obj2 = int(self.obj1_offsets[-1]), int(self.obj2_offsets[-1])
        self.n1, self.n2 = n1a, n2a
        return self

    @property
    def total_assoc(self) -> int:
        return int(self.offsets[-1])

    @property
    def packed_bytes(self) -> int:
        return int(self._lib.rb_packed_bytes(self._h))

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _dev(self, x, dtype):
        if isinstance(x, torch.Tensor):
            assert x.device == self.device and x.dtype == dtype and x.is_contiguous(), \
                f"expected contiguous {dtype} tensor on {self.device}"
            return x
        return torch.as_tensor(np.ascontiguousarray(x), dtype=dtype).to(self.device)

    # ------------------------------------------------------------------ stages (device tensors)
    def affinity(self, D1, D2, A=None):
        D1 = self._dev(D1, torch.float64)
        D2 = self._dev(D2, torch.float64)
        assert D1.numel() == self.total_obj1 * self.d and D2.numel() == self.total_obj2 * self.d
        if not self.all_to_all:
            A = self._dev(A, torch.int32)
            assert A.numel() == 2 * self.total_assoc
        self._keep = (D1, D2, A)
        self._check(self._lib.rb_affinity(self._h, _ptr(D1), _ptr(D2), _ptr(A) if not self.all_to_all else None,
                                          self._stream()))

    def solve(self, u0, want_u: bool = False):
        u0 = self._dev(u0, torch.float64)
        assert u0.numel() == self.total_assoc
        n = max(self.total_assoc, 1)
        B = max(self.B, 1)
        dev = self.device
        out = dict(
            nodes=torch.empty(n, dtype=torch.int32, device=dev),
            sel=torch.empty((n, 2), dtype=torch.int32, device=dev),
            n_nodes=torch.zeros(B, dtype=torch.int32, device=dev),
            u=torch.empty(n, dtype=torch.float64, device=dev) if want_u else None,
            score=torch.zeros(B, dtype=torch.float64, device=dev),
            ifinal=torch.zeros(B, dtype=torch.int32, device=dev),
            n_passes=torch.zeros(B, dtype=torch.int32, device=dev),
        )
        self._check(self._lib.rb_solve(self._h, _ptr(u0), _ptr(out["nodes"]), _ptr(out["sel"]), _ptr(out["n_nodes"]),
                                       _ptr(out["u"]), _ptr(out["score"]), _ptr(out["ifinal"]),
                                       _ptr(out["n_passes"]), self._stream()))
        return out

    def submap_similarity(self, frames0: Sequence, frames1: Sequence):
        """(S0, S1) float64 device tensor: max pairwise cosine between the stacked frame descriptors of every
        submap of run 0 and every submap of run 1 (rb_submap_similarity; reference roman/map/map.py:145-162).
        frames0[i]: (N_i, D) or (D,) array-like."""
        def table(frames):
            rows = [np.atleast_2d(np.asarray(f, dtype=np.float64)) for f in frames]
            if any(r.shape[0] == 0 for r in rows):
                raise ValueError("submap without a descriptor (numpy raises on the empty max as well)")
            off = np.concatenate([[0], np.cumsum([r.shape[0] for r in rows])]).astype(np.int64)
            return (np.concatenate(rows) if rows else np.zeros((0, 1))), off
        F0, off0 = table(frames0)
        F1, off1 = table(frames1)
        S0, S1 = len(off0) - 1, len(off1) - 1
        sim = torch.empty((S0, S1), dtype=torch.float64, device=self.device)
        if S0 == 0 or S1 == 0:
            return sim
        if F0.shape[1] != F1.shape[1]:
            raise ValueError("descriptor lengths differ")
        t = [self._dev(F0, torch.float64), self._dev(F1, torch.float64), self._dev(off0, torch.int64),
             self._dev(off1, torch.int64)]
        self._check(self._lib.rb_submap_similarity(self._h, S0, S1, int(F0.shape[1]), _ptr(t[2]), _ptr(t[3]),
                                                   _ptr(t[0]), _ptr(t[1]), _ptr(sim), self._stream()))
        self._keep_sim = t
        return sim

    def arun(self, D1, D2, sel, n_sel):
        D1 = self._dev(D1, torch.float64)
        D2 = self._dev(D2, torch.float64)
        sel = self._dev(sel, torch.int32)
        n_sel = self._dev(n_sel, torch.int32)
        B = max(self.B, 1)
        T = torch.empty((B, 4, 4), dtype=torch.float64, device=self.device)
        status = torch.zeros(B, dtype=torch.int32, device=self.device)
        self._check(self._lib.rb_arun(self._h, _ptr(D1), _ptr(D2), _ptr(sel), _ptr(n_sel), _ptr(T), _ptr(status),
                                      self._stream()))
        return T, status

    def align_batch(self, D1, D2, u0, A=None, out: Optional[BatchResult] = None) -> BatchResult:
        """affinity + solve + Arun for the planned batch; inputs/outputs stay on the device."""
        D1 = self._dev(D1, torch.float64)
        D2 = self._dev(D2, torch.float64)
        u0 = self._dev(u0, torch.float64)
        if not self.all_to_all:
            A = self._dev(A, torch.int32)
        if out is None:
            out = self.alloc_result()
        self._keep = (D1, D2, A, u0)
        self._check(self._lib.rb_align_batch(
            self._h, _ptr(D1), _ptr(D2), _ptr(A) if not self.all_to_all else None, _ptr(u0), _ptr(out.sel),
            _ptr(out.n_sel), _ptr(out.T), _ptr(out.status), _ptr(out.score), _ptr(out.n_passes), self._stream()))
        return out

    def alloc_result(self) -> BatchResult:
        n = max(self.total_assoc, 1)
        B = max(self.B, 1)
        dev = self.device
        return BatchResult(
            sel=torch.empty((n, 2), dtype=torch.int32, device=dev),
            n_sel=torch.zeros(B, dtype=torch.int32, device=dev),
            T=torch.empty((B, 4, 4), dtype=torch.float64, device=dev),
            status=torch.zeros(B, dtype=torch.int32, device=dev),
            score=torch.zeros(B, dtype=torch.float64, device=dev),
            n_passes=torch.zeros(B, dtype=torch.int32, device=dev),
            offsets=self.offsets.copy())

    def get_dense(self, b: int, want_M: bool = True, want_C: bool = True):
        m = int(self.offsets[b + 1] - self.offsets[b])
        M = torch.empty((m, m), dtype=torch.float64, device=self.device) if want_M else None
        Cm = torch.empty((m, m), dtype=torch.float64, device=self.device) if want_C else None
        self._check(self._lib.rb_get_dense(self._h, b, _ptr(M), _ptr(Cm), self._stream()))
        return M, Cm

    def set_dense(self, b: int, M, Cm=None):
        M = self._dev(M, torch.float64)
        Cm = None if Cm is None else self._dev(Cm, torch.float64)
        self._check(self._lib.rb_set_dense(self._h, b, _ptr(M), _ptr(Cm), self._stream()))

    # ------------------------------------------------------------------ host-buffer path (C ABI does the copies)
    def align_batch_host(self, D1: np.ndarray, D2: np.ndarray, u0: np.ndarray, A: Optional[np.ndarray] = None,
                         out: Optional[dict] = None) -> dict:
        """numpy (or pinned torch CPU) buffers in, numpy out, through rb_align_batch_host."""
        def hp(x):
            if x is None:
                return None
            if isinstance(x, torch.Tensor):
                return C.c_void_p(x.data_ptr())
            return C.c_void_p(x.ctypes.data)
        n = max(self.total_assoc, 1)
        B = max(self.B, 1)
        if out is None:
            out = dict(sel=np.empty((n, 2), dtype=np.int32), n_sel=np.zeros(B, dtype=np.int32),
                       T=np.empty((B, 4, 4)), status=np.zeros(B, dtype=np.int32), score=np.zeros(B),
                       n_passes=np.zeros(B, dtype=np.int32))
        self._check(self._lib.rb_align_batch_host(self._h, hp(D1), hp(D2), hp(A), hp(u0), hp(out["sel"]),
                                                  hp(out["n_sel"]), hp(out["T"]), hp(out["status"]),
                                                  hp(out["score"]), hp(out["n_passes"])))
        return out


def make_invariant_params(kind: int = 0, point_dim: int = 3, ratio_feature_dim: int = 0, cos_feature_dim: int = 0,
                          fusion_method: int = 0, sigma: float = 0.01, epsilon: float = 0.06, mindist: float = 0.0,
                          distance_weight: float = 1.0, ratio_weight: float = 1.0, cosine_weight: float = 1.0,
                          ratio_epsilon=None, cosine_min: float = 0.85, cosine_max: float = 0.95,
                          gravity_guided: bool = False, gravity_unc_ang_rad: float = 0.0, drift_aware: bool = False,
                          drift_scale_sigma: bool = False, drift_scale: float = 0.1) -> RbInvariantParams:
    """Keyword constructor with the reference defaults (roman.h:27-53)."""
    p = RbInvariantParams()
    p.kind, p.point_dim, p.ratio_feature_dim, p.cos_feature_dim = kind, point_dim, ratio_feature_dim, cos_feature_dim
    p.fusion_method, p.gravity_guided = int(fusion_method), int(gravity_guided)
    p.drift_aware, p.drift_scale_sigma = int(drift_aware), int(drift_scale_sigma)
    p.sigma, p.epsilon, p.mindist = sigma, epsilon, mindist
    p.distance_weight, p.ratio_weight, p.cosine_weight = distance_weight, ratio_weight, cosine_weight
    re = np.zeros(_lib.RB_MAX_RATIO_FEATURES)
    if ratio_epsilon is not None:
        r = np.asarray(ratio_epsilon, dtype=np.float64).reshape(-1)
        re[:len(r)] = r
    for i in range(_lib.RB_MAX_RATIO_FEATURES):
        p.ratio_epsilon[i] = re[i]
    p.cosine_min, p.cosine_max = cosine_min, cosine_max
    p.gravity_unc_ang_rad, p.drift_scale = gravity_unc_ang_rad, drift_scale
    return p


def method_invariant_kwargs(method: str = "roman", dim: int = 3, sigma: float = 0.4, epsilon: float = 0.6,
                            mindist: float = 0.2, epsilon_shape: float = 0.0, cosine_min: float = 0.5,
                            cosine_max: float = 0.7, semantics_dim: int = 768,
                            gravity_unc_ang_rad: float = 0.0872665) -> dict:
    """
    SubmapAlignParams.get_object_registration + ROMANRegistration.__init__ as one table
    (reference roman/params/submap_align_params.py:86-125, roman/align/roman_registration.py:38-80):
    method name -> ROMAN invariant parameters, with SubmapAlignParams' defaults.
    """
    from .synth import METHOD_FLAGS
    gravity, volume, pca, extent, semantics = METHOD_FLAGS[method]
    r = 3 * pca + 1 * volume + 3 * extent
    return dict(kind=0, point_dim=dim, ratio_feature_dim=r, cos_feature_dim=semantics_dim if semantics else 0,
                sigma=sigma, epsilon=epsilon, mindist=mindist, ratio_epsilon=np.ones(r) * epsilon_shape,
                cosine_min=cosine_min, cosine_max=cosine_max, gravity_guided=gravity,
                gravity_unc_ang_rad=gravity_unc_ang_rad if gravity else 0.0)
