"""
Multi-GPU driver pieces: problems (submap pairs) are independent, so they are partitioned
round-robin over the ranks (one process per GPU) with NO data-path collective; the only exchange
is one all-gather of fixed-size per-pair records at the end:

    record = [status, n_inliers, F, reserved, T(16)]          20 float64
    inliers = (max_k, 2) int32, rows beyond n_inliers = -1     optional, same call pattern

The functions are backend-agnostic (NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist

RECORD_DOUBLES = 20


def shard_indices(n_items: int, world: int, rank: int) -> np.ndarray:
    """Round-robin partition: item g goes to rank g mod world."""
    return np.arange(rank, n_items, world, dtype=np.int64)


def shard_size(n_items: int, world: int) -> int:
    """Per-rank slot count (ranks with fewer items pad with status = -1 records)."""
    return (n_items + world - 1) // world


def pack_records(status, n_sel, score, T, slots: int, device=None) -> torch.Tensor:
    """(slots, 20) float64 records from per-pair results of this rank (torch tensors or arrays)."""
    as_t = lambda x, dt: x.to(dt) if isinstance(x, torch.Tensor) else torch.as_tensor(np.asarray(x)).to(dt)
    status, n_sel, score, T = as_t(status, torch.float64), as_t(n_sel, torch.float64), as_t(score, torch.float64), as_t(T, torch.float64)
    n = status.shape[0]
    dev = device if device is not None else status.device
    rec = torch.full((slots, RECORD_DOUBLES), float("nan"), dtype=torch.float64, device=dev)
    rec[:, 0] = -1.0  # empty slot
    rec[:n, 0] = status.to(dev)
    rec[:n, 1] = n_sel.to(dev)
    rec[:n, 2] = score.to(dev)
    rec[:n, 3] = 0.0
    rec[:n, 4:] = T.reshape(n, 16).to(dev)
    return rec


def pack_inliers(assoc: Sequence[np.ndarray], slots: int, max_k: int, device=None) -> torch.Tensor:
    out = torch.full((slots, max_k, 2), -1, dtype=torch.int32)
    for b, a in enumerate(assoc):
        a = np.asarray(a).reshape(-1, 2)[:max_k]
        if len(a):
            out[b, :len(a)] = torch.from_numpy(a.astype(np.int32))
    return out.to(device) if device is not None else out


def all_gather_records(rec: torch.Tensor, inl: Optional[torch.Tensor] = None):
    """One all-gather of the fixed-size records (and, if given, the padded inlier lists)."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    if world == 1:
        return rec.unsqueeze(0), (None if inl is None else inl.unsqueeze(0))
    # concatenated-along-dim-0 output layout: accepted by both NCCL and gloo
    g_rec = torch.empty((world * rec.shape[0],) + tuple(rec.shape[1:]), dtype=rec.dtype, device=rec.device)
    dist.all_gather_into_tensor(g_rec, rec.contiguous())
    g_rec = g_rec.view((world,) + tuple(rec.shape))
    g_inl = None
    if inl is not None:
        g_inl = torch.empty((world * inl.shape[0],) + tuple(inl.shape[1:]), dtype=inl.dtype, device=inl.device)
        dist.all_gather_into_tensor(g_inl, inl.contiguous())
        g_inl = g_inl.view((world,) + tuple(inl.shape))
    return g_rec, g_inl


def unshard(g_rec: torch.Tensor, n_items: int, g_inl: Optional[torch.Tensor] = None):
    """Back to global problem order.  Returns dict(status, n_inliers, score, T, [inliers])."""
    world, slots = g_rec.shape[0], g_rec.shape[1]
    rec = g_rec.cpu().numpy()
    order = np.empty((n_items, 2), dtype=np.int64)
    for g in range(n_items):
        order[g] = (g % world, g // world)
    assert order[:, 1].max(initial=0) < max(slots, 1)
    r = rec[order[:, 0], order[:, 1]]
    out = dict(status=r[:, 0].astype(np.int32), n_inliers=r[:, 1].astype(np.int32), score=r[:, 2].copy(),
               T=r[:, 4:].reshape(n_items, 4, 4).copy())
    if g_inl is not None:
        inl = g_inl.cpu().numpy()[order[:, 0], order[:, 1]]
        out["inliers"] = [inl[g, :out["n_inliers"][g]].copy() for g in range(n_items)]
    return out


def register_pairs_distributed(registration, pairs: Sequence[Tuple[List, List]], u0: Optional[Sequence[np.ndarray]] = None,
                               max_k: Optional[int] = None, device=None):
    """
    All-pairs loop-closure search over the ranks: each rank registers the pairs g with
    g mod world == rank through `registration.register_batch`, then one all-gather makes the
    complete result table available on every rank, in the order of `pairs`.
    `max_k` (rows of the padded inlier lists) defaults to the largest inlier count over all ranks (one tiny
    all-reduce), so no list is ever truncated; a caller-given max_k smaller than a list raises.
    """
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    mine = shard_indices(len(pairs), world, rank)
    assoc, T, status = registration.register_batch([pairs[g] for g in mine],
                                                   None if u0 is None else [u0[g] for g in mine])
    slots = shard_size(len(pairs), world)
    n_sel = np.array([len(a) for a in assoc], dtype=np.int32)
    T4 = np.tile(np.eye(4), (len(mine), 1, 1))
    if len(mine):
        k = T.shape[1]
        T4[:, :k - 1, :k - 1] = T[:, :k - 1, :k - 1]
        T4[:, :k - 1, 3] = T[:, :k - 1, k - 1]
    k_local = int(n_sel.max(initial=0))
    if max_k is None:
        kt = torch.tensor([k_local], dtype=torch.int64, device=device if device is not None else "cpu")
        if world > 1:
            dist.all_reduce(kt, op=dist.ReduceOp.MAX)
        max_k = max(int(kt.item()), 1)
    elif k_local > max_k:
        raise ValueError(f"max_k = {max_k} is smaller than an inlier list of this rank ({k_local} associations)")
    scores = getattr(registration, "last_scores", None)
    rec = pack_records(status, n_sel, np.zeros(len(mine)) if scores is None or len(scores) != len(mine) else scores, T4, slots, device=device)
    inl = pack_inliers(assoc, slots, max_k, device=device)
    g_rec, g_inl = all_gather_records(rec, inl)
    return unshard(g_rec, len(pairs), g_inl)
