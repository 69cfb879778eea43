#!/usr/bin/env python
"""
bench.py — submap-pair alignments/sec on the BASELINE.json workload.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--batch B]

One "step" = one pass of the hot path (K1a + K1 + K2 + K3) over one batch of B synthetic submap
pairs per GPU.  Workload at every N (weak scaling, B pairs per GPU): n1=40, n2=50 objects
(m = 2000 all-to-all correspondences), method 'roman' (gravity + PCA + volume + 768-d semantics),
SubmapAlignParams defaults — the configuration BASELINE.json's metric is quoted on.
Printed JSON line: see the task contract; everything at the top level is measured in the library's
default DENSE format (every tile of the packed affinity matrix written by K1 and streamed by K2 each
pass): `roofline` is for the dominant kernel (K2), `kernels` lists K1 and K2 separately.  `fast_path`
repeats value / e2e / per-kernel times with the opt-in sparse fast path (rb_set_tile_skip).

--impl reference times the reference algorithm's CPU restatement (oracle/, kind "port": the
reference itself needs Eigen3 and cannot be compiled in this image) on all host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N1, N2, SEM_DIM, METHOD = 40, 50, 768, "roman"
M_ASSOC = N1 * N2
WORKLOAD = "synthetic submap pairs n1=40 n2=50 (m=2000 all-to-all), method roman, D=768, SubmapAlignParams defaults"
METRIC = "submap-pair alignments/sec at N=2000 correspondences"
UNIT = "alignments/s"


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.rows = []
        self.first = 0
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.idx}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def wait_ready(self, timeout=5.0):
        """nvidia-smi takes a while to print its first line; do not start timing before it samples."""
        t0 = time.time()
        while self.proc and not self.rows and time.time() - t0 < timeout:
            time.sleep(0.02)

    def mark(self):
        self.first = len(self.rows)

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in (self.rows[self.first:] or self.rows[-1:]):
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax = float(f[2])
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


def make_inputs(seed0, B):
    from roman_b200 import synth
    D1, D2, u0, _ = synth.make_batch(seed0, B, N1, N2, method=METHOD, semantics_dim=SEM_DIM)
    return np.ascontiguousarray(D1), np.ascontiguousarray(D2), np.ascontiguousarray(u0)


def oracle_params():
    from oracle import clipper_np as cn
    from roman_b200.engine import method_invariant_kwargs
    kw = method_invariant_kwargs(METHOD, semantics_dim=SEM_DIM)
    return cn.InvariantParams(**kw), cn.SolverParams()


def cpu_baseline(rounds=16, threads=0, seed0=10_000):
    """
    Time the CPU restatement, one pair per thread on all host threads (throughput-fair mode of
    BASELINE.md §3): `rounds` sweeps over 2*cores distinct pairs (~1 s of wall, cores x ~1 s of CPU work).
    """
    import oracle
    O = oracle.COracle()
    ip, sp = oracle_params()
    if threads <= 0:
        threads = host_threads()  # all logical CPUs this process may use (torchrun would pin OMP_NUM_THREADS=1)
    cores = threads
    nd = 2 * cores
    D1, D2, u0 = make_inputs(seed0, nd)
    O.align_batch(ip, sp, D1[:cores], D2[:cores], u0[:cores], threads=threads)  # warm the library / thread pool
    t0 = time.perf_counter()
    for _ in range(rounds):
        r = O.align_batch(ip, sp, D1, D2, u0, threads=threads)
    dt = time.perf_counter() - t0
    n = rounds * nd
    return {"value": n / dt, "unit": UNIT, "cores": int(r["threads"]), "kind": "port",
            "sample": f"{n} alignments ({rounds} sweeps over {nd} distinct pairs of the same workload), one pair per "
                      f"thread (OpenMP over pairs), {dt:.2f} s wall",
            "host_cpus": os.cpu_count(), "pairs": n, "seconds": dt}, r


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    rounds = 8
    times = []
    last = None
    for s in range(args.warmup + args.steps):
        cb, _ = cpu_baseline(rounds, seed0=20_000)
        if s >= args.warmup:
            times.append(cb["seconds"])
            last = cb
    sample = last["pairs"]
    t = float(np.mean(times))
    v = sample / t
    cb = dict(last)
    cb["value"] = v
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "pairs_per_step": sample, "note": "CPU restatement of the reference "
                       "(oracle/clipper_oracle.c, kind=port: the reference needs Eigen3 which is absent); one pair per host thread"},
            "cpu_baseline": cb,
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def host_threads():
    """Physical cores this process may use: the restatement is memory-bound, SMT siblings only slow it down
    (measured on the 2 x 32-core host: 64 threads 400-600 alignments/s, 128 threads 250/s)."""
    try:
        allowed = os.sched_getaffinity(0)
    except AttributeError:
        return os.cpu_count() or 1
    cores = set()
    try:
        phys = core = None
        cpu = None
        for line in open("/proc/cpuinfo"):
            if line.startswith("processor"):
                cpu = int(line.split(":")[1])
            elif line.startswith("physical id"):
                phys = int(line.split(":")[1])
            elif line.startswith("core id"):
                core = int(line.split(":")[1])
            elif not line.strip() and cpu is not None:
                if cpu in allowed:
                    cores.add((phys, core))
                cpu = None
    except OSError:
        pass
    return len(cores) if cores else len(allowed)


def run_b200(args):
    import torch
    import torch.distributed as dist
    from roman_b200.engine import Engine, make_invariant_params, method_invariant_kwargs

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B = args.batch
    K, W = args.steps, args.warmup
    d_row = 3 + 3 + 1 + SEM_DIM
    m = M_ASSOC

    eng = Engine(local)
    eng.set_invariant_params(make_invariant_params(**method_invariant_kwargs(METHOD, semantics_dim=SEM_DIM)))
    if args.promote >= 0:
        eng.set_solver_promote(args.promote)
    eng.plan(d_row, [N1] * B, [N2] * B)
    # problems are partitioned round-robin over ranks: global pair g = b*world + rank
    D1h, D2h, u0h = make_inputs(1000 + (rank + args.rank_offset) * B, B)
    pin = lambda a: torch.from_numpy(a).pin_memory()
    D1p, D2p, u0p = pin(D1h.reshape(-1)), pin(D2h.reshape(-1)), pin(u0h.reshape(-1))
    D1d, D2d, u0d = D1p.to(dev), D2p.to(dev), u0p.to(dev)
    res = eng.alloc_result()
    rec = torch.zeros((B, 20), dtype=torch.float64, device=dev)       # status, n_inliers, F, pad, T[16]
    gathered = torch.zeros((world * B, 20), dtype=torch.float64, device=dev) if world > 1 else None  # (rank-major)

    def step():
        eng.align_batch(D1d, D2d, u0d, out=res)
        if world > 1:  # the path's only exchange: all-gather of fixed-size per-pair records
            rec[:, 0] = res.status.to(torch.float64)
            rec[:, 1] = res.n_sel.to(torch.float64)
            rec[:, 2] = res.score
            rec[:, 4:] = res.T.reshape(B, 16)
            dist.all_gather_into_tensor(gathered, rec)

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(tile_skip, sampler=None):
        """W warm-ups, then exactly K steps between device events; max over ranks."""
        eng.set_tile_skip(tile_skip)
        eng.plan(d_row, [N1] * B, [N2] * B)  # the format decides the solver's cluster size; planning is untimed
        for _ in range(W):
            step()
        sync()
        if sampler is not None:
            sampler.mark()  # only samples taken from here on (the timed region) are reported
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        stage = {"prep": 0.0, "affinity": 0.0, "solve": 0.0, "arun": 0.0}
        sync()
        ev0.record()
        for _ in range(K):
            step()
            for k, v in eng.stage_ms().items():  # CUDA events on the launch stream (syncs on the step's last event)
                stage[k] += v / K
        ev1.record()
        sync()
        t = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), stage

    def e2e_run():
        """K calls of the host-buffer C ABI entry point: pinned host inputs, H2D + kernels + D2H inside the timed region."""
        out_h = None
        sync()
        for _ in range(max(1, min(W, 3))):
            out_h = eng.align_batch_host(D1p, D2p, u0p, out=out_h)
        sync()
        t0 = time.perf_counter()
        for _ in range(K):
            out_h = eng.align_batch_host(D1p, D2p, u0p, out=out_h)
        wall = time.perf_counter() - t0
        te = torch.tensor([wall], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        return world * B * K / float(te.item()), out_h

    eng.set_profiling(True)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        sampler.wait_ready()
    # ---- main measurement: the library's default (dense) format — every tile of the packed upper triangle is written
    #      by K1 and streamed by K2 each pass: the kernels and byte counts of BASELINE.json's north_star / SURVEY 8d
    ms, stage = timed(False, sampler if rank == 0 else None)
    clocks = sampler.stop() if rank == 0 else None
    value = world * B * K / (ms * 1e-3)
    passes = res.n_passes.cpu().numpy().astype(np.int64)
    streamed = eng.solver_streamed_passes().cpu().numpy().astype(np.int64)   # passes that read M from HBM
    promoted = int((eng.solver_resident().cpu().numpy()[:B] == 3).sum())
    cluster_dense = eng.solver_cluster
    e2e_value, out_h = e2e_run()
    h2d = (D1p.numel() + D2p.numel() + u0p.numel()) * 8
    d2h = sum(v.nbytes for v in out_h.values())

    # ---- same batch through the opt-in sparse fast path (rb_set_tile_skip): only tiles holding a non-zero are stored,
    #      K1 lists the edges, K2 solves from a shared-memory CSR
    do_fast = (args.format == "both")
    if do_fast:
        ms_f, stage_f = timed(True)
        passes_f = res.n_passes.cpu().numpy().astype(np.int64)
        stored = eng.stored_tiles().cpu().numpy().astype(np.int64)
        resident = eng.solver_resident().cpu().numpy().astype(np.int64)
        edges = eng.edge_counts().cpu().numpy().astype(np.int64)
        cluster_fast = eng.solver_cluster
        e2e_fast, _ = e2e_run()
        eng.set_tile_skip(False)
        eng.plan(d_row, [N1] * B, [N2] * B)

    if rank == 0:
        peaks, which = measured_peaks()
        peak = float(peaks["hbm_gbs"])
        gbs = lambda nbytes, t_ms: nbytes / (t_ms * 1e-3) / 1e9
        in_bytes = B * (8.0 * d_row * (N1 + N2) + 8.0 * m)
        # algorithmic bytes, dense format (SURVEY 8d with the packed-upper layout, s = 8): K1 writes and K2 reads per
        # pass s*m(m-1)/2; K1 also reads both feature tables and writes the single scores
        tri = 8.0 * m * (m - 1) / 2
        # K2: only passes that stream the matrix count (a problem still iterating after 8 streamed passes is extracted
        # into shared memory — rb_set_solver_promote — and its remaining passes move no HBM bytes)
        k2_bytes = float(streamed.sum()) * (tri + 24.0 * m)
        k1_bytes = B * tri + in_bytes
        k1_ms = stage["prep"] + stage["affinity"]
        n_tiles = (m // 32 + (m % 32 > 0)) * (m // 32 + (m % 32 > 0) + 1) // 2
        # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` capture of this
        # kernel at this configuration (profiles/r1_solver_final5_dense_B592_summary.txt); null for other batch sizes
        traffic = 52.599338e9 + 170.937856e6 if (B == 592 and world == 1 and args.rank_offset == 0 and args.promote < 0) else None
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "pairs_per_gpu_per_step": B, "parallelism": f"round-robin pairs over {world} GPU(s)",
                       "l2": f"inputs larger than L2: packed affinity working set {eng.packed_bytes / 1e9:.2f} GB per GPU per step",
                       "layout": "packed upper triangle, 32x32 fp64 tiles, dense format (every tile written by K1 and streamed by K2 each pass)",
                       "mean_passes": float(passes.mean()), "max_passes": int(passes.max()),
                       "mean_streamed_passes": float(streamed.mean()), "promoted_problems": promoted,
                       "promote_after_streamed_passes": 8, "solver_cluster": int(cluster_dense)},
            "roofline": {"kernel": "rb_solver_kernel (K2)", "bound": "hbm", "achieved": gbs(k2_bytes, stage["solve"]),
                         "peak": peak, "unit": "GB/s", "frac": gbs(k2_bytes, stage["solve"]) / peak, "traffic": traffic,
                         "peak_source": which, "algorithmic_bytes_per_launch": k2_bytes, "ms_per_launch": stage["solve"],
                         "note": "peak is the measured COPY bandwidth (reads + writes); K2 only reads, and a read-only "
                                 "stream can exceed it - against ncu's own DRAM peak the committed capture sits at 87.9 %"},
            "kernels": {"K1_affinity": {"achieved": gbs(k1_bytes, k1_ms), "frac": gbs(k1_bytes, k1_ms) / peak, "ms_per_launch": k1_ms,
                                        "algorithmic_bytes_per_launch": k1_bytes,
                                        "launches": "rb_prep_table_kernel + rb_block_mask_kernel + rb_affinity_kernel"},
                        "K2_solver": {"achieved": gbs(k2_bytes, stage["solve"]), "frac": gbs(k2_bytes, stage["solve"]) / peak,
                                      "ms_per_launch": stage["solve"]},
                        "K3_arun_ms": stage["arun"]},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "api": "rb_align_batch_host (pinned host buffers; H2D pipelined with K1a/K1/K2/K3 over up to 5 problem ranges, then D2H)"},
            "gpu_launches": 5 * K,
            "clocks": clocks,
        }
        if do_fast:
            # fast path bytes: a problem solved from K1's edge list reads 12 B per edge once; one built from its tiles reads
            # them once; a streamed one reads them every pass; plus u0 in, scratch and outputs (~24 m) and the tile flags
            per_prob = np.where(resident == 2, edges * 12, np.where(resident == 1, stored * 8192, passes_f * stored * 8192))
            k2f = float((per_prob + 24 * m).sum()) + float(B * n_tiles)
            k1f = float(stored.sum()) * 8192 + B * n_tiles + float(np.minimum(edges, 32 * (m // 32 + (m % 32 > 0))).sum()) * 12 + in_bytes
            k1f_ms = stage_f["prep"] + stage_f["affinity"]
            line["fast_path"] = {
                "what": "same batch with rb_set_tile_skip(1), the opt-in sparse fast path (SURVEY 8f rank 2): K1 stores only tiles "
                        "holding a non-zero and lists the edges, K2 solves every problem of this workload from a shared-memory CSR "
                        "built from that list; identical ordered inliers, u and F within the oracle tolerances. These kernels move "
                        "almost no HBM bytes: they are latency / fp64 bound by construction, the HBM rooflines above do not apply",
                "value": world * B * K / (ms_f * 1e-3), "ms_per_step": ms_f / K,
                "e2e": {"value": e2e_fast, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                        "note": "PCIe-bound: the host-to-device copy of the fp64 descriptors alone takes ~6.8 ms per step"},
                "kernels": {"K1_affinity": {"ms_per_launch": k1f_ms, "bytes_per_launch": k1f, "achieved": gbs(k1f, k1f_ms)},
                            "K2_solver": {"ms_per_launch": stage_f["solve"], "bytes_per_launch": k2f, "achieved": gbs(k2f, stage_f["solve"])},
                            "K3_arun_ms": stage_f["arun"]},
                "resident_problems": int((resident > 0).sum()), "from_edge_list": int((resident == 2).sum()),
                "mean_edges": float(edges.mean()), "mean_stored_tiles": float(stored.mean()), "tiles_per_problem": int(n_tiles),
                "solver_cluster": int(cluster_fast)}
        if world == 1 and not args.no_cpu:
            cb, _ = cpu_baseline(16)
            line["cpu_baseline"] = cb
        emit(line)
    if world > 1:
        dist.destroy_process_group()


_RESULT_OUT = None


def emit(line: dict):
    """The ONE line of stdout: main() points fd 1 at stderr for the whole run (NCCL, torchrun and library banners would
    otherwise land on stdout next to the result) and keeps the real stdout for this call."""
    out = _RESULT_OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    global _RESULT_OUT
    sys.stdout.flush()
    _RESULT_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=592, help="submap pairs per GPU per step (4 x 148 SMs)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--rank-offset", type=int, default=0,
                    help="dev: shift the synthetic seeds by this many ranks (run another rank's batch on this GPU)")
    ap.add_argument("--promote", type=int, default=-1, help="dev: rb_set_solver_promote value (-1 = library default)")
    ap.add_argument("--format", default="both", choices=["both", "dense"],
                    help="both: also time the opt-in sparse fast path (fast_path block); dense: default format only")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
