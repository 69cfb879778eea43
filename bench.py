#!/usr/bin/env python
"""
bench.py — submap-pair alignments/sec on the BASELINE.json workloads.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--config metric|cfg3|cfg5] [--batch B]

One "step" = one pass of the hot path (K1a + K1 + K2 + K3) over one batch of synthetic submap pairs per GPU.

  --config metric (default)  B = 592 pairs per GPU, n1=40, n2=50 objects (m = 2000 all-to-all correspondences), method
                   'roman' (gravity + PCA + volume + 768-d semantics), SubmapAlignParams defaults, the generator of
                   SURVEY §8d — the configuration BASELINE.json's metric is quoted on.
  --config cfg3    B = 256 pairs per GPU, n1 = n2 = 32 (m = 1024) — BASELINE.json configs[2].
  --config cfg5    the all-pairs loop-closure search (BASELINE.json configs[4]; reference loop
                   roman/align/submap_align.py:93-200): per GPU one robot pair = 20 x 20 submaps of <= 40 objects, every
                   submap pair one problem (400 per GPU, ragged m <= 1600), object tables shared between problems
                   (rb_plan_indexed).  With N GPUs every GPU searches its own robot pair (a multi-robot session has one
                   such search per pair of robots); a rank's problems are ordered longest-first, which is also the
                   order the persistent solver queue takes them in.
All configs are weak scaling (fixed work per GPU).  The only inter-GPU exchange is one all-gather of fixed-size per-pair
records per step.

Printed JSON line: see the task contract.  Everything at the top level is measured in the library's default DENSE
format (every tile of the packed affinity matrix written by K1 and streamed by K2 each pass): `roofline` is for the
dominant kernel (K2), `kernels` lists K1 and K2 separately; `parity_checked` pairs of the timed batch are compared
with the CPU oracle AFTER the timing (ordered inliers identical, SE(3) within 1e-4).  N = 1 adds `fast_path` (same
batch, opt-in sparse format), `config.variants.cone` (metric config with cone-shaped descriptors: realistic graph
density) and `cpu_baseline`.

--impl reference times the reference algorithm on the host cores: oracle/_ref (the reference's own C++ sources compiled
against the committed Eigen stand-in) when it was built, else the C restatement oracle/clipper_oracle.c.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SEM_DIM, METHOD = 768, "roman"
D_ROW = 3 + 3 + 1 + SEM_DIM
CFG5_POOL = 4   # robot pairs a cfg5 run rotates through (see run_b200)
GATHER_RING = 4 # N > 1: all-gathers of per-pair records that may be in flight behind the next steps' kernels
METRIC = "submap-pair alignments/sec at N=2000 correspondences"
UNIT = "alignments/s"
T_TOL = 1e-4


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


# --------------------------------------------------------------------------------------------------------- workloads
class Workload:
    """One rank's batch: shapes, host inputs (flat float64), and per-problem views for the CPU side."""

    def __init__(self, name, text, n1, n2, D1, D2, u0, obj1=None, obj2=None, meta=None):
        self.name, self.text = name, text
        self.n1, self.n2 = np.asarray(n1, dtype=np.int32), np.asarray(n2, dtype=np.int32)
        self.B = len(self.n1)
        self.m = self.n1.astype(np.int64) * self.n2.astype(np.int64)
        self.D1, self.D2, self.u0 = D1, D2, u0          # (rows, d) tables / flat start vectors
        self.obj1, self.obj2 = obj1, obj2               # row offsets into shared tables, or None (private rows)
        if obj1 is None:
            self.obj1 = np.concatenate([[0], np.cumsum(self.n1)[:-1]]).astype(np.int64)
            self.obj2 = np.concatenate([[0], np.cumsum(self.n2)[:-1]]).astype(np.int64)
            self.indexed = False
        else:
            self.indexed = True
        self.aoff = np.concatenate([[0], np.cumsum(self.m)]).astype(np.int64)
        self.meta = meta or {}

    def plan(self, eng):
        if self.indexed:
            eng.plan(D_ROW, self.n1, self.n2, obj1_offsets=self.obj1, obj2_offsets=self.obj2,
                     total_obj1=len(self.D1), total_obj2=len(self.D2))
        else:
            eng.plan(D_ROW, self.n1, self.n2)

    def pair(self, b):
        """(D1 (n1,d), D2 (n2,d), u0 (m,)) of problem b."""
        return (self.D1[self.obj1[b]:self.obj1[b] + self.n1[b]], self.D2[self.obj2[b]:self.obj2[b] + self.n2[b]],
                self.u0[self.aoff[b]:self.aoff[b + 1]])

    @property
    def table_bytes(self):
        return 8.0 * D_ROW * (len(self.D1) + len(self.D2))


def build_workload(cfg, rank, world, batch=0, desc_common=None, rank_offset=0):
    from roman_b200 import synth
    if cfg in ("metric", "cfg3"):
        n1, n2 = (40, 50) if cfg == "metric" else (32, 32)
        B = batch or (592 if cfg == "metric" else 256)
        dc = 0.0 if desc_common is None else desc_common
        seed0 = (1000 if cfg == "metric" else 2000) + (rank + rank_offset) * B
        D1, D2, u0, _ = synth.make_batch(seed0, B, n1, n2, method=METHOD, semantics_dim=SEM_DIM, desc_common=dc)
        text = (f"synthetic submap pairs n1={n1} n2={n2} (m={n1 * n2} all-to-all), method roman, D=768, "
                "SubmapAlignParams defaults")
        if dc != 0.0:
            text += f", descriptors in a cone (desc_common={dc})"
        return Workload(cfg, text, [n1] * B, [n2] * B, np.ascontiguousarray(D1.reshape(B * n1, -1)),
                        np.ascontiguousarray(D2.reshape(B * n2, -1)), np.ascontiguousarray(u0.reshape(-1)))
    assert cfg == "cfg5"
    S = 20
    dc = 1.0 if desc_common is None else desc_common
    # one robot pair (world) per GPU: a rank uploads and searches its own pair's tables only.  (Dealing the N x 400
    # problems of all pairs longest-first over the ranks balances K2 to +-2.5 %, but then every rank needs every pair's
    # tables: 8 x 12 MB per step through the shared host PCIe cost more than the balance won — measured.)
    worlds = [synth.make_all_pairs_world(5000 + rank + rank_offset, S, S, 40, method=METHOD, semantics_dim=SEM_DIM, desc_common=dc)]
    t0 = [t for w in worlds for t in w.tables[0]]
    t1 = [t for w in worlds for t in w.tables[1]]
    off0 = np.concatenate([[0], np.cumsum([len(t) for t in t0])]).astype(np.int64)
    off1 = np.concatenate([[0], np.cumsum([len(t) for t in t1])]).astype(np.int64)
    probs = [(0, i, j) for i in range(S) for j in range(S)]
    msz = np.array([len(t0[i]) * len(t1[j]) for (_, i, j) in probs])
    order = np.argsort(-msz, kind="stable")                      # longest problem first: the order the solver queue takes them in
    sel = [probs[int(g)] for g in order]
    if batch:
        sel = sel[:batch]
    a = np.array([w * S + i for (w, i, j) in sel])
    b = np.array([w * S + j for (w, i, j) in sel])
    n1 = np.array([len(t0[x]) for x in a], dtype=np.int32)
    n2 = np.array([len(t1[x]) for x in b], dtype=np.int32)
    u0 = np.concatenate([np.random.default_rng(7_000_000 + 1000 * (rank + rank_offset) + 20 * i + j).uniform(0, 1, int(x) * int(y))
                         for (w, i, j), x, y in zip(sel, n1, n2)])
    text = (f"all-pairs loop-closure search: per GPU one robot pair of {S} x {S} synthetic submaps (<= 40 objects, ragged "
            f"m <= 1600), every submap pair one problem, shared object tables, method roman, D=768, descriptors in a cone "
            f"(desc_common={dc}), SubmapAlignParams defaults")
    return Workload("cfg5", text, n1, n2, np.ascontiguousarray(np.concatenate(t0)), np.ascontiguousarray(np.concatenate(t1)),
                    np.ascontiguousarray(u0), obj1=off0[a], obj2=off1[b], meta={"pairs": sel, "submaps": S})


def oracle_params():
    from oracle import clipper_np as cn
    from roman_b200.engine import method_invariant_kwargs
    kw = method_invariant_kwargs(METHOD, semantics_dim=SEM_DIM)
    return cn.InvariantParams(**kw), cn.SolverParams()


def host_threads():
    """Physical cores this process may use (SMT siblings share the FP units and the memory pipes; measured on the
    2 x 32-core host: 64 threads beat 128)."""
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


class CpuArm:
    """The reference algorithm on the host: oracle/_ref (the reference's own sources) if built, else the C restatement."""

    def __init__(self):
        import oracle
        self.ip, self.sp = oracle_params()
        self.ref = None
        try:
            from oracle import ref as oref
            if oref.available():
                self.ref = oref.RefClipper()
        except Exception:
            self.ref = None
        self.port = oracle.COracle()
        self.kind = "reference" if self.ref is not None else "port"

    def align(self, D1, D2, u0, parallelize=False):
        if self.ref is not None:
            return self.ref.align_pair(self.ip, self.sp, D1, D2, u0, parallelize=parallelize)
        return self.port.align_pair(self.ip, self.sp, D1, D2, u0, parallelize=parallelize)

    def throughput(self, wl, idx, threads):
        """One pair per host thread (throughput-fair mode of BASELINE.md §3)."""
        with ThreadPoolExecutor(max_workers=threads) as ex:
            list(ex.map(lambda b: self.align(*wl.pair(b)), idx[:threads]))          # warm the pool / the allocator
            t0 = time.perf_counter()
            res = list(ex.map(lambda b: self.align(*wl.pair(b)), idx))
            dt = time.perf_counter() - t0
        return len(idx) / dt, dt, res

    def latency(self, wl, idx):
        """One pair at a time, OpenMP inside the affinity build, solver on one core (clipper.cpp:39,95 vs :243) —
        what a user of clipperpy experiences per pair."""
        self.align(*wl.pair(idx[0]), parallelize=True)
        t0 = time.perf_counter()
        for b in idx:
            self.align(*wl.pair(b), parallelize=True)
        dt = time.perf_counter() - t0
        return len(idx) / dt, dt


def cpu_baseline(wl, budget_s=12.0, threads=0, latency=True):
    """Bounded sample of the SAME workload on the host cores: about `budget_s` seconds of wall per mode."""
    arm = CpuArm()
    if threads <= 0:
        threads = host_threads()
    # calibrate on a few pairs, then size the sample (at least two rounds over the threads, at most the batch x 4)
    probe = list(range(min(wl.B, max(2, min(threads, 8)))))
    t0 = time.perf_counter()
    for b in probe[:2]:
        arm.align(*wl.pair(b))
    per = (time.perf_counter() - t0) / 2
    n = int(min(max(2 * threads, budget_s / max(per, 1e-4) * threads * 0.5), 4 * wl.B))
    idx = [i % wl.B for i in range(n)]
    v, dt, _ = arm.throughput(wl, idx, threads)
    out = {"value": v, "unit": UNIT, "cores": threads, "kind": arm.kind,
           "sample": f"{n} alignments of the same workload (problems 0..{min(n, wl.B) - 1} of the batch, cycled), one pair per "
                     f"host thread, {dt:.2f} s wall", "host_cpus": os.cpu_count(), "pairs": n, "seconds": dt,
           "mode": "throughput-fair (one problem per core)"}
    if latency:
        nl = int(min(max(8, budget_s / max(per, 1e-4) * 0.5), wl.B))
        os.environ.pop("OMP_NUM_THREADS", None)
        lv, ldt = arm.latency(wl, list(range(nl)))
        out["latency_mode"] = {"value": lv, "unit": UNIT, "sample": f"{nl} alignments, one at a time, OpenMP over the pair index "
                               f"inside the affinity build on all host threads, solver on one core, {ldt:.2f} s wall",
                               "ms_per_alignment": 1e3 * ldt / nl}
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = build_workload(args.config, 0, 1, args.batch)
    arm = CpuArm()
    threads = host_threads()
    # one step = a bounded sample: two rounds over the host threads (cycled through the batch)
    n = max(2 * threads, 16)
    times = []
    for s in range(args.warmup + args.steps):
        idx = [(s * n + i) % wl.B for i in range(n)]
        v, dt, _ = arm.throughput(wl, idx, threads)
        if s >= args.warmup:
            times.append(dt)
    t = float(np.mean(times))
    v = n / t
    cb = {"value": v, "unit": UNIT, "cores": threads, "kind": arm.kind,
          "sample": f"{n} alignments per step of the same workload, one pair per host thread", "host_cpus": os.cpu_count(),
          "pairs": n, "seconds": t, "mode": "throughput-fair (one problem per core)"}
    line = {"impl": "reference", "metric": METRIC if wl.name == "metric" else f"submap-pair alignments/sec ({wl.name})", "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl.text, "name": wl.name, "pairs_per_gpu_per_step": wl.B,   # same keys / values as the GPU arm's
                       "sample_pairs_per_step": n,                                               # what one CPU step processes of it
                       "note": ("the reference's own C++ sources (dependencies/clipper: clipper.cpp, utils.cpp, invariants/roman.cpp) "
                                "compiled against the committed Eigen stand-in (oracle/_ref), plus the numpy Arun of T_align restated in C"
                                if arm.kind == "reference" else
                                "CPU restatement of the reference (oracle/clipper_oracle.c, kind=port)") + "; one pair per host thread"},
            "cpu_baseline": cb,
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


# --------------------------------------------------------------------------------------------------------- GPU arm
def check_parity(wl, host_out, idx):
    """Compare problems `idx` of a finished batch (host result dict of align_batch_host) with the CPU oracle:
    ordered inlier associations identical, status equal, SE(3) within 1e-4 relative Frobenius."""
    arm = CpuArm()
    bad = []
    for b in idx:
        ref = arm.align(*wl.pair(b))
        o, k = int(wl.aoff[b]), int(host_out["n_sel"][b])
        got = host_out["sel"][o:o + k]
        ok = np.array_equal(got, ref["sel"]) and int(host_out["status"][b]) == int(ref["status"])
        if ok and ref["status"] == 0:
            ok = np.linalg.norm(host_out["T"][b] - ref["T"]) / np.linalg.norm(ref["T"]) <= T_TOL
        if not ok:
            bad.append(int(b))
    return bad, arm.kind


def traffic_for(key):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed ncu capture of
    this exact configuration (profiles/traffic.json, written by tools/ncu_traffic.py from the .ncu-rep); None otherwise."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        with open(p) as f:
            t = json.load(f).get(key)
    except (OSError, ValueError):
        return None, f"no committed ncu capture for {key}"
    if not t:
        return None, f"no committed ncu capture for {key}"
    return float(t["dram_bytes_read"]) + float(t["dram_bytes_write"]), t.get("source", "")


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
    K, W = args.steps, args.warmup

    def make_engine():
        e = Engine(local)
        e.set_invariant_params(make_invariant_params(**method_invariant_kwargs(METHOD, semantics_dim=SEM_DIM)))
        if args.promote >= 0:
            e.set_solver_promote(args.promote)
        if args.cluster > 0:
            e.set_solver_cluster(args.cluster)
        elif args.config == "cfg5":   # the preset roman_b200.align.align_submap_pairs applies to an all-pairs search
            from roman_b200.align.submap_align import ALL_PAIRS_SOLVER_CLUSTER
            e.set_solver_cluster(ALL_PAIRS_SOLVER_CLUSTER)
        e.set_profiling(True)
        return e

    eng = make_engine()
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a).reshape(-1)).pin_memory()

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    class Bound:
        """A workload planned on an engine with its device / pinned buffers."""

        def __init__(self, wl, engine=None):
            self.wl = wl
            self.eng = engine if engine is not None else eng
            self.D1p, self.D2p, self.u0p = pin(wl.D1), pin(wl.D2), pin(wl.u0)
            self.D1d, self.D2d, self.u0d = self.D1p.to(dev), self.D2p.to(dev), self.u0p.to(dev)
            B = wl.B
            # N > 1: per-pair records (status, n_inliers, F, pad, T[16]) and their all-gather, a ring of GATHER_RING in flight
            self.rec = [torch.zeros((B, 20), dtype=torch.float64, device=dev) for _ in range(GATHER_RING)] if world > 1 else None
            self.gathered = [torch.zeros((world * B, 20), dtype=torch.float64, device=dev) for _ in range(GATHER_RING)] if world > 1 else None
            self.work = [None] * GATHER_RING
            self.nstep = 0
            self.res = None
            self.out_h = None

        def plan(self, tile_skip):
            self.eng.set_tile_skip(tile_skip)
            self.wl.plan(self.eng)     # the format decides the solver's cluster size; planning is untimed
            self.res = self.eng.alloc_result()

        def step(self):
            res = self.res
            self.eng.align_batch(self.D1d, self.D2d, self.u0d, out=res)
            if world > 1:
                # The path's only exchange: all-gather of fixed-size per-pair records.  It is ASYNCHRONOUS: NCCL runs it on
                # its own stream behind this step's kernels, and the launch stream goes straight on with the next step; a
                # slot of the ring is waited for only when it comes round again (and all of them before the timed region
                # ends).  A blocking all-gather makes every step as long as the slowest rank's — with data-dependent work
                # per search that cost the all-pairs config 15 % at N = 8.
                k = self.nstep % GATHER_RING
                self.nstep += 1
                if self.work[k] is not None:
                    self.work[k].wait()
                rec = self.rec[k]
                rec[:, 0] = res.status.to(torch.float64)
                rec[:, 1] = res.n_sel.to(torch.float64)
                rec[:, 2] = res.score
                rec[:, 4:] = res.T.reshape(-1, 16)
                self.work[k] = dist.all_gather_into_tensor(self.gathered[k], rec, async_op=True)

        def drain(self):
            """The launch stream waits for every all-gather still in flight."""
            for k, w in enumerate(self.work):
                if w is not None:
                    w.wait()
                    self.work[k] = None

        def host_step(self):
            self.out_h = self.eng.align_batch_host(self.D1p, self.D2p, self.u0p, out=self.out_h)

        def timed(self, tile_skip, K, W, sampler=None):
            ms, stage, _ = timed_pool([self], tile_skip, K, W, sampler)
            return ms, stage

        def e2e(self, K, W):
            v, last, h2d, d2h = e2e_pool([self], K, W)
            return v, last.out_h, h2d, d2h

    def timed_pool(pool, tile_skip, K, W, sampler=None):
        """W warm-ups (of every member), then exactly K steps between device events; max over ranks.  Step s of rank r
        runs member (r + s) mod len(pool): with one member that is the same batch every step, with several (cfg5) every
        rank works through the same rotation of workloads."""
        for b in pool:
            b.plan(tile_skip)
            for _ in range(W):
                b.step()
            b.drain()
        sync()
        if sampler is not None:
            sampler.mark()  # only samples taken from here on (the timed region) are reported
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        stage = {"prep": 0.0, "affinity": 0.0, "solve": 0.0, "arun": 0.0}
        counts = [0] * len(pool)
        sync()
        ev0.record()
        for s_ in range(K):
            k = (rank + s_) % len(pool)
            b = pool[k]
            b.step()
            counts[k] += 1
            for kk, v in b.eng.stage_ms().items():  # CUDA events on the launch stream (syncs on the step's last event)
                stage[kk] += v / K
        for b in pool:
            b.drain()      # the timed region ends when the last all-gather has landed
        ev1.record()
        sync()
        t = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
        smax = torch.tensor([stage["solve"]], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            all_solve = [torch.zeros_like(smax) for _ in range(world)]
            dist.all_gather(all_solve, smax)
            stage["solve_per_rank"] = [float(x.item()) for x in all_solve]
        return float(t.item()), stage, counts

    def e2e_pool(pool, K, W):
        """K calls of the host-buffer C ABI entry point: pinned host inputs, H2D + kernels + D2H inside the timed region."""
        sync()
        for b in pool:
            for _ in range(max(1, min(W, 3))):
                b.host_step()
        sync()
        last = pool[0]
        t0 = time.perf_counter()
        for s_ in range(K):
            last = pool[(rank + s_) % len(pool)]
            last.host_step()
        wall = time.perf_counter() - t0
        te = torch.tensor([wall], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        h2d, d2h = last.eng.host_copy_bytes()
        return world * last.wl.B * K / float(te.item()), last, h2d, d2h

    # cfg5: a stream of robot-pair searches dealt round-robin over the GPUs — a pool of CFG5_POOL robot pairs (one engine
    # each, planned up front), step s of rank r searches pair (r + s) mod pool.  Every rank, at every N, works through
    # the same rotation, so the N-GPU efficiency measures the machine and not which robot pair a rank happened to get
    # (the work of a search depends on its data: K2 of one 400-problem search ranges from 5.9 to 8.9 ms over seeds).
    if args.config == "cfg5":
        pool = [Bound(build_workload("cfg5", k, world, args.batch, args.desc_common, args.rank_offset), eng if k == 0 else make_engine())
                for k in range(CFG5_POOL)]
    else:
        pool = [Bound(build_workload(args.config, rank, world, args.batch, args.desc_common, args.rank_offset))]
    wl = pool[0].wl
    B = wl.B
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        sampler.wait_ready()

    def cat(counts, f):
        """Per-problem arrays of the K timed steps, concatenated (member k appears counts[k] times)."""
        return np.concatenate([np.tile(np.asarray(f(b)), c) for b, c in zip(pool, counts) if c])

    # ---- main measurement: the library's default (dense) format — every tile of the packed upper triangle is written
    #      by K1 and streamed by K2 each pass: the kernels and byte counts of BASELINE.json's north_star / SURVEY 8d
    ms, stage, counts = timed_pool(pool, False, K, W, sampler if rank == 0 else None)
    clocks = sampler.stop() if rank == 0 else None
    value = world * B * K / (ms * 1e-3)
    m = cat(counts, lambda b: b.wl.m)
    passes = cat(counts, lambda b: b.res.n_passes.cpu().numpy().astype(np.int64))
    streamed = cat(counts, lambda b: b.eng.solver_streamed_passes().cpu().numpy().astype(np.int64)[:B])   # passes that read M from HBM
    promoted = int(round(float((cat(counts, lambda b: b.eng.solver_resident().cpu().numpy()[:B]) >= 3).sum()) / K))
    live = cat(counts, lambda b: b.eng.live_counts().cpu().numpy().astype(np.int64)[:B])
    cluster_dense = eng.solver_cluster
    packed_bytes = eng.packed_bytes
    e2e_value, last, h2d, d2h = e2e_pool(pool, K, W)
    out_h, wl_chk = last.out_h, last.wl

    # ---- same batch through the opt-in sparse fast path (rb_set_tile_skip): only tiles holding a non-zero are stored,
    #      K1 lists the edges, K2 solves from a shared-memory CSR.  N = 1 only: the N-GPU sweep stays short.
    do_fast = (args.format == "both") and world == 1
    if do_fast:
        ms_f, stage_f, counts_f = timed_pool(pool, True, K, W)
        passes_f = cat(counts_f, lambda b: b.res.n_passes.cpu().numpy().astype(np.int64))
        stored = cat(counts_f, lambda b: b.eng.stored_tiles().cpu().numpy().astype(np.int64)[:B])
        resident = cat(counts_f, lambda b: b.eng.solver_resident().cpu().numpy().astype(np.int64)[:B])
        edges = cat(counts_f, lambda b: b.eng.edge_counts().cpu().numpy().astype(np.int64)[:B])
        m_f = cat(counts_f, lambda b: b.wl.m)
        cluster_fast = eng.solver_cluster
        e2e_fast, _, _, _ = e2e_pool(pool, K, W)
        for b in pool:
            b.eng.set_tile_skip(False)

    if rank == 0:
        peaks, which = measured_peaks()
        peak = float(peaks["hbm_gbs"])
        gbs = lambda nbytes, t_ms: nbytes / (t_ms * 1e-3) / 1e9
        # algorithmic bytes PER LAUNCH (the arrays hold all K timed steps), dense format (SURVEY 8d with the packed-upper
        # layout, s = 8): K1 writes and K2 reads per pass s*m(m-1)/2; K1 also reads the feature tables (shared tables
        # once) and writes the single scores.
        tri = 8.0 * m * (m - 1) / 2
        tbl = lambda w: (w.table_bytes if w.indexed else float((8.0 * D_ROW * (w.n1 + w.n2)).sum())) + 8.0 * float(w.m.sum())
        in_bytes = sum(c * tbl(b.wl) for b, c in zip(pool, counts)) / K
        # K2: only passes that stream the matrix count (a problem still iterating after 8 streamed passes is extracted
        # into shared memory — rb_set_solver_promote — and its remaining passes move no HBM bytes)
        k2_bytes = float((streamed * (tri + 24.0 * m)).sum()) / K
        k1_bytes = float(tri.sum()) / K + in_bytes
        k1_ms = stage["prep"] + stage["affinity"]
        nbt = (m + 31) // 32
        n_tiles = nbt * (nbt + 1) // 2
        promote = 8 if args.promote < 0 else args.promote
        tkey = f"{wl.name}:B{B}:desc{args.desc_common if args.desc_common is not None else 'default'}:promote{promote}:dense:K2"
        traffic, tsrc = traffic_for(tkey) if (world == 1 and args.rank_offset == 0) else (None, "N > 1 / shifted batch: not captured")
        # parity of the timed batch, after the timing: the host result of the last end-to-end call against the oracle
        nchk = B if wl.name in ("cfg3", "cfg5") and world == 1 else min(B, 64)
        idx = list(range(B)) if nchk == B else [int(x) for x in np.linspace(0, B - 1, nchk).astype(int)]
        if args.no_parity:
            bad, okind, idx = [], None, []
        else:
            bad, okind = check_parity(wl_chk, out_h, idx)
            if bad:
                print(f"PARITY MISMATCH on problems {bad[:20]} of the timed batch", file=sys.stderr)
        line = {
            "metric": METRIC if wl.name == "metric" else f"submap-pair alignments/sec ({wl.name})", "value": value, "unit": UNIT,
            "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl.text, "name": wl.name, "pairs_per_gpu_per_step": B,
                       "parallelism": (f"one robot pair (20 x 20 submaps = {B} problems, longest first) per GPU per step, {world} GPU(s); "
                                       f"step s of rank r searches pair (r + s) mod {CFG5_POOL} of a pool of {CFG5_POOL}" if wl.name == "cfg5"
                                       else f"round-robin pairs over {world} GPU(s)"),
                       "l2": f"inputs larger than L2: packed affinity working set {packed_bytes / 1e9:.2f} GB per GPU per step",
                       "layout": "packed upper triangle, 32x32 fp64 tiles, dense format (every tile written by K1 and streamed by K2 each pass)",
                       "m_min": int(m.min()), "m_max": int(m.max()), "m_mean": float(m.mean()),
                       "mean_passes": float(passes.mean()), "max_passes": int(passes.max()),
                       "mean_streamed_passes": float(streamed.mean()), "promoted_problems": promoted,
                       "promote_after_streamed_passes": promote, "solver_cluster": int(cluster_dense),
                       "mean_live_associations": float(live.mean())},
            "roofline": {"kernel": "rb_solver_kernel (K2)", "bound": "hbm", "achieved": gbs(k2_bytes, stage["solve"]),
                         "peak": peak, "unit": "GB/s", "frac": gbs(k2_bytes, stage["solve"]) / peak, "traffic": traffic,
                         "traffic_source": tsrc,
                         "peak_source": which, "algorithmic_bytes_per_launch": k2_bytes, "ms_per_launch": stage["solve"],
                         "note": "peak is the measured COPY bandwidth (reads + writes); K2 only reads, and a read-only "
                                 "stream can exceed it"},
            "kernels": {"K1_affinity": {"achieved": gbs(k1_bytes, k1_ms), "frac": gbs(k1_bytes, k1_ms) / peak, "ms_per_launch": k1_ms,
                                        "K1a_prep_ms": stage["prep"], "K1_affinity_ms": stage["affinity"],
                                        "algorithmic_bytes_per_launch": k1_bytes,
                                        "launches": "rb_prep_table_kernel + rb_block_mask_kernel + rb_affinity_kernel"},
                        "K2_solver": {"achieved": gbs(k2_bytes, stage["solve"]), "frac": gbs(k2_bytes, stage["solve"]) / peak,
                                      "ms_per_launch": stage["solve"]},
                        "K3_arun_ms": stage["arun"]},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "pcie_h2d_gbs": h2d / (B / (e2e_value / world)) / 1e9,
                    "api": "rb_align_batch_host (pinned host buffers; H2D pipelined with K1a/K1/K2/K3 over problem ranges, "
                           "compact D2H of the inlier rows)"},
            "gpu_launches": 5 * K,
            "parity_checked": len(idx), "parity_mismatches": len(bad), "parity_oracle": okind,
            "clocks": clocks,
        }
        if world > 1:
            line["config"]["K2_ms_per_rank"] = stage.get("solve_per_rank")
        if do_fast:
            # fast path bytes: a problem solved from K1's edge list reads 12 B per edge once; one built from its tiles reads
            # them once; a streamed one reads them every pass; plus u0 in, scratch and outputs (~24 m) and the tile flags
            nbt_f = (m_f + 31) // 32
            n_tiles_f = nbt_f * (nbt_f + 1) // 2
            per_prob = np.where(resident == 2, edges * 12, np.where(resident == 1, stored * 8192, passes_f * stored * 8192))
            k2f = (float((per_prob + 24 * m_f).sum()) + float(n_tiles_f.sum())) / K
            k1f = (float(stored.sum()) * 8192 + float(n_tiles_f.sum()) + float(np.minimum(edges, 32 * nbt_f).sum()) * 12) / K + in_bytes
            k1f_ms = stage_f["prep"] + stage_f["affinity"]
            line["fast_path"] = {
                "what": "same batch with rb_set_tile_skip(1), the opt-in sparse fast path (SURVEY 8f rank 2): K1 stores only tiles "
                        "holding a non-zero and lists the edges, K2 solves problems whose edges fit from a shared-memory CSR built "
                        "from that list; identical ordered inliers, u and F within the oracle tolerances. These kernels move "
                        "few HBM bytes: they are latency / fp64 bound by construction, the HBM rooflines above do not apply",
                "value": world * B * K / (ms_f * 1e-3), "ms_per_step": ms_f / K,
                "e2e": {"value": e2e_fast, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
                "kernels": {"K1_affinity": {"ms_per_launch": k1f_ms, "K1a_prep_ms": stage_f["prep"], "bytes_per_launch": k1f, "achieved": gbs(k1f, k1f_ms)},
                            "K2_solver": {"ms_per_launch": stage_f["solve"], "bytes_per_launch": k2f, "achieved": gbs(k2f, stage_f["solve"])},
                            "K3_arun_ms": stage_f["arun"]},
                "resident_problems": int(round((resident > 0).sum() / K)), "from_edge_list": int(round((resident == 2).sum() / K)),
                "streamed_problems": int(round((resident == 0).sum() / K)),
                "mean_edges": float(edges.mean()), "max_edges": int(edges.max()), "mean_stored_tiles": float(stored.mean()),
                "mean_tiles_per_problem": float(n_tiles_f.mean()), "solver_cluster": int(cluster_fast)}
        if world == 1 and wl.name == "metric" and args.desc_common is None and not args.no_variants:
            line["config"]["variants"] = {"cone": cone_variant(Bound, eng, K, W, peak)}
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu_baseline(pool[0].wl)
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def cone_variant(Bound, eng, K, W, peak):
    """The metric config with cone-shaped descriptors (desc_common = 1: unrelated objects have cosines well above
    cosine_min, as CLIP / DINO embeddings do), i.e. a realistic graph density: K1 becomes transcendental-bound,
    K2 runs tens of passes.  Device-resident timing of both formats; no change to the headline."""
    wl = build_workload("metric", 0, 1, 0, 1.0)
    bw = Bound(wl)
    Kc = max(3, K // 3)
    ms, st = bw.timed(False, Kc, max(2, W))
    passes = bw.res.n_passes.cpu().numpy().astype(np.int64)
    streamed = eng.solver_streamed_passes().cpu().numpy().astype(np.int64)
    live = eng.live_counts().cpu().numpy().astype(np.int64)
    m = wl.m
    tri = 8.0 * m * (m - 1) / 2
    k2_bytes = float((streamed * (tri + 24.0 * m)).sum())
    ms_f, st_f = bw.timed(True, Kc, max(2, W))
    resident = eng.solver_resident().cpu().numpy().astype(np.int64)[:wl.B]
    edges = eng.edge_counts().cpu().numpy().astype(np.int64)[:wl.B]
    stored = eng.stored_tiles().cpu().numpy().astype(np.int64)[:wl.B]
    eng.set_tile_skip(False)
    sm = eng.sm_count
    fp64_peak = sm * 64 * 2 * 1.965e9 / 1e12           # 64 fp64 FMA lanes per SM per clock at clocks.max.sm
    pairs_eval = float((live * (live - 1) / 2).sum())
    flops = 60.0 * pairs_eval                           # ~60 fp64-equivalent flops per evaluated pair
    k1_ms = st["prep"] + st["affinity"]
    gbs = lambda nbytes, t_ms: nbytes / (t_ms * 1e-3) / 1e9
    return {
        "what": "metric config, desc_common = 1.0 (descriptors in a cone): realistic density; device-resident inputs",
        "steps": Kc, "value": wl.B * Kc / (ms * 1e-3), "ms_per_step": ms / Kc,
        "mean_live_associations": float(live.mean()), "mean_edges": float(edges.mean()), "max_edges": int(edges.max()),
        "density_pct": float(100.0 * (edges / (m * (m - 1) / 2)).mean()),
        "mean_passes": float(passes.mean()), "max_passes": int(passes.max()),
        "K1": {"ms_per_launch": k1_ms, "K1a_prep_ms": st["prep"], "evaluated_pairs_per_launch": pairs_eval,
               "fp64_tflops": flops / (k1_ms * 1e-3) / 1e12, "fp64_peak_tflops": fp64_peak,
               "frac_of_fp64_peak": flops / (k1_ms * 1e-3) / 1e12 / fp64_peak,
               "hbm_gbs": gbs(float(tri.sum()), k1_ms), "frac_of_hbm_peak": gbs(float(tri.sum()), k1_ms) / peak},
        "K2": {"ms_per_launch": st["solve"], "achieved_gbs": gbs(k2_bytes, st["solve"]), "frac": gbs(k2_bytes, st["solve"]) / peak,
               "mean_streamed_passes": float(streamed.mean())},
        "fast_path": {"value": wl.B * Kc / (ms_f * 1e-3), "ms_per_step": ms_f / Kc, "K1_ms": st_f["prep"] + st_f["affinity"],
                      "K2_ms": st_f["solve"], "resident_from_edge_list": int((resident == 2).sum()),
                      "resident_from_tiles": int((resident == 1).sum()), "streamed": int((resident == 0).sum()),
                      "mean_stored_tiles": float(stored.mean())},
    }


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
    ap.add_argument("--config", default="metric", choices=["metric", "cfg3", "cfg5"])
    ap.add_argument("--batch", type=int, default=0, help="submap pairs per GPU per step (default: 592 = 4 x 148 SMs for "
                                                         "metric, 256 for cfg3, all 400 for cfg5)")
    ap.add_argument("--desc-common", type=float, default=None, help="descriptor cone strength of the generator (default: 0 "
                                                                    "for metric / cfg3 = SURVEY 8d verbatim, 1 for cfg5)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-variants", action="store_true", help="skip config.variants (the cone-descriptor leg)")
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle comparison of the timed batch")
    ap.add_argument("--rank-offset", type=int, default=0,
                    help="dev: shift the synthetic seeds by this many ranks (run another rank's batch on this GPU)")
    ap.add_argument("--promote", type=int, default=-1, help="dev: rb_set_solver_promote value (-1 = library default)")
    ap.add_argument("--cluster", type=int, default=0, help="dev: CTAs per problem in the solver (0 = library default)")
    ap.add_argument("--format", default="both", choices=["both", "dense"],
                    help="both: also time the opt-in sparse fast path (fast_path block, N = 1); dense: default format only")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
