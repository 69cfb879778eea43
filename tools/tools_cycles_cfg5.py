"""Dev tool: per-CTA cycle breakdown of the solver on the cfg5 workload (tile-skip format: everything resident)."""
import sys, ctypes as C, numpy as np, torch
sys.path.insert(0, '.')
import bench
from roman_b200.engine import Engine, make_invariant_params, method_invariant_kwargs
eng = Engine(0)
eng._lib.rb_debug_solver_cycles.restype = C.c_int
eng._lib.rb_debug_solver_cycles(eng._h, None, 0)
eng.set_invariant_params(make_invariant_params(**method_invariant_kwargs("roman", semantics_dim=768)))
eng.set_profiling(True)
wl = bench.build_workload("cfg5", 0, 1)
D1 = torch.from_numpy(wl.D1.reshape(-1)).cuda(); D2 = torch.from_numpy(wl.D2.reshape(-1)).cuda(); u0 = torch.from_numpy(wl.u0).cuda()
for CL, TS in ((1, True), (1, False), (2, False)):
    eng.set_solver_cluster(CL)
    eng.set_tile_skip(TS)
    wl.plan(eng)
    for it in range(3):
        res = eng.align_batch(D1, D2, u0)
    ms = eng.stage_ms()
    n = min(eng._lib.rb_get_solver_max_clusters(eng._h) * CL, wl.B * CL)
    buf = np.zeros((n, 8), dtype=np.int64)
    eng._lib.rb_debug_solver_cycles(eng._h, C.c_void_p(buf.ctypes.data), n)
    f = buf.astype(float) / 1.965e3  # us at 1965 MHz
    p = res.n_passes.cpu().numpy()
    print(f"skip={TS} C={CL} ctas={n} solve={ms['solve']:.3f} ms | per-CTA us: total mean {f[:,0].mean():.0f} max {f[:,0].max():.0f} | matvec {f[:,1].mean():.0f} (max {f[:,1].max():.0f}) "
          f"combine {f[:,2].mean():.0f} queue {f[:,3].mean():.0f} (max {f[:,3].max():.0f}) round {f[::CL,4].mean():.0f} build {f[:,5].mean():.0f} | passes total {p.sum()} max {p.max()}; "
          f"matvec us per pass {f[:,1].sum() / p.sum() / CL:.2f}; total us per pass {(f[:,0]-f[:,3]).sum() / p.sum() / CL:.2f}")
