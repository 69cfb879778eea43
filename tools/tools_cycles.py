"""Dev tool: per-CTA cycle breakdown of the solver."""
import sys, ctypes as C, numpy as np, torch
sys.path.insert(0, '.')
import bench
from roman_b200.engine import Engine, make_invariant_params, method_invariant_kwargs
B = int(sys.argv[1])
eng = Engine(0)
eng._lib.rb_debug_solver_cycles.restype = C.c_int
eng._lib.rb_debug_solver_cycles(eng._h, None, 0)
eng.set_invariant_params(make_invariant_params(**method_invariant_kwargs("roman", semantics_dim=768)))
eng.set_profiling(True)
D1h, D2h, u0h = bench.make_inputs(1000, B)
D1 = torch.from_numpy(D1h.reshape(-1)).cuda(); D2 = torch.from_numpy(D2h.reshape(-1)).cuda(); u0 = torch.from_numpy(u0h.reshape(-1)).cuda()
import itertools
for CL, TS in itertools.product([1, 2, 4], [True, False]):
    eng.set_solver_cluster(CL)
    eng.set_tile_skip(TS)
    eng.plan(775, [40] * B, [50] * B)
    for it in range(3):
        res = eng.align_batch(D1, D2, u0)
    ms = eng.stage_ms()
    n = eng._lib.rb_get_solver_max_clusters(eng._h) * CL
    n = min(n, B * CL)
    buf = np.zeros((n, 8), dtype=np.int64)
    eng._lib.rb_debug_solver_cycles(eng._h, C.c_void_p(buf.ctypes.data), n)
    f = buf.astype(float) / 1.965e3  # us at 1965 MHz (approx)

    print(f"skip={TS} C={CL} ctas={n} solve={ms['solve']:.3f} ms | per-CTA us: total {f[:,0].mean():.0f} matvec {f[:,1].mean():.0f} (max {f[:,1].max():.0f}) combine {f[:,2].mean():.0f} queue {f[:,3].mean():.0f} round(lead) {f[::CL,4].mean():.0f} build {f[:,5].mean():.0f} | passes {res.n_passes.float().sum().item():.0f}")
