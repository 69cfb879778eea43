"""Dev tool: stage times vs batch size (not part of the product or the tests)."""
import sys, numpy as np, torch
sys.path.insert(0, '.')
import bench
from roman_b200.engine import Engine, make_invariant_params, method_invariant_kwargs
eng = Engine(0)
eng.set_invariant_params(make_invariant_params(**method_invariant_kwargs("roman", semantics_dim=768)))
eng.set_profiling(True)
Bmax = int(sys.argv[1]) if len(sys.argv) > 1 else 1184
D1h, D2h, u0h = bench.make_inputs(1000, min(Bmax, 296))
rep = (Bmax + 295) // 296
D1h = np.concatenate([D1h] * rep)[:Bmax]; D2h = np.concatenate([D2h] * rep)[:Bmax]; u0h = np.concatenate([u0h] * rep)[:Bmax]
import itertools
for B, CL, TS in itertools.product([592], [1, 2, 4], [True, False]):
    if B > Bmax: break
    eng.set_solver_cluster(CL)
    eng.set_tile_skip(TS)
    eng.plan(775, [40] * B, [50] * B)
    D1 = torch.from_numpy(D1h[:B].reshape(-1)).cuda(); D2 = torch.from_numpy(D2h[:B].reshape(-1)).cuda(); u0 = torch.from_numpy(u0h[:B].reshape(-1)).cuda()
    res = eng.alloc_result()
    acc = {}
    for it in range(6):
        eng.align_batch(D1, D2, u0, out=res)
        ms = eng.stage_ms()
        if it >= 2:
            for k, v in ms.items(): acc[k] = acc.get(k, 0) + v / 4
    p = res.n_passes.cpu().numpy()
    gb = p.sum() * (8.0 * 2000 * 1999 / 2 + 24 * 2000) / 1e9
    print(f"B={B:5d} C={CL} skip={TS} prep {acc['prep']:.3f} aff {acc['affinity']:.3f} solve {acc['solve']:.3f} arun {acc['arun']:.3f} ms | passes mean {p.mean():.2f} max {p.max()} | K2 {gb/acc['solve']*1e3:.0f} GB/s | K1 {B*16.0e6/1e9/(acc['affinity'])*1e3:.0f} GB/s(aff only)")
