"""Dev tool: run one align_batch with a given cluster size (for ncu)."""
import sys, numpy as np, torch
sys.path.insert(0, '.')
import bench
from roman_b200.engine import Engine, make_invariant_params, method_invariant_kwargs
B = int(sys.argv[1]); CL = int(sys.argv[2])
eng = Engine(0)
eng.set_invariant_params(make_invariant_params(**method_invariant_kwargs("roman", semantics_dim=768)))
eng.set_solver_cluster(CL)
if len(sys.argv) > 3:
    eng.set_tile_skip(bool(int(sys.argv[3])))
D1h, D2h, u0h = bench.make_inputs(1000, B)
eng.plan(775, [40] * B, [50] * B)
D1 = torch.from_numpy(D1h.reshape(-1)).cuda(); D2 = torch.from_numpy(D2h.reshape(-1)).cuda(); u0 = torch.from_numpy(u0h.reshape(-1)).cuda()
for it in range(3):
    res = eng.align_batch(D1, D2, u0)
torch.cuda.synchronize()
print(eng.solver_cluster, res.n_passes.float().mean().item())
