import sys, ctypes as C, numpy as np, torch
sys.path.insert(0, '.')
import bench
from roman_b200.engine import Engine, make_invariant_params, method_invariant_kwargs
B = 592
eng = Engine(0)
eng._lib.rb_debug_solver_cycles.restype = C.c_int
eng._lib.rb_debug_solver_cycles(eng._h, None, 0)
eng.set_invariant_params(make_invariant_params(**method_invariant_kwargs("roman", semantics_dim=768)))
D1h, D2h, u0h = bench.make_inputs(1000, B)
D1 = torch.from_numpy(D1h.reshape(-1)).cuda(); D2 = torch.from_numpy(D2h.reshape(-1)).cuda(); u0 = torch.from_numpy(u0h.reshape(-1)).cuda()
eng.set_solver_cluster(int(sys.argv[1]) if len(sys.argv) > 1 else 1)
eng.plan(775, [40] * B, [50] * B)
res = eng.align_batch(D1, D2, u0); torch.cuda.synchronize()
buf = np.zeros((4, 8), dtype=np.int64)
eng._lib.rb_debug_solver_cycles(eng._h, C.c_void_p(buf.ctypes.data), 4)
d = buf[1].astype(float)
us = lambda c: c / 1.965e3
print(f"warp3 of CTAs 0-7: passes {d[5]:.0f} tiles {d[3]:.0f} live rows {d[4]:.0f} | per pass: total {us(d[6])/d[5]:.1f} us wait {us(d[0])/d[5]:.1f} comp {us(d[1])/d[5]:.1f} rowend {us(d[2])/d[5]:.1f} | per tile: wait {us(d[0])/d[3]:.2f} comp {us(d[1])/d[3]:.2f} | per live row: rowend {us(d[2])/d[4]:.2f} us")
