"""Dev tool: wall time of roman_b200.align.align_submap_pairs on one 20 x 20 all-pairs search (the cfg5 robot pair of
bench.py, seed 5000): whole function (host packing, planning, upload, kernels, read-back, vectorised epilogue) and the
batched registration inside it."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from roman_b200 import synth
from roman_b200.align import SubmapAlignParams, align_submap_pairs
from roman_b200.align.objects import objects_from_rows

w = synth.make_all_pairs_world(5000, 20, 20, 40, method="roman", semantics_dim=768, desc_common=1.0)
sub0 = [objects_from_rows(t, "roman", 768) for t in w.tables[0]]
sub1 = [objects_from_rows(t, "roman", 768) for t in w.tables[1]]
smp = SubmapAlignParams(method="roman")
rng = np.random.default_rng(0)
u0 = lambda i, j, m: rng.uniform(0, 1, m)
if "--dry" in sys.argv:
    print("submaps", len(sub0), len(sub1), "objects", sum(map(len, sub0)), sum(map(len, sub1)))
    sys.exit(0)
for it in range(4):
    timing = {}
    t0 = time.perf_counter()
    res = align_submap_pairs(smp, sub0, sub1, u0=u0, timing=timing)
    dt = time.perf_counter() - t0
    print(f"call {it}: align_submap_pairs {dt * 1e3:.1f} ms for {timing.get('n_registered')} pairs "
          f"(registration {timing.get('registration_seconds', 0) * 1e3:.1f} ms; "
          f"{int((res.status_mat == 0).sum())} aligned)", flush=True)
