"""Dev tool: single-pair latency through the reference-facing API (register + T_align), vs batch."""
import sys, time, numpy as np, torch
sys.path.insert(0, '.')
from roman_b200 import synth
from roman_b200.align import SubmapAlignParams
from roman_b200.align.objects import objects_from_rows
reg = SubmapAlignParams(method="roman").get_object_registration()
p = synth.make_pair(0, 40, 50)
m1, m2 = objects_from_rows(p.D1, "roman", 768), objects_from_rows(p.D2, "roman", 768)
for it in range(3):
    reg.register(m1, m2, u0=p.u0)
def t(f, n=20):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): f()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
print("register (pack + score + solve + fetch): %.2f ms" % t(lambda: reg.register(m1, m2, u0=p.u0)))
a = reg.register(m1, m2, u0=p.u0)
print("T_align: %.2f ms" % t(lambda: reg.T_align(m1, m2, a)))
rows1 = np.array([reg._object_to_clipper_list(o) for o in m1]); rows2 = np.array([reg._object_to_clipper_list(o) for o in m2])
print("  of which python packing of the rows: %.2f ms" % t(lambda: (np.array([reg._object_to_clipper_list(o) for o in m1]), np.array([reg._object_to_clipper_list(o) for o in m2]))))
from roman_b200 import clipperpy
def core():
    c = reg._setup_clipper(); c.score_pairwise_and_single_consistency(rows1.T, rows2.T, clipperpy.utils.create_all_to_all(40, 50)); c.solve(p.u0); return c.get_selected_associations()
print("  clipperpy score+solve+get (arrays ready): %.2f ms" % t(core))
print("register_batch of 64 pairs: %.2f ms per pair" % (t(lambda: reg.register_batch([(m1, m2)] * 64, [p.u0] * 64), 3) / 64))
