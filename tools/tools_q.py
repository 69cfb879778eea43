import sys; sys.path.insert(0,'.')
from roman_b200.engine import Engine
e=Engine(0)
for m in [(40,50),(63,64),(20,20)]:
  for c in [1,2,4,8]:
    e.set_solver_cluster(c); e.plan(3,[m[0]]*300,[m[1]]*300)
    print(m, c, e.solver_cluster, e._lib.rb_get_solver_max_clusters(e._h))
