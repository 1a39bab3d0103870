"""ncu target: one ESV2007 estimate on a 2M-triangle grid (k_vertex_means + k_indicators + segment reductions)."""
import sys, time
sys.path.insert(0, '.')
import numpy as np
import dune_hdd_b200 as hdd
s = int(sys.argv[1]) if len(sys.argv) > 1 else 512
g = hdd.grids.simplex(s, partitions=(8, 8))
d = hdd.BlockSWIPDG(g, hdd.problems.ESV2007()); d.init()
u = np.cos(np.arange(g.n_dofs) * 1e-3)
for _ in range(3):
    t = time.time(); eta = d.estimate(u, "eta_ESV2007"); dt = time.time() - t
print("cells", g.n_cells, "eta", eta, "estimate wall ms", dt * 1e3)
