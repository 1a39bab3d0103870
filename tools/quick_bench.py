import time, sys, numpy as np
sys.path.insert(0, '.')
import dune_hdd_b200 as hdd
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
t=time.time(); g = hdd.grids.cube(n); print("grid s", time.time()-t, g.n_cells)
t=time.time(); d = hdd.SWIPDG(g, hdd.problems.ESV2007()); print("create s", time.time()-t)
t=time.time(); d.init(); print("init s", time.time()-t)
for _ in range(3): ta = d.assemble()
nnz = 16*(g.n_cells + (g.cell_neigh>=0).sum())
print("assemble s", ta, "GB/s", (8*nnz + 8*g.n_dofs)/ta/1e9, "DoFs/s", g.n_dofs/ta)
u, info = d.uncached_solve({"type":"cg.diagonal","precision":1e-10,"max_iter":100000}, return_info=True, copy_to_host=False)
print(info)
bytes_it = 8*nnz + 4*5*g.n_cells + 8*g.n_cells + 8*g.n_dofs*2 + 80*g.n_dofs
print("CG GB/s (block-CSR bytes)", bytes_it/info["seconds_per_iteration"]/1e9)
