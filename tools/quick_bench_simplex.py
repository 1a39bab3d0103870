import time, sys, numpy as np
sys.path.insert(0, '.')
import ctypes as C
import dune_hdd_b200 as hdd
from dune_hdd_b200 import capi
s = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
t=time.time(); g = hdd.grids.simplex(s, partitions=(8,8)); print("grid s", time.time()-t, g.n_cells)
for prob, mu in ((hdd.problems.ESV2007(), None), (hdd.problems.OS2014ParametricESV2007(), 0.5)):
    t=time.time(); d = hdd.BlockSWIPDG(g, prob); d.init(); capi.check(capi.lib().hdd_sync(d._h)); print(prob.name, "create+init s", time.time()-t)
    for _ in range(3): ta = d.assemble()
    print("  assemble ms", ta*1e3, "DoFs/s", g.n_dofs/ta)
    u, info = d.uncached_solve({"type":"cg.diagonal","precision":1e-10,"max_iter":100000}, mu=mu, return_info=True, copy_to_host=False)
    print("  cg", info["iterations"], "s", info["seconds"], "ms/it", info["seconds_per_iteration"]*1e3)
    L = capi.lib()
    for which, name in ((0,"spmv"),(1,"update"),(2,"direction"),(3,"assembly")):
        sec, byt = C.c_double(), C.c_double()
        capi.check(L.hdd_profile_kernel(d._h, which, 10, C.byref(sec))); capi.check(L.hdd_kernel_bytes(d._h, which, C.byref(byt)))
        print("  %-10s %.3f ms  %.0f GB/s" % (name, sec.value*1e3, byt.value/sec.value/1e9))
    u, info = d.uncached_solve({"type":"cg.diagonal","precision":1e-10,"max_iter":100000}, mu=mu, return_info=True, copy_to_host=False)
    prm = None if mu is None else {"mu": mu, "mu_bar": mu, "mu_hat": 1.0, "parameter_range_min": 0.1, "parameter_range_max": 1.0}
    typ = "eta_ESV2007" if mu is None else "eta_OS2014"
    for _ in range(3):
        t=time.time(); eta = d.estimate(None, typ, prm); te=time.time()-t
    print("  estimate", typ, eta, "ms", te*1e3)
    del d
