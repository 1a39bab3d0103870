"""p = 2 timing on one GPU: python tools/quick_bench_p2.py [cells per side] [cube|simplex] [cg iterations]"""
import sys, time, json
import numpy as np
sys.path.insert(0, '.')
import dune_hdd_b200 as hdd
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
kind = sys.argv[2] if len(sys.argv) > 2 else "cube"
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 200
g = hdd.grids.cube(n) if kind == "cube" else hdd.grids.simplex(n)
d = hdd.SWIPDG(g, hdd.problems.ESV2007(), polorder=2)
t = time.time(); d.init(); t_init = time.time() - t
for _ in range(3):
    ta = d.assemble()
nl = d.n_loc
nnz = nl * nl * (g.n_cells + int((g.cell_neigh >= 0).sum()))
out = {"kind": kind, "cells": g.n_cells, "dofs": d.num_dofs(), "nnz": nnz, "init_s": t_init, "assemble_s": ta,
       "assemble_GBs": (8 * nnz + 8 * d.num_dofs()) / ta / 1e9, "assemble_DoFs_s": d.num_dofs() / ta}
for typ in ("cg.diagonal", "cg.blockdiagonal"):
    try:
        u, info = d.uncached_solve({"type": typ, "precision": 1e-30, "max_iter": iters}, return_info=True, copy_to_host=False)
    except hdd.discretizations.linear_solver_failed as e:
        info = None
        msg = str(e)
    import ctypes as C
    from dune_hdd_b200 import capi
    L = capi.lib()
    res = {}
    for which, name in ((0, "spmv"), (1, "update"), (2, "direction"), (3, "assembly")):
        sec, byt = C.c_double(), C.c_double()
        capi.check(L.hdd_profile_kernel(d._h, which, 10, C.byref(sec)))
        capi.check(L.hdd_kernel_bytes(d._h, which, C.byref(byt)))
        res[name] = {"ms": sec.value * 1e3, "GBs": byt.value / sec.value / 1e9}
    out[typ] = res
print(json.dumps(out))
