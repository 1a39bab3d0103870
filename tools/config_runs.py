"""Scale runs of the BASELINE configs that are not the bench line (one GPU), JSON on stdout:
    python tools/config_runs.py spe10 [nx ny]     config 2 scaled up: SPE10-shaped Q1 grid, synthetic log-normal permeability
    python tools/config_runs.py p2 [n] [iters]    config 5, polOrder 2 (Q2) on the n x n grid: assembly + CG per-iteration times
"""
import ctypes as C
import json
import sys
import time

import numpy as np

sys.path.insert(0, '.')
import dune_hdd_b200 as hdd  # noqa: E402
from dune_hdd_b200 import capi  # noqa: E402


def kernel_times(d, which=((0, "spmv"), (1, "update"), (2, "direction"), (3, "assembly"))):
    L = capi.lib()
    out = {}
    for w, name in which:
        sec, byt = C.c_double(), C.c_double()
        capi.check(L.hdd_profile_kernel(d._h, w, 5, C.byref(sec)))
        capi.check(L.hdd_kernel_bytes(d._h, w, C.byref(byt)))
        out[name] = {"ms": sec.value * 1e3, "GBs": byt.value / sec.value / 1e9, "frac_of_6553.9": byt.value / sec.value / 1e9 / 6553.9}
    return out


def spe10(nx, ny):
    g = hdd.grids.cube(nx, ny, (0.0, 0.0), (5.0, 1.0))
    prob = hdd.problems.Spe10Model1(g)  # k(x) nearest-neighbour upsampled from the 100 x 20 synthetic layer
    d = hdd.SWIPDG(g, prob)
    t = time.time(); d.init(); t_init = time.time() - t
    ta = min(d.assemble() for _ in range(3))
    out = {"config": "spe10-shaped %dx%d Q1 cells on [0,5]x[0,1], contrast %.1e" % (nx, ny, prob.diffusion_tensor[:, 0].max() / prob.diffusion_tensor[:, 0].min()),
           "dofs": d.num_dofs(), "init_s": t_init, "assemble_ms": ta * 1e3, "assemble_DoFs_s": d.num_dofs() / ta}
    for typ, maxit in (("cg.mg", 5000), ("cg.blockdiagonal", 3000)):
        try:
            _, info = d.uncached_solve({"type": typ, "precision": 1e-10, "max_iter": maxit}, return_info=True, copy_to_host=False)
            out[typ] = {k: info[k] for k in ("iterations", "seconds", "relative_residual", "converged")}
        except hdd.discretizations.linear_solver_failed as e:
            out[typ] = {"not_converged": str(e)[-120:]}
    out["kernels"] = kernel_times(d)
    return out


def p2(n, iters):
    g = hdd.grids.cube(n)
    d = hdd.SWIPDG(g, hdd.problems.ESV2007(), polorder=2)
    t = time.time(); d.init(); t_init = time.time() - t
    ta = min(d.assemble() for _ in range(3))
    out = {"config": "config 5, Q2 on %dx%d cells" % (n, n), "dofs": d.num_dofs(), "nnz": 81 * (g.n_cells + int((g.cell_neigh >= 0).sum())),
           "init_s": t_init, "assemble_ms": ta * 1e3, "assemble_DoFs_s": d.num_dofs() / ta}
    for typ in ("cg.diagonal", "cg.blockdiagonal"):
        try:
            _, info = d.uncached_solve({"type": typ, "precision": 1e-10, "max_iter": iters}, return_info=True, copy_to_host=False)
        except hdd.discretizations.linear_solver_failed as e:
            info = None
            msg = str(e)
        out[typ] = {"kernels": kernel_times(d), "note": "capped at %d iterations" % iters if info is None else info["iterations"],
                    "residual": msg[-60:] if info is None else info["relative_residual"]}
    return out


if __name__ == "__main__":
    what = sys.argv[1]
    if what == "spe10":
        nx, ny = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (5120, 1024)
        print(json.dumps(spe10(nx, ny)))
    else:
        n = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
        print(json.dumps(p2(n, int(sys.argv[3]) if len(sys.argv) > 3 else 200)))
