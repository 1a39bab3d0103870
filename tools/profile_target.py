"""ncu targets, one small job per mode (run from the repository root on a GPU box):
    python tools/profile_target.py est [s]     one eta_ESV2007 evaluation on 8 s^2 triangles (default s = 1448)
    python tools/profile_target.py p1 [s]      P1 assembly on 8 s^2 triangles
    python tools/profile_target.py q2 [n]      Q2 assembly on n^2 cells (default 2048)
    python tools/profile_target.py q1 [n]      Q1 assembly + cg.mg solve on n^2 cells (default 4096)
Prints the CUDA-event kernel time next to the algorithmic bytes, so the ncu capture has its plain-run number beside it."""
import ctypes as C
import sys
import time

import numpy as np

sys.path.insert(0, '.')
import dune_hdd_b200 as hdd  # noqa: E402
from dune_hdd_b200 import capi  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "est"
L = capi.lib()


def kernel(d, which, reps=5):
    sec, byt = C.c_double(), C.c_double()
    capi.check(L.hdd_profile_kernel(d._h, which, reps, C.byref(sec)))
    capi.check(L.hdd_kernel_bytes(d._h, which, C.byref(byt)))
    return sec.value * 1e3, byt.value / sec.value / 1e9


if mode in ("est", "p1"):
    s = int(sys.argv[2]) if len(sys.argv) > 2 else 1448
    g = hdd.grids.simplex(s, partitions=(8, 8))
    d = hdd.BlockSWIPDG(g, hdd.problems.ESV2007())
    d.init()
    for _ in range(3):
        ta = d.assemble()
    print("P1 assemble (matrix + rhs) ms", ta * 1e3, "kernel ms / GB/s", kernel(d, 3))
    if mode == "est":
        v = g.xy[g.cell_verts]
        u = (np.cos(0.5 * np.pi * v[..., 0]) * np.cos(0.5 * np.pi * v[..., 1])).reshape(-1)
        for _ in range(2):
            t = time.time(); eta = d.estimate(u, "eta_ESV2007"); dt = time.time() - t
        print("cells", g.n_cells, "eta", eta, "estimate wall ms", dt * 1e3)
        print("estimator pass ms / GB/s", kernel(d, 4), "indicator kernel", kernel(d, 5))
elif mode == "q2":
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
    g = hdd.grids.CubeProvider(n)
    d = hdd.SWIPDG(g, hdd.problems.ESV2007(), polorder=2)
    d.init()
    for _ in range(3):
        ta = d.assemble()
    nnz = 81 * (n * n + 4 * n * (n - 1))
    print("Q2 %d^2 assemble (matrix + rhs) ms" % n, ta * 1e3, "kernel ms / GB/s", kernel(d, 3), "8 B/nnz GB", 8 * nnz / 1e9)
else:
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
    g = hdd.grids.CubeProvider(n, partitions=(8, 8))
    d = hdd.BlockSWIPDG(g, hdd.problems.ESV2007())
    d.init()
    for _ in range(2):
        ta = d.assemble()
        u, info = d.uncached_solve({"type": "cg.mg", "precision": 1e-10, "max_iter": 1000}, return_info=True, copy_to_host=False)
    print("Q1 %d^2 assemble ms" % n, ta * 1e3, "cg.mg", info["iterations"], "iterations", info["seconds"], "s")
    print("assembly kernel ms / GB/s", kernel(d, 3))
