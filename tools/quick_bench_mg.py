import sys, json
sys.path.insert(0, '.')
import dune_hdd_b200 as hdd
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
g = hdd.grids.cube(n, partitions=(8, 8) if n % 8 == 0 else (1, 1))
d = hdd.BlockSWIPDG(g, hdd.problems.ESV2007())
d.init()
out = {"n": n}
for typ in sys.argv[2:] or ["cg.mg", "cg.mg"]:
    u, info = d.uncached_solve({"type": typ, "precision": 1e-10, "max_iter": 200000}, return_info=True, copy_to_host=False)
    e = d.error_norms(*hdd.problems.ESV2007_EXACT, order=5)
    out.setdefault(typ, []).append({"iterations": info["iterations"], "seconds": info["seconds"], "relres": info["relative_residual"], "L2": e["L2"]})
print(json.dumps(out))
