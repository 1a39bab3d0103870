"""config 4 at scale on one GPU: block-SWIPDG P1 on 8 s^2 triangles, 8 x 8 subdomains: python tools/quick_bench_config4.py [s]"""
import sys, time, json
sys.path.insert(0, '.')
import dune_hdd_b200 as hdd
s = int(sys.argv[1]) if len(sys.argv) > 1 else 1408
t = time.time(); g = hdd.grids.simplex(s, partitions=(8, 8)); t_grid = time.time() - t
t = time.time(); d = hdd.BlockSWIPDG(g, hdd.problems.ESV2007()); d.init(); t_init = time.time() - t
out = {"triangles": g.n_cells, "dofs": g.n_dofs, "grid_s": t_grid, "create_init_s": t_init}
for _ in range(3):
    ta = d.assemble()
out["assemble_ms"] = ta * 1e3
for typ, maxit in (("cg.mg", 500), ("cg.mg", 500), ("cg.blockdiagonal", 300)):
    try:
        u, info = d.uncached_solve({"type": typ, "precision": 1e-10, "max_iter": maxit}, return_info=True, copy_to_host=False)
        out[typ] = {"iterations": info["iterations"], "seconds": info["seconds"], "residual": d.residual()}
    except hdd.discretizations.linear_solver_failed as e:
        out[typ] = str(e)[-90:]
    if typ == "cg.mg":
        out["errors"] = d.error_norms(*hdd.problems.ESV2007_EXACT, order=5)
        t = time.time(); out["eta_ESV2007"] = d.estimate(None, "eta_ESV2007"); out["estimate_wall_ms"] = 1e3 * (time.time() - t)
print(json.dumps(out))
