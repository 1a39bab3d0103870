"""Run under torchrun on a multi-GPU box: the N-rank BlockSWIPDG path against the oracle / single-rank results.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dune_hdd_b200 as hdd  # noqa: E402
from oracle import oracle as o  # noqa: E402


def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    comm = hdd.parallel.init_comm(rank, world, lr)
    # the last case is large enough for the TMA SpMV and therefore for the peer-memory (CUDA IPC / NVLink) halo read
    # alu 64: 128 lattice rows, enough for the strip-distributed multigrid on triangles
    for kind, n in (("alu", 8), ("alu", 64), ("sgrid", 32), ("sgrid", 256)):
        g = (hdd.grids.simplex if kind == "alu" else hdd.grids.cube)(n, partitions=(8, 8) if n in (64, 256) else (4, 4))
        roff = hdd.parallel.rank_cell_offsets(g, world)
        rng = (int(roff[rank]), int(roff[rank + 1]))
        prob = hdd.problems.OS2014ParametricESV2007() if kind == "alu" else hdd.problems.ESV2007()
        d = hdd.BlockSWIPDG(g, prob, device=lr, cell_range=rng, comm=comm)
        d.init()
        nl = g.n_loc
        m = o.Mesh(o.SIMPLEX if kind == "alu" else o.CUBE, g.xy, g.cell_verts, g.cell_neigh)
        rp, col = o.pattern(m)
        mu = 0.5 if kind == "alu" else None
        fac = o.os2014_factor(0.5) if kind == "alu" else o.const(1.0)
        A = o.assemble_lhs(m, fac, None, rp, col)
        b = o.assemble_rhs(m, o.esv2007_force())
        r0, r1 = rng[0] * nl, rng[1] * nl
        # pattern rows of this rank, global columns
        rp_l, col_l = d.pattern()
        assert np.array_equal(rp_l, rp[r0:r1 + 1] - rp[r0]) and np.array_equal(col_l, col[rp[r0]:rp[r1]]), "pattern"
        Al = d.system_matrix().freeze_parameter(mu)
        assert np.abs(Al - A[rp[r0]:rp[r1]]).max() <= 1e-12 * np.abs(A).max(), "entries"
        x = np.cos(np.arange(g.n_dofs, dtype=float))
        y = d.apply(x[r0:r1], mu)  # needs the halo of x: apply() only has the owned part -> halo exchange inside
        y_ref = o.spmv(rp, col, A, x)[r0:r1]
        assert np.abs(y - y_ref).max() <= 1e-12 * np.abs(y_ref).max(), "spmv with halo exchange"
        import scipy.sparse.linalg as spla
        u_ref = spla.spsolve(o.to_scipy(rp, col, A).tocsc(), b)
        u, info = d.solve({"type": "cg.diagonal", "precision": 1e-13, "max_iter": 50000}, mu=mu, return_info=True)
        assert np.abs(u - u_ref[r0:r1]).max() <= 1e-8 * np.abs(u_ref).max(), "solution"
        if kind == "sgrid":
            # multigrid-preconditioned CG: distributed DG level, replicated vertex hierarchy
            um, im = d.solve({"type": "cg.mg", "precision": 1e-13, "max_iter": 500}, return_info=True)
            assert np.abs(um - u_ref[r0:r1]).max() <= 1e-8 * np.abs(u_ref).max(), "cg.mg solution"
            assert im["iterations"] <= 70, "cg.mg iterations"
            ub, ib = d.solve({"type": "cg.blockdiagonal", "precision": 1e-13, "max_iter": 50000}, return_info=True)
            assert np.abs(ub - u_ref[r0:r1]).max() <= 1e-8 * np.abs(u_ref).max(), "cg.blockdiagonal solution"
            # products and error norms are global over the ranks
            import ctypes as C
            from dune_hdd_b200 import capi
            ids = ["l2", "penalty"]
            capi.check(capi.lib().hdd_swipdg_only_these_products(d._h, (C.c_char_p * 2)(*[i.encode() for i in ids]), 2))
            one = np.ones(r1 - r0)
            assert abs(d.get_product("l2").apply2(one, one) - 4.0) <= 1e-11, "l2 product over all ranks"
            Pm = o.to_scipy(rp, col, o.assemble_product(m, "penalty", rp, col))
            assert abs(d.get_product("penalty").apply2(x[r0:r1], x[r0:r1]) - x @ (Pm @ x)) <= 1e-10 * abs(x @ (Pm @ x)), "penalty"
            e = d.error_norms(*hdd.problems.ESV2007_EXACT, vector=u_ref[r0:r1], order=5)
            e_ref = o.error_norms(m, u_ref, o.esv2007_exact(), order=5)
            assert abs(e["L2"] - e_ref["L2"]) <= 1e-8 * e_ref["L2"], "L2 error over all ranks"
            if rank == 0:
                print("cg.mg iterations", im["iterations"], "cg.blockdiagonal", ib["iterations"])
            if n == 32:  # p = 2 on the distributed mesh (generic halo exchange with 9 DoFs per cell)
                d2 = hdd.BlockSWIPDG(g, prob, polorder=2, device=lr, cell_range=rng, comm=comm)
                d2.init()
                m2 = m.with_polorder(2)
                rp2, col2 = o.pattern(m2)
                A2 = o.assemble_lhs(m2, fac, None, rp2, col2)
                b2 = o.assemble_rhs(m2, o.esv2007_force())
                q0, q1 = rng[0] * 9, rng[1] * 9
                assert np.abs(d2.system_matrix().affine_part() - A2[rp2[q0]:rp2[q1]]).max() <= 1e-12 * np.abs(A2).max(), "p2 entries"
                u2_ref = spla.spsolve(o.to_scipy(rp2, col2, A2).tocsc(), b2)
                u2 = d2.solve({"type": "cg.diagonal", "precision": 1e-13, "max_iter": 50000})
                assert np.abs(u2 - u2_ref[q0:q1]).max() <= 1e-8 * np.abs(u2_ref).max(), "p2 solution"
                del d2
        if kind == "alu":
            # multigrid-preconditioned CG on the triangle lattice (replicated vertex levels on N ranks)
            ua, ia = d.solve({"type": "cg.mg", "precision": 1e-13, "max_iter": 500}, mu=mu, return_info=True)
            assert np.abs(ua - u_ref[r0:r1]).max() <= 1e-8 * np.abs(u_ref).max(), "cg.mg solution on simplices"
            assert ia["iterations"] <= 45, "cg.mg iterations on simplices"
            if rank == 0:
                print("cg.mg iterations on triangles", n, ia["iterations"])
            prm = {"mu": 0.5, "mu_bar": 0.5, "mu_hat": 1.0, "parameter_range_min": 0.1, "parameter_range_max": 1.0}
            ind_ref = o.indicators(m, u_ref, o.esv2007_force(), o.os2014_factor(0.5), a_hat=o.os2014_factor(1.0),
                                   a_bar=o.os2014_factor(0.5), a_min=o.os2014_factor(0.1), a_max=o.os2014_factor(1.0))
            ind = d.indicators(u, prm)
            for k in ("nc2", "res2", "df2", "dfstar2", "resstar2"):
                ref = ind_ref[k][rng[0]:rng[1]]
                assert np.abs(ind[k] - ref).max() <= 1e-7 * np.abs(ind_ref[k]).max(), k
            etas = {t: d.estimate(u, t, prm) for t in ("eta_NC_OS2014", "eta_R_OS2014", "eta_DF_OS2014", "eta_OS2014", "eta_OS2014_*")}
            assert abs(etas["eta_NC_OS2014"] - np.sqrt(ind_ref["nc2"].sum())) <= 1e-7 * etas["eta_NC_OS2014"]
            assert abs(etas["eta_DF_OS2014"] - np.sqrt(ind_ref["df2"].sum())) <= 1e-7 * etas["eta_DF_OS2014"]
            loc = d.estimate_local(u, "eta_OS2014", prm)
            assert loc.shape == (g.n_subdomains,) and np.all(loc > 0)
            if rank == 0:
                print("estimates", etas)
        if kind == "sgrid" and n == 256:
            # the same grid from the cube provider (tables written on the device, rank-local closed forms)
            pr = hdd.grids.CubeProvider(n, partitions=(8, 8))
            dp = hdd.BlockSWIPDG(pr, prob, device=lr, cell_range=rng, comm=comm)
            dp.init()
            rp_p, col_p = dp.pattern()
            assert np.array_equal(rp_p, rp_l) and np.array_equal(col_p, col_l), "cube provider pattern"
            assert np.array_equal(dp.system_matrix().affine_part(), Al), "cube provider entries"
            up, ip = dp.solve({"type": "cg.mg", "precision": 1e-13, "max_iter": 500}, return_info=True)
            assert np.abs(up - u_ref[r0:r1]).max() <= 1e-8 * np.abs(u_ref).max(), "cube provider cg.mg solution"
            assert ip["iterations"] == im["iterations"], "cube provider cg.mg iterations"
            uq = dp.solve({"type": "cg.diagonal", "precision": 1e-13, "max_iter": 50000})
            assert np.abs(uq - u_ref[r0:r1]).max() <= 1e-8 * np.abs(u_ref).max(), "cube provider Jacobi solution"
            del dp
        if rank == 0:
            print("multi-GPU check ok:", kind, n, "world", world, "iterations", info["iterations"])
        del d
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
