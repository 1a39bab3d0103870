"""Shared helpers of the parity tests: the same problem stated for the product (dune_hdd_b200) and for the oracle."""
import numpy as np

from oracle import oracle as o


def oracle_mesh(grid):
    return o.Mesh(o.SIMPLEX if grid.kind == 0 else o.CUBE, grid.xy, grid.cell_verts, grid.cell_neigh)


def oracle_system(grid, factor, force, tensor=None):
    m = oracle_mesh(grid)
    rp, col = o.pattern(m)
    A = o.assemble_lhs(m, factor, tensor, rp, col)
    b = o.assemble_rhs(m, force)
    return m, rp, col, A, b


def direct_solve(rp, col, val, b):
    import scipy.sparse.linalg as spla
    return spla.spsolve(o.to_scipy(rp, col, val).tocsc(), b)


def rel(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)
