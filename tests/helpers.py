"""Shared helpers of the parity tests: the same problem stated for the product (dune_hdd_b200) and for the oracle."""
import json
import os

import numpy as np

from oracle import oracle as o

_GOLDEN = None


def golden(stem, type, partitioning="-", mus="-"):
    """A golden vector of the reference's own tests (tests/golden/reference_expectations.json, transcribed from the
    reference's test/<stem>.cxx by tests/golden/extract_expectations.py): one value per refinement level."""
    global _GOLDEN
    if _GOLDEN is None:
        with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_expectations.json")) as f:
            _GOLDEN = json.load(f)
    return _GOLDEN[stem][partitioning][mus][type]["values"]


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
