"""CPU checks of the oracle parts no reference golden pins (polOrder 2, product matrices, Neumann data): since the
reference prints nothing for them ("parity unpinned"), the oracle is checked against mathematical
identities instead - convergence orders against the analytic ESV2007 solution, symmetry, partition of unity,
consistency of the products with the system matrix."""
import numpy as np
import pytest
import scipy.sparse.linalg as spla

from oracle import oracle as o


def _solve(m):
    rp, col = o.pattern(m)
    A = o.assemble_lhs(m, o.const(1.0), None, rp, col)
    b = o.assemble_rhs(m, o.esv2007_force())
    S = o.to_scipy(rp, col, A)
    return S, b, spla.spsolve(S.tocsc(), b)


def test_p2_simplex_converges_with_third_order():
    errs = []
    for n in (2, 4, 8):
        m = o.mesh_bisect(n, -1.0, 1.0, 2).with_polorder(2)
        S, b, u = _solve(m)
        assert abs(S - S.T).max() <= 1e-13 * abs(S).max()
        errs.append(o.error_norms(m, u, o.esv2007_exact(), order=8))
    assert np.log2(errs[1]["L2"] / errs[2]["L2"]) > 2.8
    assert np.log2(errs[1]["H1_semi"] / errs[2]["H1_semi"]) > 1.8


def test_q2_with_the_reference_quadrature_order_converges():
    """volume rule order(a) + 2(p-1) = 2 -> 2x2 Gauss: under-integrated for Q2 exactly as the 1-point rule is for Q1
    (SURVEY 0.4); the scheme still converges, with second order in L2"""
    errs = []
    for n in (4, 8, 16):
        m = o.mesh_cube(n, n, -1.0, 1.0, -1.0, 1.0).with_polorder(2)
        S, b, u = _solve(m)
        assert m.n_dofs == 9 * n * n
        assert abs(S - S.T).max() <= 1e-13 * abs(S).max()
        errs.append(o.error_norms(m, u, o.esv2007_exact(), order=8)["L2"])
    assert np.log2(errs[1] / errs[2]) > 1.8
    assert errs[2] < 0.25 * 1.13e-2  # far below the Q1 error of the committed golden at the same h


@pytest.mark.parametrize("kind,p", [("alu", 1), ("alu", 2), ("sgrid", 1), ("sgrid", 2)])
def test_products_identities(kind, p):
    m = (o.mesh_bisect(2, -1.0, 1.0, 2) if kind == "alu" else o.mesh_cube(6, 6, -1.0, 1.0, -1.0, 1.0)).with_polorder(p)
    rpv, colv = o.pattern_volume(m)
    rp, col = o.pattern(m)
    one = np.ones(m.n_dofs)
    M = o.to_scipy(rpv, colv, o.assemble_product(m, "l2", rpv, colv))
    assert abs(one @ (M @ one) - 4.0) <= 1e-12                   # |Omega|
    H = o.to_scipy(rpv, colv, o.assemble_product(m, "h1_semi", rpv, colv))
    assert abs(H @ one).max() <= 1e-12 and abs(H - H.T).max() <= 1e-13
    B = o.to_scipy(rpv, colv, o.assemble_product(m, "boundary_l2", rpv, colv))
    assert abs(one @ (B @ one) - 8.0) <= 1e-12                   # |dOmega|
    E = o.to_scipy(rpv, colv, o.assemble_product(m, "elliptic", rpv, colv, factor=o.const(3.0)))
    assert abs(E - 3.0 * H).max() <= 1e-12
    P = o.to_scipy(rp, col, o.assemble_product(m, "penalty", rp, col))
    assert abs(P - P.T).max() <= 1e-13
    # jumps of a continuous function vanish: the penalty form of u = 1 only sees the Dirichlet boundary
    sb = 14.0 if p == 1 else 38.0
    h = 0.5 if kind == "alu" else 2.0 / 6.0  # boundary face length
    assert abs(one @ (P @ one) - sb / h * 8.0) <= 1e-9
    if p == 1 and kind == "alu":
        # on P1 simplices the volume rule is exact, so system = elliptic + consistency terms + penalty: the difference
        # A - E - P is the (symmetric) consistency part, which vanishes on constants tested against interior bubbles
        A = o.to_scipy(rp, col, o.assemble_lhs(m, o.const(1.0), None, rp, col))
        C = A - P - o.to_scipy(rpv, colv, o.assemble_product(m, "h1_semi", rpv, colv))
        assert abs(C - C.T).max() <= 1e-12
        assert abs(C.diagonal()).max() > 0


def test_neumann_rhs_integrates_the_data_over_the_neumann_faces():
    for kind, p in (("alu", 1), ("sgrid", 2)):
        m = (o.mesh_bisect(2, -1.0, 1.0, 2) if kind == "alu" else o.mesh_cube(4, 4, -1.0, 1.0, -1.0, 1.0)).with_polorder(p)
        nf = 3 if kind == "alu" else 4
        bt = np.ones((m.nc, nf), np.uint8)
        cv = m.xy[m.cv]
        left = (cv[:, :, 0].mean(axis=1) < 0)
        bt[left[:, None] & (m.nb < 0)] = 2
        gn = o.fn([(1.0, o.FN_ONE), (1.0, o.FN_Y)], 1)
        b = o.assemble_rhs(m, None, neumann=gn, bnd_type=bt)
        # sum_i b_i = int_{Gamma_N} g_N; Gamma_N = boundary faces of cells with centre x < 0
        on = (bt == 2) & (m.nb < 0)
        fv = [(0, 1), (0, 2), (1, 2)] if kind == "alu" else [(0, 2), (1, 3), (0, 1), (2, 3)]
        total = 0.0
        for c, f in zip(*np.nonzero(on)):
            a, e = cv[c, fv[f][0]], cv[c, fv[f][1]]
            total += np.linalg.norm(e - a) * (1.0 + 0.5 * (a[1] + e[1]))
        assert abs(b.sum() - total) <= 1e-12
        assert np.all(o.assemble_rhs(m, None, neumann=gn) == 0.0)  # AllDirichlet: no Neumann faces


def test_threaded_oracle_equals_the_serial_walk():
    """or_set_threads only changes who adds which entry (atomically): same matrix, rhs and CG iterates to rounding"""
    m = o.mesh_cube(48, 48, -1.0, 1.0, -1.0, 1.0)
    rp, col = o.pattern(m)
    A1 = o.assemble_lhs(m, o.const(1.0), None, rp, col)
    b1 = o.assemble_rhs(m, o.esv2007_force())
    x1, it1, _ = o.cg(rp, col, A1, b1, precond=1, rtol=1e-30, maxit=30)
    try:
        o.set_threads(4)
        A4 = o.assemble_lhs(m, o.const(1.0), None, rp, col)
        b4 = o.assemble_rhs(m, o.esv2007_force())
        x4, it4, _ = o.cg(rp, col, A4, b4, precond=1, rtol=1e-30, maxit=30)
    finally:
        o.set_threads(1)
    assert np.abs(A1 - A4).max() <= 1e-14 * np.abs(A1).max() and np.array_equal(b1, b4)
    assert it1 == it4 and np.abs(x1 - x4).max() <= 1e-11 * np.abs(x1).max()


def test_reference_indicators_localise_the_energy_error():
    """studies.localize_energy (compute_reference_indicators, test/linearelliptic-swipdg.hh:133-223) on oracle data: the
    per-cell energies of u_ref - P u_h from the elliptic product blocks add up to the squared energy norm, and the
    indicators are that energy per coarse cell over (total * children)"""
    from dune_hdd_b200.studies import localize_energy
    coarse, fine = o.mesh_bisect(2, -1.0, 1.0, 2), o.mesh_bisect(2, -1.0, 1.0, 6)
    def solve(m):
        rp, col = o.pattern(m)
        import scipy.sparse.linalg as spla
        return spla.spsolve(o.to_scipy(rp, col, o.assemble_lhs(m, o.os2014_factor(0.5), None, rp, col)).tocsc(),
                            o.assemble_rhs(m, o.esv2007_force()))
    father = o.fathers(coarse, fine)
    d = solve(fine) - o.prolong(coarse, solve(coarse), fine, father)
    rpv, colv = o.pattern_volume(fine)
    vals = o.assemble_product(fine, "elliptic", rpv, colv, factor=o.os2014_factor(0.5))
    total = d @ (o.to_scipy(rpv, colv, vals) @ d)
    ind = localize_energy(vals.reshape(fine.nc, 3, 3), d.reshape(fine.nc, 3), father, coarse.nc)
    assert ind.shape == (coarse.nc,) and (ind > 0).all()
    children = np.bincount(father, minlength=coarse.nc)
    assert set(children) == {16} and abs((ind * children).sum() - 1.0) < 1e-12
    local = np.array([d[3 * c:3 * c + 3] @ vals[9 * c:9 * c + 9].reshape(3, 3) @ d[3 * c:3 * c + 3] for c in range(fine.nc)])
    assert abs(local.sum() - total) <= 1e-12 * total
    assert np.allclose(ind, np.bincount(father, weights=local) / (total * 16), rtol=1e-12)
