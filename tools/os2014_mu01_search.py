"""Search for the arithmetic behind the reference's mu = 0.1 OS2014 goldens
(/root/reference/test/linearelliptic-block-swipdg-expectations_os2014_2daluconform.cxx:155-167 and :185-197).

TEST INFRASTRUCTURE: drives the CPU oracle, never the product.

Every golden whose solve has a constant diffusion factor (all ESV2007 rows, the OS2014 rows with mu = 1) is reproduced
by the oracle to the three printed digits.  The rows whose *solve* uses mu = 0.1 (a_f = 1 + 0.675 sin(4 pi (x + y/2)))
are not.  This script sweeps the variants of the dune-gdt arithmetic that are plausible for a non-constant factor and
prints, for (mu, mu_bar, mu_hat) = (0.1, 0.1, 0.1) and (0.1, 0.1, 1), all six golden columns on the four levels plus the
energy error the goldens imply (eta / eff), next to the reference's numbers:

  * quadrature order of the factor in the volume and in the face terms of the solve (independently, 0 ... 5),
  * the factor as written / evaluated at the cell centre / P0-projected (in the solve, and in solve + estimator),
  * the penalty's a_f at the face midpoint instead of the quadrature point,
  * harmonic instead of arithmetic face mean of a_f in the penalty, weights omega from a_f K instead of K
    (for a continuous factor evaluated pointwise both coincide with the restatement - they only differ for the P0
    variants, and are swept there),
  * quadrature order of the factor inside the estimator (flux reconstruction and eta_DF integrand).

Usage:  python tools/os2014_mu01_search.py [--levels 4] [--out tests/golden/os2014_mu01_search.txt]
"""
import argparse
import itertools
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import oracle as o  # noqa: E402
from tests.helpers import direct_solve, golden  # noqa: E402

OS14 = "linearelliptic-block-swipdg-expectations_os2014_2daluconform"
COLS = ("eta_DF_OS2014", "eta_DF_OS2014_*", "eta_OS2014", "eta_OS2014_*", "eff_OS2014_mu", "eff_OS2014_*_mu")
MU = 0.1


def a_exact(x, y, mu):
    return 1.0 + 0.75 * (1.0 - mu) * np.sin(4.0 * np.pi * (x + 0.5 * y))


def factor(mesh, mu, kind, order=3):
    """the diffusion factor at mu as the oracle sees it: 'exact' = the Expression, 'centre' = its value at the cell
    centre, 'mean' = its cell mean (7-point rule)"""
    if kind == "exact":
        f = o.os2014_factor(mu)
        f.order = order
        return f
    v = mesh.xy[mesh.cv]
    if kind == "centre":
        c = v.mean(axis=1)
        return o.cellwise(a_exact(c[:, 0], c[:, 1], mu))
    x, y, w = o.element_rule(o.SIMPLEX, 5)
    pts = v[:, 0, None, :] + x[None, :, None] * (v[:, 1, None, :] - v[:, 0, None, :]) \
        + y[None, :, None] * (v[:, 2, None, :] - v[:, 0, None, :])
    return o.cellwise((a_exact(pts[..., 0], pts[..., 1], mu) * w).sum(axis=1) * 2.0)


def subdomain_eta_r(mesh, res2, amin, k=4):
    c = mesh.xy[mesh.cv].mean(axis=1)
    sx = np.clip(((c[:, 0] + 1.0) / 2.0 * k).astype(int), 0, k - 1)
    sy = np.clip(((c[:, 1] + 1.0) / 2.0 * k).astype(int), 0, k - 1)
    sub = sy * k + sx
    diam = np.sqrt(2.0) * 2.0 / k
    tot = 0.0
    for s in range(k * k):
        sel = sub == s
        tot += diam ** 2 / np.pi ** 2 / amin[sel].min() * res2[sel].sum()
    return np.sqrt(tot)


def solve(mesh, f, flags, vol_order, face_order):
    o.lib().or_set_variant(flags, vol_order, face_order)
    try:
        rp, col = o.pattern(mesh)
        A = o.assemble_lhs(mesh, f, None, rp, col)
    finally:
        o.lib().or_set_variant(0, -1, -1)
    b = o.assemble_rhs(mesh, o.esv2007_force())
    return direct_solve(rp, col, A, b)


_MESH = {}


def mesh(level):
    if level not in _MESH:
        _MESH[level] = o.mesh_bisect(4, -1.0, 1.0, 2 + 2 * level)
    return _MESH[level]


def run_variant(v, levels):
    """returns {(mu_hat, column): [value per level]} and the energy errors"""
    ref = mesh(4)
    u_ref = solve(ref, factor(ref, MU, v["solve_factor"], 3), v["flags"], v["vol_order"], v["face_order"])
    rpv, colv = o.pattern_volume(ref)
    E = o.to_scipy(rpv, colv, o.assemble_product(ref, "elliptic", rpv, colv, factor=o.os2014_factor(MU)))
    out = {}
    for level in range(levels):
        m = mesh(level)
        u = solve(m, factor(m, MU, v["solve_factor"], 3), v["flags"], v["vol_order"], v["face_order"])
        d = u_ref - o.prolong(m, u, ref)
        energy = float(np.sqrt(d @ (E @ d)))
        out.setdefault("energy", []).append(energy)
        ek, eo = v["est_factor"], v["est_order"]
        for mu_hat in (0.1, 1.0):
            ind = o.indicators(m, u, o.esv2007_force(), factor(m, MU, ek, eo), a_hat=factor(m, mu_hat, ek, eo),
                               a_bar=factor(m, MU, ek, eo), a_min=o.os2014_factor(0.1), a_max=o.os2014_factor(1.0))
            e_nc, e_df, e_dfs = (float(np.sqrt(ind[k].sum())) for k in ("nc2", "df2", "dfstar2"))
            e_r = subdomain_eta_r(m, ind["res2"], ind["amin"])
            e_rs = subdomain_eta_r(m, ind["resstar2"], ind["amin"])
            ratio = MU / mu_hat  # alpha = gamma = theta(mu) / theta(mu_hat), single component
            eta = e_nc + e_r + max(np.sqrt(ratio), 1.0 / np.sqrt(ratio)) * e_df
            eta_star = e_nc + e_rs + e_dfs / np.sqrt(ratio)
            vals = dict(zip(COLS, (e_df, e_dfs, eta, eta_star, eta / energy, eta_star / energy)))
            for c in COLS:
                out.setdefault((mu_hat, c), []).append(vals[c])
    return out


def run_amplitude(A, levels):
    """the whole pipeline with the problem's amplitude changed: a(mu) = 1 + A (1 - mu) sin(4 pi (x + y/2)), reference A = 0.75"""
    def fac(mu):
        return o.fn([(1.0, o.FN_ONE), (A * (1.0 - mu), o.FN_OS_SIN)], 3)

    ref = mesh(4)
    u_ref = solve(ref, fac(MU), 0, -1, -1)
    rpv, colv = o.pattern_volume(ref)
    E = o.to_scipy(rpv, colv, o.assemble_product(ref, "elliptic", rpv, colv, factor=fac(MU)))
    out = {}
    for level in range(levels):
        m = mesh(level)
        u = solve(m, fac(MU), 0, -1, -1)
        d = u_ref - o.prolong(m, u, ref)
        energy = float(np.sqrt(d @ (E @ d)))
        out.setdefault("energy", []).append(energy)
        for mu_hat in (0.1, 1.0):
            ind = o.indicators(m, u, o.esv2007_force(), fac(MU), a_hat=fac(mu_hat), a_bar=fac(MU), a_min=fac(0.1), a_max=fac(1.0))
            e_nc, e_df, e_dfs = (float(np.sqrt(ind[k].sum())) for k in ("nc2", "df2", "dfstar2"))
            e_r = subdomain_eta_r(m, ind["res2"], ind["amin"])
            e_rs = subdomain_eta_r(m, ind["resstar2"], ind["amin"])
            ratio = MU / mu_hat
            eta = e_nc + e_r + max(np.sqrt(ratio), 1.0 / np.sqrt(ratio)) * e_df
            eta_star = e_nc + e_rs + e_dfs / np.sqrt(ratio)
            vals = dict(zip(COLS, (e_df, e_dfs, eta, eta_star, eta / energy, eta_star / energy)))
            for c in COLS:
                out.setdefault((mu_hat, c), []).append(vals[c])
    return out


def deviation(res, levels):
    """largest relative deviation from the goldens over every column, both rows, all levels"""
    worst = 0.0
    for mu_hat in (0.1, 1.0):
        for c in COLS:
            g = golden(OS14, c, "[4 4 1]", "0.1,0.1,%g" % mu_hat)
            for level in range(levels):
                worst = max(worst, abs(res[(mu_hat, c)][level] - g[level]) / abs(g[level]))
    return worst


def table(name, res, levels):
    lines = ["variant: %s   (largest relative deviation %.1f %%)" % (name, 100 * deviation(res, levels))]
    for mu_hat in (0.1, 1.0):
        lines.append("  (mu, mu_bar, mu_hat) = (0.1, 0.1, %g)" % mu_hat)
        for c in COLS:
            g = golden(OS14, c, "[4 4 1]", "0.1,0.1,%g" % mu_hat)
            lines.append("    %-18s here %s | reference %s" % (
                c, " ".join("%9.3e" % x for x in res[(mu_hat, c)]), " ".join("%9.2e" % x for x in g[:levels])))
        g_eta = golden(OS14, "eta_OS2014", "[4 4 1]", "0.1,0.1,%g" % mu_hat)
        g_eff = golden(OS14, "eff_OS2014_mu", "[4 4 1]", "0.1,0.1,%g" % mu_hat)
        lines.append("    %-18s here %s | implied   %s" % (
            "energy_mu", " ".join("%9.3e" % x for x in res["energy"]),
            " ".join("%9.2e" % (a / b) for a, b in list(zip(g_eta, g_eff))[:levels])))
    return "\n".join(lines)


def variants():
    base = dict(solve_factor="exact", est_factor="exact", flags=0, vol_order=-1, face_order=-1, est_order=3)
    yield "restatement (factor at the quadrature points, order 3)", base
    for vo, fo in itertools.product((0, 1, 2, 3, 5), (0, 1, 2, 3, 5)):
        if (vo, fo) != (3, 3):
            yield "solve: factor order %d in the volume term, %d on the faces" % (vo, fo), dict(base, vol_order=vo, face_order=fo)
    for eo in (0, 1, 2, 5):
        yield "estimator: factor order %d" % eo, dict(base, est_order=eo)
    yield "penalty a_f at the face midpoint", dict(base, flags=1)
    for kind in ("centre", "mean"):
        for est in ("exact", kind):
            for flags, what in ((0, "arithmetic mean, weights from K"), (2, "harmonic mean"), (4, "weights and gamma from a_f K")):
                yield "solve: factor P0 (%s), estimator: factor %s, %s" % (kind, "as written" if est == "exact" else "P0", what), \
                    dict(base, solve_factor=kind, est_factor=est, flags=flags)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--levels", type=int, default=4)
    ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden", "os2014_mu01_search.txt"))
    ap.add_argument("--quick", action="store_true", help="restatement only")
    ap.add_argument("--amplitudes", action="store_true",
                    help="also scan the amplitude and the direction of the problem's factor (is it the data or the grid's orientation, "
                         "not the arithmetic, that differs?)")
    args = ap.parse_args()
    results = []
    for name, v in variants():
        res = run_variant(v, args.levels)
        results.append((deviation(res, args.levels), name, res))
        print("%6.1f %%  %s" % (100 * results[-1][0], name), flush=True)
        if args.quick:
            break
    results.sort(key=lambda r: r[0])
    lines = ["# written by tools/os2014_mu01_search.py - largest relative deviation from the reference's mu = 0.1 goldens",
             "# (test/linearelliptic-block-swipdg-expectations_os2014_2daluconform.cxx:155-167, :185-197), %d levels" % args.levels, ""]
    lines += ["%6.1f %%  %s" % (100 * d, n) for d, n, _ in results]
    lines.append("")
    shown = {results[0][1], results[1][1] if len(results) > 1 else results[0][1]}
    for d, n, res in results:
        if n in shown or n.startswith("restatement"):
            lines += [table(n, res, args.levels), ""]
    if args.amplitudes:
        lines += ["# amplitude scan: a(mu) = 1 + A (1 - mu) sin(4 pi (x + y/2)) in solve, estimator and energy product (reference A = 0.75).",
                  "# No A fits all levels: the goldens' eta_DF asks for A = 0.78 on the coarsest and 0.765 on the finest level, the implied",
                  "# energy error for 0.855 ... 0.78 - and at A = 0.75 the reference has the larger eta_DF next to the smaller",
                  "# eta_NC + eta_R (level 3: 0.183 + 0.088 against 0.178 + 0.095 here), so it is not the amplitude of the data.", ""]
        for A in (0.65, 0.70, 0.75, 0.775, 0.80, 0.85):
            res = run_amplitude(A, args.levels)
            lines.append("A = %.3f  deviation %5.1f %%  energy %s  eta_DF(0.1) %s  eta(0.1) %s" % (
                A, 100 * deviation(res, args.levels), " ".join("%.3e" % x for x in res["energy"]),
                " ".join("%.3e" % x for x in res[(0.1, "eta_DF_OS2014")]), " ".join("%.3e" % x for x in res[(0.1, "eta_OS2014")])))
            print(lines[-1], flush=True)
        lines += ["", "# the eight images of the factor's direction (1, 1/2) under the symmetries of the square (equivalently: the grid",
                  "# mirrored or rotated against the data; the ESV2007 force and the bisection grids are invariant, so the mu = 1 rows",
                  "# could not tell): identical to three digits, the grid's orientation is not it either.", ""]
        import ctypes as C
        L = o.lib()
        L.or_set_os_direction.argtypes = [C.c_double, C.c_double]
        L.or_set_os_direction.restype = None
        base = dict(solve_factor="exact", est_factor="exact", flags=0, vol_order=-1, face_order=-1, est_order=3)
        try:
            for a, b in ((1, .5), (-1, .5), (1, -.5), (-1, -.5), (.5, 1), (-.5, 1), (.5, -1), (-.5, -1)):
                L.or_set_os_direction(a, b)
                res = run_variant(base, args.levels)
                lines.append("direction (%4g, %4g)  deviation %5.1f %%  energy %s  eta_DF(0.1) %s" % (
                    a, b, 100 * deviation(res, args.levels), " ".join("%.3e" % x for x in res["energy"]),
                    " ".join("%.3e" % x for x in res[(0.1, "eta_DF_OS2014")])))
                print(lines[-1], flush=True)
        finally:
            L.or_set_os_direction(1.0, 0.5)
        lines.append("")
    text = "\n".join(lines)
    if args.out:
        with open(args.out, "w") as f:
            f.write(text)
    print(text)
    return 0 if results[0][0] <= 0.006 else 1


if __name__ == "__main__":
    sys.exit(main())
