"""Convergence studies: the reference's ``SWIPDGStudy`` / ``BlockSWIPDGStudy`` (test/linearelliptic-swipdg.hh:56-330,
test/linearelliptic-block-swipdg.hh:60-400, base classes test/linearelliptic.hh:60-530) - solve the test case on every
level of its grid ladder and tabulate the error norms, the estimators, the effectivities and the experimental orders of
convergence.  Everything a column needs runs on the device: assembly, solve, prolongation (``hdd_prolong``), norms
(``hdd_error_norms`` / ``hdd_product_apply2``), ``hdd_estimate``.

As in the reference (test/linearelliptic.hh:146-214) the error is measured **on the reference level**: the level
solution is prolonged onto the reference grid and compared there with the exact solution or, for test cases without
one, with the discrete solution of the reference level in the L2 / H1-semi / elliptic product norms.
``norms_on="level"`` takes the norms against the analytic solution on the level grid itself instead (no reference-level
discretization; the difference is below the printed digits of the expectations)."""
import math

from . import estimators, problems
from .discretizations import SWIPDG, BlockSWIPDG


def localize_energy(blocks, diff, group, n_groups):
    """blocks [n, nl, nl], diff [n, nl], group [n] -> per group: sum_T d_T^T B_T d_T / (total * cells in the group)
    (the "average" loop of test/linearelliptic-swipdg.hh:209-211)"""
    import numpy as np
    local = np.einsum("ci,cij,cj->c", diff, blocks, diff)
    sums = np.bincount(group, weights=local, minlength=n_groups)
    counts = np.bincount(group, minlength=n_groups)
    return sums / (local.sum() * np.maximum(counts, 1))


class _StudyBase:
    disc_class = SWIPDG
    estimator_class = estimators.SWIPDG
    effectivity_ids = ()  # eta_<id> gets the column(s) eff_<id>[_<parameter>] = eta_<id> / energy error

    def __init__(self, test_case, polorder=1, solver_options=None, device=0, norms_on="reference"):
        if norms_on not in ("reference", "level"):
            raise ValueError("norms_on is 'reference' or 'level'")
        if norms_on == "level" and not test_case.provides_exact_solution():
            raise ValueError("a test case without exact solution needs the reference level")
        self.test_case, self.polorder, self.device, self.norms_on = test_case, polorder, device, norms_on
        self.solver_options = solver_options or {"type": "cg.blockdiagonal", "precision": 1e-12, "max_iter": 200000}
        self._reference = None

    # ---- what there is to compute ------------------------------------------------------------------------------
    def _parameter_ids(self):
        return sorted(self.test_case.parameters()) if self._parametric() else []

    def _parametric(self):
        return bool(self.test_case.parameters())

    def available_norms(self):
        """test/linearelliptic-swipdg.hh:260-265, test/linearelliptic-block-swipdg.hh:222-232"""
        norms = ["L2", "H1_semi"]
        if self._parametric():
            return norms + ["energy_" + k for k in self._parameter_ids()]
        return norms + ["energy"]

    def _energy_norms(self):
        return [n for n in self.available_norms() if n.startswith("energy")]

    def available_estimators(self, disc):
        """test/linearelliptic-block-swipdg.hh:283-298: the estimators plus one effectivity per energy norm"""
        ret = list(self.estimator_class.available(disc))
        for id in self.effectivity_ids:
            if "eta_" + id in ret:
                ret += ["eff_" + id + n[len("energy"):] for n in self._energy_norms()]
        return ret

    # ---- discretizations ---------------------------------------------------------------------------------------
    def _make(self, grid, products=()):
        return self.disc_class(grid, self.test_case.problem(), polorder=self.polorder, device=self.device,
                               only_these_products=products)

    def _mu(self):
        p = self.test_case.parameters()
        return p.get("mu") if p else None

    def reference(self):
        """compute_reference_solution() (test/linearelliptic.hh:226-247): (grid, discretization, solution or None)"""
        if self._reference is None:
            exact = self.test_case.provides_exact_solution()
            grid = self.test_case.reference_grid()
            disc = self._make(grid, () if exact else ("l2", "h1_semi", "elliptic"))
            disc.init()
            u = None if exact else disc.solve(self.solver_options, mu=self._mu())
            self._reference = (grid, disc, u)
        return self._reference

    def error_norms(self, disc, u):
        """current_error_norm(type) for every norm (test/linearelliptic.hh:188-216) -> {name: value}"""
        names = self.available_norms()
        if self.norms_on == "level":
            n = disc.error_norms(*problems.ESV2007_EXACT, vector=u, order=5, mu=self._mu())
            return {k: n[k] for k in names}
        _, ref, u_ref = self.reference()
        prolonged = ref.prolong(disc, u)
        if self.test_case.provides_exact_solution():
            n = ref.error_norms(*problems.ESV2007_EXACT, vector=prolonged, order=5, mu=self._mu())
            return {k: n["energy" if k.startswith("energy") else k] for k in names}
        diff = u_ref - prolonged
        out = {}
        for k in names:
            if k == "L2":
                out[k] = ref.get_product("l2").induced_norm(diff)
            elif k == "H1_semi":
                out[k] = ref.get_product("h1_semi").induced_norm(diff)
            else:  # Products::Elliptic of problem.with_mu(parameters[id]) (test/linearelliptic-block-swipdg.hh:252-270)
                mu = self.test_case.parameters()[k[len("energy_"):]] if k != "energy" else None
                out[k] = ref.get_product("elliptic").induced_norm(diff, mu=mu)
        return out

    # ---- localization (compute_reference_indicators / compute_indicators) ----------------------------------------
    def _groups(self, disc, father):
        """the coarse entity every reference-level cell is accounted to: its father cell"""
        return father, disc.grid.n_cells

    def reference_indicators(self, disc, u):
        """compute_reference_indicators() (test/linearelliptic-swipdg.hh:133-223, test/linearelliptic-block-swipdg.hh:122-199):
        the elliptic energy of (reference solution - prolonged solution) per reference-level cell, summed over the cells
        of each coarse entity and divided by (total * number of fine cells of the entity).  Like the reference this
        exists for test cases without an exact solution only.  The per-cell energies come from the device-assembled
        "elliptic" product blocks (volume pattern: one dense block per cell)."""
        if self.test_case.provides_exact_solution():
            raise NotImplementedError("you_have_to_implement_this (test/linearelliptic-swipdg.hh:157-158)")
        from . import grids
        grid_r, ref, u_ref = self.reference()
        father = grids.fathers(disc.grid, grid_r)
        diff = (u_ref - ref.prolong(disc, u, father)).reshape(grid_r.n_cells, ref.n_loc)
        blocks = ref.get_product("elliptic").freeze_parameter(self._mu()).reshape(grid_r.n_cells, ref.n_loc, ref.n_loc)
        group, n_groups = self._groups(disc, father)
        return localize_energy(blocks, diff, group, n_groups)

    def indicators(self, disc, u, type):
        """compute_indicators(type): estimate_local of the estimator class"""
        return self.estimator_class.estimate_local(disc, u, type, self.test_case.parameters() or None)

    # ---- the study ---------------------------------------------------------------------------------------------
    def run(self, only_these_norms=None, only_these_estimators=None, levels=None):
        """-> {"size": [...], "h": [...], "iterations": [...], "<norm>": [...], "<estimator>": [...], "eoc": {column: [...]}}
        levels: the refinements to compute (default: all of the test case's ladder)"""
        table = {"size": [], "h": [], "iterations": []}
        prm = self.test_case.parameters() or None
        levels = range(self.test_case.num_refinements() + 1) if levels is None else levels
        for level in levels:
            grid = self.test_case.level_grid(level)
            disc = self._make(grid)
            disc.init()
            u, info = disc.solve(self.solver_options, mu=self._mu(), return_info=True)
            table["size"].append(grid.n_cells)
            table["h"].append(math.sqrt(4.0 / grid.n_cells))  # |Omega| = 4 for the ESV2007 / OS2014 cases
            table["iterations"].append(info["iterations"])
            norms = self.error_norms(disc, u)
            for name, value in norms.items():
                if only_these_norms is None or name in only_these_norms:
                    table.setdefault(name, []).append(value)
            for est in self.available_estimators(disc):
                if only_these_estimators is not None and est not in only_these_estimators:
                    continue
                if est.startswith("eff_"):
                    id = next(i for i in sorted(self.effectivity_ids, key=len, reverse=True) if est[4:].startswith(i))
                    eta = self.estimator_class.estimate(disc, u, "eta_" + id, prm)
                    table.setdefault(est, []).append(eta / norms["energy" + est[4 + len(id):]])
                else:
                    table.setdefault(est, []).append(self.estimator_class.estimate(disc, u, est, prm))
            del disc
        table["eoc"] = {}
        for name, col in table.items():
            if name in ("size", "h", "iterations", "eoc") or name.startswith("eff"):
                continue
            table["eoc"][name] = [math.log(col[l] / col[l + 1]) / math.log(table["h"][l] / table["h"][l + 1])
                                  if col[l] > 0 and col[l + 1] > 0 else float("nan") for l in range(len(col) - 1)]
        return table


class SWIPDGStudy(_StudyBase):
    """test/linearelliptic-swipdg.hh: Discretizations::SWIPDG + Estimators::SWIPDG on the ladder of the test case"""
    effectivity_ids = ("ESV2007", "ESV2007_alt")


class BlockSWIPDGStudy(_StudyBase):
    """test/linearelliptic-block-swipdg.hh: Discretizations::BlockSWIPDG + Estimators::BlockSWIPDG; the test case carries
    the partitioning and the parameters mu, mu_bar, mu_hat, parameter_range_min / _max.  Effectivity ids as
    test/linearelliptic-block-swipdg.hh:276-279 (the longer id first where one contains the other)."""
    disc_class = BlockSWIPDG
    estimator_class = estimators.BlockSWIPDG
    effectivity_ids = ("OS2014_*", "OS2014")

    def _groups(self, disc, father):
        """... its father's subdomain (ms_grid->subdomainOf(father_entity), test/linearelliptic-block-swipdg.hh:178)"""
        return disc.grid.cell_subdomain[father], disc.grid.n_subdomains
