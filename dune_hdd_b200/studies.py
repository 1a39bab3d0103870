"""Convergence studies: the reference's ``SWIPDGStudy`` / ``BlockSWIPDGStudy`` (test/linearelliptic-swipdg.hh:56-330,
test/linearelliptic-block-swipdg.hh, base class test/linearelliptic.hh:60-260) - solve the test case on every level of
its grid ladder and tabulate the error norms, the estimators, the effectivities and the experimental orders of
convergence.  Everything a column needs runs on the device: assembly, solve, ``hdd_error_norms``, ``hdd_estimate``.

Difference to the reference, stated once: norms are taken against the analytic solution on the level grid itself
(the reference prolongs to the reference level; the difference is below the printed digits of its expectations,
SURVEY.md 9.2), so a test case without an exact solution has no norm / effectivity columns here."""
import math

from . import estimators, problems
from .discretizations import SWIPDG, BlockSWIPDG


class _StudyBase:
    disc_class = SWIPDG
    estimator_class = estimators.SWIPDG
    energy_estimators = ()  # estimators that get an effectivity column eff_* = eta / energy error

    def __init__(self, test_case, polorder=1, solver_options=None, device=0):
        self.test_case, self.polorder, self.device = test_case, polorder, device
        self.solver_options = solver_options or {"type": "cg.blockdiagonal", "precision": 1e-12, "max_iter": 200000}

    def available_norms(self):
        return ["L2", "H1_semi", "energy"] if self.test_case.provides_exact_solution() else []

    def _make(self, grid):
        return self.disc_class(grid, self.test_case.problem(), polorder=self.polorder, device=self.device)

    def _mu(self):
        p = self.test_case.parameters()
        return p.get("mu") if p else None

    def available_estimators(self, disc):
        ret = list(self.estimator_class.available(disc))
        if self.available_norms():
            ret += ["eff" + e[3:] for e in self.energy_estimators if e in ret]
        return ret

    def run(self, only_these_norms=None, only_these_estimators=None):
        """-> {"size": [...], "h": [...], "iterations": [...], "<norm>": [...], "<estimator>": [...], "eoc": {column: [...]}}"""
        table = {"size": [], "h": [], "iterations": []}
        prm = self.test_case.parameters() or None
        for level in range(self.test_case.num_refinements() + 1):
            grid = self.test_case.level_grid(level)
            disc = self._make(grid)
            disc.init()
            u, info = disc.solve(self.solver_options, mu=self._mu(), return_info=True)
            table["size"].append(grid.n_cells)
            table["h"].append(math.sqrt(4.0 / grid.n_cells))  # |Omega| = 4 for the ESV2007 / OS2014 cases
            table["iterations"].append(info["iterations"])
            norms = {}
            if self.available_norms():
                norms = disc.error_norms(*problems.ESV2007_EXACT, vector=u, order=5, mu=self._mu())
                for name in self.available_norms():
                    if only_these_norms is None or name in only_these_norms:
                        table.setdefault(name, []).append(norms[name])
            for est in self.available_estimators(disc):
                if only_these_estimators is not None and est not in only_these_estimators:
                    continue
                if est.startswith("eff"):
                    eta = self.estimator_class.estimate(disc, u, "eta" + est[3:], prm)
                    table.setdefault(est, []).append(eta / norms["energy"])
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
    energy_estimators = ("eta_ESV2007", "eta_ESV2007_alt")


class BlockSWIPDGStudy(_StudyBase):
    """test/linearelliptic-block-swipdg.hh: Discretizations::BlockSWIPDG + Estimators::BlockSWIPDG; the test case carries
    the partitioning and the parameters mu, mu_bar, mu_hat, parameter_range_min / _max"""
    disc_class = BlockSWIPDG
    estimator_class = estimators.BlockSWIPDG
    energy_estimators = ("eta_OS2014", "eta_OS2014_*")
