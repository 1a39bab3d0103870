"""Test cases = problem + grid ladder + parameters (testcases/{base,ESV2007,OS2014,spe10}.hh)."""
from . import grids, problems


class ESV2007:
    """testcases/ESV2007.hh: [-1,1]^2, AllDirichlet, exact solution cos(pi x/2) cos(pi y/2).
    ALU simplex ladder 128/512/2048/8192 (reference 32768); SGrid ladder 64/256/1024/4096 (reference 16384)."""

    def __init__(self, grid_type="alu", num_refinements=3, partitions=(1, 1)):
        self.grid_type, self.num_refinements_, self.partitions = grid_type, num_refinements, partitions
        self.problem_ = problems.ESV2007(3)

    def num_refinements(self):
        return self.num_refinements_

    def level_grid(self, refinement):
        if self.grid_type == "alu":
            return grids.simplex(4 * 2 ** refinement, partitions=self.partitions)
        return grids.cube(8 * 2 ** refinement, partitions=self.partitions)

    def reference_grid(self):
        return self.level_grid(self.num_refinements_ + 1)

    def problem(self):
        return self.problem_

    def provides_exact_solution(self):
        return True

    exact_solution = staticmethod(problems.esv2007_exact)

    def parameters(self):
        return {}


class ESV2007Multiscale(ESV2007):
    """testcases/ESV2007.hh:141-172 with num_partitions "[px py 1]"."""

    def __init__(self, num_partitions=(1, 1), num_refinements=3, grid_type="alu"):
        super().__init__(grid_type, num_refinements, tuple(num_partitions)[:2])

    def partitioning(self):
        return "[%d %d 1]" % self.partitions


class OS2014ParametricESV2007Multiscale(ESV2007Multiscale):
    """testcases/OS2014.hh: the parametric problem on the ESV2007 multiscale grids; parameter range [0.1, 1]."""

    def __init__(self, parameters, num_partitions=(1, 1), num_refinements=3):
        super().__init__(num_partitions, num_refinements, "alu")
        self.problem_ = problems.OS2014ParametricESV2007(3)
        self.parameters_ = dict(parameters)
        self.parameters_.setdefault("parameter_range_min", 0.1)
        self.parameters_.setdefault("parameter_range_max", 1.0)

    def provides_exact_solution(self):
        return False

    def parameters(self):
        return self.parameters_


class Spe10Model1:
    """testcases/spe10.hh:262-311: [0,5]x[0,1], 100x20 cells (SGrid) ladder 2000 -> 8000, reference 32000."""

    def __init__(self, num_refinements=1, permeability=None):
        self.num_refinements_ = num_refinements
        self.permeability = problems.spe10_synthetic_permeability() if permeability is None else permeability

    def level_grid(self, refinement):
        f = 2 ** refinement
        return grids.cube(100 * f, 20 * f, (0.0, 0.0), (5.0, 1.0))

    def problem(self, grid):
        return problems.Spe10Model1(grid, self.permeability)
