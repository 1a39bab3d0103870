"""Problem data: mirrors ``LinearElliptic::ProblemInterface`` (problems/interfaces.hh:84-144) - diffusion_factor,
diffusion_tensor, force, dirichlet, neumann as affinely decomposed functions - for the problems the BASELINE
configs name: ESV2007 (problems/ESV2007.hh), OS2014::ParametricESV2007 (problems/OS2014.hh) and Spe10::Model1
(problems/spe10.hh, with a synthetic permeability field because perm_case1.dat is not shipped).
"""
import numpy as np

from . import capi


class Function:
    """A localised scalar data function (hdd_function)."""

    def __init__(self, kind, order=0, value=0.0, cell_values=None, expression=None, name=""):
        self.kind, self.order, self.value, self.name = kind, int(order), float(value), name
        self.cell_values = None if cell_values is None else capi.as_f64(cell_values)
        self.expression = expression

    def to_c(self):
        f = capi.hdd_function()
        f.kind, f.order, f.value = self.kind, self.order, self.value
        f.cell_values = capi.ptr(self.cell_values)
        f.expression = None if self.expression is None else self.expression.encode()
        return f


def Constant(value, name=""):
    """Stuff::Functions::Constant"""
    return Function(capi.HDD_FN_CONSTANT, 0, value=value, name=name)


def Cellwise(values, name=""):
    """piecewise constant data (Spe10::Model1, Indicator, Checkerboard) evaluated per cell by the host"""
    return Function(capi.HDD_FN_CELLWISE, 0, cell_values=values, name=name)


def Expression(expression, order, name=""):
    """Stuff::Functions::Expression("x", expression, order)"""
    return Function(capi.HDD_FN_EXPRESSION, order, expression=expression, name=name)


class AffinelyDecomposable:
    """Pymor::Functions::AffinelyDecomposableDefault: sum_q theta_q(mu) component_q + affine_part."""

    def __init__(self, affine_part=None, components=(), coefficients=()):
        self.affine = affine_part
        self.components = list(components)
        self.coefficients = list(coefficients)
        assert len(self.components) == len(self.coefficients)

    def parametric(self):
        return len(self.components) > 0

    def has_affine_part(self):
        return self.affine is not None

    def num_components(self):
        return len(self.components)

    def to_c(self, keep):
        a = capi.hdd_affine_function()
        a.n_components = len(self.components)
        if self.components:
            structs = [c.to_c() for c in self.components]
            encoded = [c.encode() for c in self.coefficients]
            arr = (capi.hdd_function * len(structs))(*structs)
            coefs = (capi.C.c_char_p * len(encoded))(*encoded)
            keep += [structs, encoded, arr, coefs]
            a.components = arr
            a.coefficients = coefs
        if self.affine is not None:
            f = self.affine.to_c()
            keep.append(f)
            a.affine_part = capi.C.pointer(f)
        return a


class Problem:
    """ProblemInterface: diffusion_factor(), diffusion_tensor(), force(), dirichlet(), neumann(), with_mu()."""

    def __init__(self, diffusion_factor, force, dirichlet=None, neumann=None, diffusion_tensor=None,
                 parameter_name=None, parameter_size=0, name="problem"):
        self.diffusion_factor = diffusion_factor
        self.force = force
        self.dirichlet = dirichlet or AffinelyDecomposable(Constant(0.0, "dirichlet"))
        self.neumann = neumann or AffinelyDecomposable(Constant(0.0, "neumann"))
        self.diffusion_tensor = None if diffusion_tensor is None else capi.as_f64(diffusion_tensor)
        self.parameter_name = parameter_name
        self.parameter_size = parameter_size
        self.name = name

    def parametric(self):
        return any(f.parametric() for f in (self.diffusion_factor, self.force, self.dirichlet, self.neumann))

    def parameter_type(self):
        return {self.parameter_name: self.parameter_size} if self.parametric() else {}

    def to_c(self):
        keep = []
        p = capi.hdd_problem()
        p.diffusion_factor = self.diffusion_factor.to_c(keep)
        p.diffusion_tensor = capi.ptr(self.diffusion_tensor)
        p.force = self.force.to_c(keep)
        p.dirichlet = self.dirichlet.to_c(keep)
        p.neumann = self.neumann.to_c(keep)
        p.parameter_name = None if self.parameter_name is None else self.parameter_name.encode()
        p.parameter_size = self.parameter_size
        p._keep = keep
        return p


ESV2007_FORCE = "0.5*pi*pi*cos(0.5*pi*x[0])*cos(0.5*pi*x[1])"


def ESV2007(integration_order=3):
    """problems/ESV2007.hh:75-81: a = 1, K = I, f = 1/2 pi^2 cos(pi x/2) cos(pi y/2), g_D = g_N = 0."""
    return Problem(AffinelyDecomposable(Constant(1.0, "diffusion_factor")),
                   AffinelyDecomposable(Expression(ESV2007_FORCE, integration_order, "force")), name="ESV2007")


# Stuff::Functions::ESV2007::Testcase1ExactSolution and its gradient as Expression strings (u, du/dx, du/dy)
ESV2007_EXACT = ("cos(0.5*pi*x[0])*cos(0.5*pi*x[1])", "-0.5*pi*sin(0.5*pi*x[0])*cos(0.5*pi*x[1])",
                 "-0.5*pi*cos(0.5*pi*x[0])*sin(0.5*pi*x[1])")


def esv2007_exact(xy):
    """Stuff::Functions::ESV2007::Testcase1ExactSolution (testcases/ESV2007.hh:41)"""
    return np.cos(0.5 * np.pi * xy[..., 0]) * np.cos(0.5 * np.pi * xy[..., 1])


def OS2014ParametricESV2007(integration_order=3):
    """problems/OS2014.hh:63-113: a(mu) = [1 + 0.75 sin(4 pi (x + y/2))] + mu [-0.75 sin(4 pi (x + y/2))]."""
    factor = AffinelyDecomposable(
        Expression("1+0.75*(sin(4*pi*(x[0]+0.5*x[1])))", integration_order, "affine_part"),
        [Expression("-0.75*(sin(4*pi*(x[0]+0.5*x[1])))", integration_order, "component_0")], ["mu"])
    return Problem(factor, AffinelyDecomposable(Expression(ESV2007_FORCE, integration_order, "force")),
                   parameter_name="mu", parameter_size=1, name="OS2014.parametricESV2007")


SPE10_MIN, SPE10_MAX = 1e-3, 998.915  # Spe10::Model1 min_value / max_value (problems/spe10.hh:154-155)


def spe10_synthetic_permeability(nx=100, ny=20, seed=20141010):
    """Synthetic log-normal stand-in for perm_case1.dat on the 100 x 20 SPE10 model-1 layer."""
    from scipy.ndimage import gaussian_filter
    rng = np.random.default_rng(seed)
    g = gaussian_filter(rng.standard_normal((ny, nx)), sigma=(1.0, 4.0))
    g = (g - g.mean()) / g.std()
    return np.clip(10.0 ** (1.5 * g), SPE10_MIN, SPE10_MAX)


def Spe10Model1(grid, permeability=None, lower_left=(0.0, 0.0), upper_right=(5.0, 1.0), channel=None,
                parametric=False):
    """problems/spe10.hh:72-185: K = k(x) I on the 100 x 20 layer grid, f = three indicator boxes
    (+2000 on [0.95,1.10]x[0.30,0.45], -1000 on [3.00,3.15]x[0.75,0.90], -1000 on [4.25,4.40]x[0.25,0.40]),
    a = 1 (+ mu-dependent channel term "-1.0*mu" * 0.9 * channel if parametric, problems/spe10.hh:141-179)."""
    k = spe10_synthetic_permeability() if permeability is None else np.asarray(permeability)
    ny, nx = k.shape
    c = grid.centers()
    ix = np.clip(((c[:, 0] - lower_left[0]) / (upper_right[0] - lower_left[0]) * nx).astype(int), 0, nx - 1)
    iy = np.clip(((c[:, 1] - lower_left[1]) / (upper_right[1] - lower_left[1]) * ny).astype(int), 0, ny - 1)
    kc = k[iy, ix]
    tensor = np.zeros((grid.n_cells, 4))
    tensor[:, 0] = kc
    tensor[:, 3] = kc
    f = np.zeros(grid.n_cells)
    for (x0, x1, y0, y1, v) in ((0.95, 1.10, 0.30, 0.45, 2000.0), (3.00, 3.15, 0.75, 0.90, -1000.0),
                                (4.25, 4.40, 0.25, 0.40, -1000.0)):
        f[(c[:, 0] >= x0) & (c[:, 0] <= x1) & (c[:, 1] >= y0) & (c[:, 1] <= y1)] += v
    if parametric:
        ch = np.zeros(grid.n_cells) if channel is None else np.asarray(channel, dtype=float)
        factor = AffinelyDecomposable(Constant(1.0, "affine_part"), [Cellwise(0.9 * ch, "component_0")], ["-1.0*mu"])
        return Problem(factor, AffinelyDecomposable(Cellwise(f, "force")), diffusion_tensor=tensor,
                       parameter_name="mu", parameter_size=1, name="Spe10.Model1.parametric")
    return Problem(AffinelyDecomposable(Constant(1.0, "diffusion_factor")), AffinelyDecomposable(Cellwise(f, "force")),
                   diffusion_tensor=tensor, name="Spe10.Model1")


def read_spe10_model1(filename, nx=100, ny=20):
    """The permeability file of SPE10 model 1 (``perm_case1.dat``, looked up outside the reference tree,
    CMakeLists.txt:76-81): nx * ny whitespace-separated values, x running fastest (upstream
    Stuff::Functions::Spe10::Model1, model1_x_elements = 100, model1_z_elements = 20; recalled, the file is not shipped).
    -> array [ny, nx] for Spe10Model1(permeability=...)."""
    with open(filename) as f:
        values = np.array(f.read().split(), dtype=np.float64)
    if values.size != nx * ny:
        raise ValueError("%s holds %d values, expected %d x %d" % (filename, values.size, nx, ny))
    return values.reshape(ny, nx)


def Thermalblock(grid, num_elements=(2, 2), lower_left=(0.0, 0.0), upper_right=(1.0, 1.0),
                 parameter_name="diffusion_factor", force=1.0, dirichlet=0.0, neumann=0.0):
    """problems/thermalblock.hh:44-128: the diffusion factor is a Pymor::Functions::Checkerboard - one indicator
    function per block of a num_elements[0] x num_elements[1] checkerboard (x fastest), weighted by the parameter
    component ``parameter_name[k]``; K = I, constant force / Dirichlet / Neumann data.  A handle carries at most
    7 parametric components (HDD_ERR_NOT_IMPLEMENTED beyond)."""
    nx, ny = int(num_elements[0]), int(num_elements[1])
    c = grid.centers()
    ix = np.clip(((c[:, 0] - lower_left[0]) / (upper_right[0] - lower_left[0]) * nx).astype(int), 0, nx - 1)
    iy = np.clip(((c[:, 1] - lower_left[1]) / (upper_right[1] - lower_left[1]) * ny).astype(int), 0, ny - 1)
    block = iy * nx + ix
    comps = [Cellwise((block == k).astype(np.float64), "component_%d" % k) for k in range(nx * ny)]
    factor = AffinelyDecomposable(None, comps, ["%s[%d]" % (parameter_name, k) for k in range(nx * ny)])
    return Problem(factor, AffinelyDecomposable(Constant(force, "force")),
                   AffinelyDecomposable(Constant(dirichlet, "dirichlet")), AffinelyDecomposable(Constant(neumann, "neumann")),
                   parameter_name=parameter_name, parameter_size=nx * ny, name="thermalblock")
