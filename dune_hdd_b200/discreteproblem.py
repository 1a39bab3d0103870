"""The configuration-file driver around the discretizations: ``DiscreteProblem`` / ``DiscreteBlockProblem``
(discreteproblem.hh:44-214, :220-440) and the example classes built on them (examples/linearelliptic/swipdg.hh:24-97,
block-swipdg.hh), SURVEY.md 8f rank 4.  Host code only: it reads ``<id>.cfg``, creates the grid arrays, the boundary
info and the problem, and hands them to ``SWIPDG`` / ``BlockSWIPDG``; nothing here touches the device.

``Configuration`` is the subset of ``Stuff::Common::Configuration`` (a Dune::ParameterTree: ``[section]`` headers,
``key = value`` lines, nested keys joined by dots, vectors written ``[a b c]``) that the driver uses.  Provider ids and
the default values of the upstream providers (``Stuff::GridProviders``, ``Stuff::Grid::BoundaryInfoProvider``,
``grid::Multiscale::MsGridProviders``) are recalled from dune-stuff / dune-grid-multiscale, which are not vendored; the
ids of the problems are the reference's own (problems/*.hh ``static_id()``)."""
import os
import time

import numpy as np

from . import grids, problems


class configuration_error(RuntimeError):
    """Stuff::Exceptions::configuration_error (discreteproblem.hh:99-101)"""


def _format(value):
    if isinstance(value, bool):
        return "true" if value else "false"
    if isinstance(value, (list, tuple, np.ndarray)):
        return "[" + " ".join(_format(v) for v in value) + "]"
    return str(value)


class Configuration:
    """flat ``{"a.b.c": "value"}`` view of a parameter tree"""

    def __init__(self, keys=None, values=None):
        self._d = {}
        if isinstance(keys, dict):
            for k, v in keys.items():
                self[k] = v
        elif isinstance(keys, str):
            self[keys] = values
        elif keys is not None:
            for k, v in zip(keys, values):
                self[k] = v

    # ---- reading ---------------------------------------------------------------------------------------------
    @classmethod
    def from_string(cls, text):
        cfg, prefix = cls(), ""
        for raw in text.splitlines():
            line = raw.split("#", 1)[0].strip()
            if not line:
                continue
            if line.startswith("[") and line.endswith("]") and "=" not in line:
                prefix = line[1:-1].strip()
                prefix = prefix + "." if prefix else ""
                continue
            if "=" not in line:
                # continuation of a multi-line list of keys (write_keys_to_file, discreteproblem.hh:190-196): alternatives
                # the user is meant to delete; the first entry stays the value
                continue
            key, value = line.split("=", 1)
            cfg._d[prefix + key.strip()] = value.strip()
        return cfg

    @classmethod
    def from_file(cls, filename):
        with open(filename) as f:
            return cls.from_string(f.read())

    def __setitem__(self, key, value):
        self._d[key] = _format(value)

    def __getitem__(self, key):
        if key not in self._d:
            raise configuration_error("missing key '%s' in the following Configuration:\n\n%s" % (key, self.report()))
        return self._d[key]

    def __contains__(self, key):
        return key in self._d

    def __len__(self):
        return len(self._d)

    def empty(self):
        return not self._d

    def has_key(self, key):
        return key in self._d

    def has_sub(self, name):
        return any(k.startswith(name + ".") for k in self._d)

    def sub(self, name):
        if not self.has_sub(name):
            raise configuration_error("missing sub '%s' in the following Configuration:\n\n%s" % (name, self.report()))
        out = Configuration()
        out._d = {k[len(name) + 1:]: v for k, v in self._d.items() if k.startswith(name + ".")}
        return out

    def add(self, other, sub_name=""):
        for k, v in other._d.items():
            self._d[(sub_name + "." if sub_name else "") + k] = v

    def get_value_keys(self):
        return [k for k in self._d if "." not in k]

    def get(self, key, default=None, type=None):
        """get< T >(key[, default]); type: str, int, float, bool or "vector" (list of floats).  Without ``type`` the type
        of ``default`` decides."""
        if key not in self._d:
            if default is None:
                self[key]  # raises
            return default
        raw = self._d[key]
        if type is None:
            type = "vector" if isinstance(default, (list, tuple, np.ndarray)) else (default.__class__ if default is not None else str)
        if type is bool:
            if raw.lower() in ("true", "1", "yes", "on"):
                return True
            if raw.lower() in ("false", "0", "no", "off"):
                return False
            raise configuration_error("'%s = %s' is not a bool" % (key, raw))
        if type == "vector":
            body = raw.strip()
            if body.startswith("[") and body.endswith("]"):
                body = body[1:-1]
            return [float(t) for t in body.replace(";", " ").replace(",", " ").split()]
        try:
            return type(raw)
        except ValueError:
            raise configuration_error("'%s = %s' is not a %s" % (key, raw, type.__name__))

    # ---- writing ---------------------------------------------------------------------------------------------
    def report(self):
        """sections in order of first appearance, top-level keys first"""
        sections = {}
        for k, v in self._d.items():
            sec, _, name = k.rpartition(".")
            sections.setdefault(sec, []).append((name, v))
        out = []
        for sec in [""] * ("" in sections) + [s for s in sections if s]:
            if sec:
                out.append("[%s]" % sec)
            out += ["%s = %s" % kv for kv in sections[sec]]
        return "\n".join(out) + "\n"

    __str__ = report


# ---- providers -------------------------------------------------------------------------------------------------
class GridProviders:
    """Stuff::GridProviders< GridType >: the cube provider (the only one of the upstream set that does not read a file)"""

    @staticmethod
    def available():
        return ["stuff.grid.provider.cube"]

    @staticmethod
    def default_config(type, sub_name=""):
        GridProviders._check(type)
        cfg = Configuration()
        cfg.add(Configuration({"lower_left": [0.0, 0.0], "upper_right": [1.0, 1.0], "num_elements": [8, 8],
                               "num_refinements": 0}), sub_name)
        return cfg

    @classmethod
    def _check(cls, type):
        if type not in cls.available():
            raise configuration_error("'%s' is not one of the available grid providers: %s" % (type, ", ".join(cls.available())))

    @classmethod
    def create(cls, type, config, grid_type="sgrid", partitions=(1, 1)):
        """-> grids.Grid.  grid_type "sgrid": SGrid<2,2> (num_elements cells, each refinement halves them); "alu":
        ALUGrid<2,2,simplex,conforming> - two triangles per element, every refinement is two bisections
        (refineStepsForHalf, testcases/base.hh:96-98); the closed-form generator needs at least one."""
        cls._check(type)
        cfg = config.sub(type) if config.has_sub(type) else config
        d = cls.default_config(type)
        ll = cfg.get("lower_left", d.get("lower_left", type="vector"), "vector")
        ur = cfg.get("upper_right", d.get("upper_right", type="vector"), "vector")
        ne = [int(v) for v in cfg.get("num_elements", d.get("num_elements", type="vector"), "vector")]
        refine = cfg.get("num_refinements", 0, int)
        if len(ne) == 1:
            ne = ne * 2
        if refine < 0 or min(ne[:2]) < 1:
            raise configuration_error("num_elements / num_refinements have to be positive")
        if grid_type == "sgrid":
            f = 2 ** refine
            return grids.cube(ne[0] * f, ne[1] * f, tuple(ll[:2]), tuple(ur[:2]), partitions=partitions)
        if grid_type != "alu":
            raise configuration_error("grid_type is 'sgrid' or 'alu'")
        if ne[0] != ne[1] or refine < 1:
            raise NotImplementedError("the simplex ladder needs num_elements = [n n] and num_refinements >= 1")
        return grids.simplex(ne[0] * 2 ** (refine - 1), tuple(ll[:2]), tuple(ur[:2]), partitions=partitions)


class MsGridProviders(GridProviders):
    """grid::Multiscale::MsGridProviders< GridType >: the cube provider with ``num_partitions``"""

    @staticmethod
    def available():
        return ["grid.multiscale.provider.cube"]

    @staticmethod
    def default_config(type, sub_name=""):
        MsGridProviders._check(type)
        cfg = Configuration()
        cfg.add(Configuration({"lower_left": [0.0, 0.0], "upper_right": [1.0, 1.0], "num_elements": [8, 8],
                               "num_refinements": 0, "num_partitions": [2, 2], "oversampling_layers": 0}), sub_name)
        return cfg

    @classmethod
    def create(cls, type, config, grid_type="sgrid"):
        cls._check(type)
        cfg = config.sub(type) if config.has_sub(type) else config
        parts = [int(v) for v in cfg.get("num_partitions", [2, 2], "vector")]
        if len(parts) == 1:
            parts = parts * 2
        return super().create(type, cfg, grid_type, partitions=tuple(parts[:2]))


class BoundaryInfoProvider:
    """Stuff::Grid::BoundaryInfoProvider: AllDirichlet / AllNeumann; -> per-face types for hdd_mesh_create (None = all
    Dirichlet)"""

    @staticmethod
    def available():
        return ["stuff.grid.boundaryinfo.alldirichlet", "stuff.grid.boundaryinfo.allneumann"]

    @staticmethod
    def default_config(type, sub_name=""):
        return Configuration()  # neither has settings of its own

    @staticmethod
    def create(config, grid):
        type = config.get("type", "stuff.grid.boundaryinfo.alldirichlet")
        if type == "stuff.grid.boundaryinfo.alldirichlet":
            return None
        if type == "stuff.grid.boundaryinfo.allneumann":
            return np.full(grid.cell_neigh.shape, 2, dtype=np.uint8)
        raise configuration_error("'%s' is not one of the available boundary infos: %s"
                                  % (type, ", ".join(BoundaryInfoProvider.available())))


class ProblemsProvider:
    """LinearElliptic::ProblemsProvider (problems.hh:47-220) for the problems of this path"""
    ESV2007 = "hdd.linearelliptic.problem.ESV2007"                   # problems/ESV2007.hh:50-53
    OS2014 = "hdd.linearelliptic.problem.OS2014.parametricESV2007"   # problems/OS2014.hh:81-84
    SPE10 = "hdd.linearelliptic.problem.spe10.model1"                # problems/spe10.hh:67-70
    THERMALBLOCK = "hdd.linearelliptic.problem.thermalblock"         # problems/thermalblock.hh:59-62

    @classmethod
    def available(cls):
        return [cls.ESV2007, cls.OS2014, cls.SPE10, cls.THERMALBLOCK]

    @classmethod
    def default_config(cls, type, sub_name=""):
        if type in (cls.ESV2007, cls.OS2014):
            d = {"integration_order": 3}  # problems/ESV2007.hh:55-65, problems/OS2014.hh:86-96
        elif type == cls.SPE10:  # problems/spe10.hh:72-93
            d = {"filename": "perm_case1.dat", "lower_left": [0.0, 0.0], "upper_right": [5.0, 1.0],
                 "parametric_channel": False}
        elif type == cls.THERMALBLOCK:  # problems/thermalblock.hh:64-97; [4 4] there, more parts than one handle carries
            d = {"diffusion_factor.lower_left": [0.0, 0.0], "diffusion_factor.upper_right": [1.0, 1.0],
                 "diffusion_factor.num_elements": [2, 2], "diffusion_factor.parameter_name": "diffusion_factor",
                 "diffusion_factor.name": "diffusion_factor", "force.value": 1, "force.name": "force",
                 "dirichlet.value": 0, "dirichlet.name": "dirichlet", "neumann.value": 0, "neumann.name": "neumann"}
        else:
            raise configuration_error("'%s' is not one of the available problems: %s" % (type, ", ".join(cls.available())))
        cfg = Configuration()
        cfg.add(Configuration(d), sub_name)
        return cfg

    @classmethod
    def create(cls, type, config, grid):
        """the grid is needed by the problems whose data are piecewise constant (localised per cell by the host)"""
        d = cls.default_config(type)
        cfg = config.sub(type) if config.has_sub(type) else config
        if type == cls.ESV2007:
            return problems.ESV2007(cfg.get("integration_order", 3, int))
        if type == cls.OS2014:
            return problems.OS2014ParametricESV2007(cfg.get("integration_order", 3, int))
        if type == cls.SPE10:
            filename = cfg.get("filename", d["filename"])
            perm = problems.read_spe10_model1(filename) if os.path.exists(filename) else None
            return problems.Spe10Model1(grid, perm, tuple(cfg.get("lower_left", [0.0, 0.0], "vector")),
                                        tuple(cfg.get("upper_right", [5.0, 1.0], "vector")),
                                        parametric=cfg.get("parametric_channel", False, bool))
        f = cfg.sub("diffusion_factor") if cfg.has_sub("diffusion_factor") else Configuration()
        value = lambda name, default: (cfg.sub(name) if cfg.has_sub(name) else Configuration()).get("value", default, float)
        return problems.Thermalblock(grid, [int(v) for v in f.get("num_elements", [2, 2], "vector")],
                                     tuple(f.get("lower_left", [0.0, 0.0], "vector")),
                                     tuple(f.get("upper_right", [1.0, 1.0], "vector")),
                                     f.get("parameter_name", "diffusion_factor"), force=value("force", 1.0),
                                     dirichlet=value("dirichlet", 0.0), neumann=value("neumann", 0.0))


# ---- discrete problems -----------------------------------------------------------------------------------------
def _write_keys(name, keys):
    pad = " " * len(name + " = ")
    return [name + " = " + keys[0]] + [pad + k for k in keys[1:]]


def write_config(filename, id, block=False):
    """DiscreteProblem::write_config (discreteproblem.hh:63-84) / DiscreteBlockProblem::write_config (:241-276): the first
    of the listed alternatives is the one in effect"""
    grid_providers = MsGridProviders if block else GridProviders
    lines = ["[%s]" % id] + _write_keys("gridprovider", grid_providers.available())
    if not block:
        lines += _write_keys("boundaryinfo", BoundaryInfoProvider.available())
    lines += _write_keys("problem", ProblemsProvider.available())
    lines += ["[logging]", "info  = true", "debug = %s" % ("false" if block else "true"), "file  = false", "visualize = true"]
    if not block:
        lines += ["[parameter]", "0.diffusion_factor = [0.1 0.1 1.0 1.0]", "1.diffusion_factor = [1.0 1.0 0.1 0.1]"]
    text = "\n".join(lines) + "\n"
    for provider in (grid_providers, ProblemsProvider):
        for type in provider.available():
            cfg = provider.default_config(type, type)
            if not cfg.empty():
                text += cfg.report()
    with open(filename, "w") as f:
        f.write(text)


class DiscreteProblem:
    """DiscreteProblem(id, arguments) (discreteproblem.hh:86-152).  arguments[0], if it is a directory, is where
    ``<id>.cfg`` is looked up (the Python examples pass os.getcwd(), examples/linearelliptic/cg_main.py:20)."""
    block = False

    def __init__(self, id, arguments=(), grid_type="sgrid", out=None):
        self._info = out if out is not None else (lambda s: None)
        directory = arguments[0] if arguments and os.path.isdir(arguments[0]) else "."
        self.config_ = Configuration.from_file(os.path.join(directory, id + ".cfg"))
        if not self.config_.has_sub(id):
            raise configuration_error("Missing sub '%s' in the following Configuration:\n\n%s" % (id, self.config_))
        self.filename_ = self.config_.get(id + ".filename", id)
        logging = self.config_.sub("logging") if self.config_.has_sub("logging") else Configuration()
        self.debug_logging_ = logging.get("debug", False, bool)
        t = time.perf_counter()
        provider = self.config_.get(id + ".gridprovider", type=str)
        self._info("creating grid with '%s'... " % provider)
        self.grid_ = (MsGridProviders if self.block else GridProviders).create(provider, self.config_, grid_type)
        n = self.grid_.n_cells
        self._info(" done (took %.3gs, has %d element%s)\n" % (time.perf_counter() - t, n, "s" if n > 1 else ""))
        if self.block:  # always AllDirichlet (discreteproblem.hh:318)
            self.boundary_info_ = Configuration("type", "stuff.grid.boundaryinfo.alldirichlet")
        else:
            type = self.config_.get(id + ".boundaryinfo", type=str)
            self.boundary_info_ = self.config_.sub(type) if self.config_.has_sub(type) else Configuration("type", type)
            if "type" not in self.boundary_info_:
                self.boundary_info_["type"] = type
        t = time.perf_counter()
        problem_type = self.config_.get(id + ".problem", type=str)
        self._info("setting up '%s'... " % problem_type)
        self.problem_ = ProblemsProvider.create(problem_type, self.config_, self.grid_)
        self._info("done (took %.3gs)\n" % (time.perf_counter() - t))
        self.visualize_ = logging.get("visualize", True, bool)

    def filename(self):
        return self.filename_

    def config(self):
        return self.config_

    def debug_logging(self):
        return self.debug_logging_

    def grid_provider(self):
        return self.grid_

    def boundary_info(self):
        return self.boundary_info_

    def boundary_types(self):
        return BoundaryInfoProvider.create(self.boundary_info_, self.grid_)

    def problem(self):
        return self.problem_

    def parameters(self):
        """the ``[parameter]`` section: ``<n>.<key> = [values]`` -> [{key: values}, ...] (examples/linearelliptic/swipdg_main.cc:45-53)"""
        out = []
        if self.config_.has_sub("parameter"):
            sub = self.config_.sub("parameter")
            while sub.has_sub(str(len(out))):
                one = sub.sub(str(len(out)))
                out.append({k: one.get(k, type="vector") for k in one.get_value_keys()})
        return out


class DiscreteBlockProblem(DiscreteProblem):
    """discreteproblem.hh:220-440: multiscale grid provider, boundary info fixed to AllDirichlet"""
    block = True


class LinearellipticExampleSWIPDG:
    """examples/linearelliptic/swipdg.hh:24-97"""
    discrete_problem_class = DiscreteProblem

    def __init__(self, grid_type="sgrid", device=0, out=None):
        self.grid_type, self.device, self._out = grid_type, device, out
        self.discrete_problem_ = None
        self.discretization_ = None

    @staticmethod
    def static_id():
        return "linearelliptic.swipdg"

    @classmethod
    def write_config_file(cls, filename=None):
        write_config(filename or cls.static_id() + ".cfg", cls.static_id(), cls.discrete_problem_class.block)

    def _make(self, dp):
        from .discretizations import SWIPDG
        return SWIPDG(dp.grid_provider(), dp.problem(), dp.boundary_types(), device=self.device)

    def initialize(self, arguments=()):
        if self.discrete_problem_ is not None:
            return
        self.discrete_problem_ = self.discrete_problem_class(self.static_id(), arguments, self.grid_type, self._out)
        self.discretization_ = self._make(self.discrete_problem_)
        self.discretization_.init()

    def discrete_problem(self):
        return self.discrete_problem_

    def discretization(self):
        return self.discretization_


class LinearellipticExampleBlockSWIPDG(LinearellipticExampleSWIPDG):
    """examples/linearelliptic/block-swipdg.hh"""
    discrete_problem_class = DiscreteBlockProblem

    @staticmethod
    def static_id():
        return "linearelliptic.block-swipdg"

    def _make(self, dp):
        from .discretizations import BlockSWIPDG
        return BlockSWIPDG(dp.grid_provider(), dp.problem(), device=self.device)
