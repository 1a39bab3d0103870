"""``LinearElliptic::Discretizations::SWIPDG`` / ``BlockSWIPDG`` on top of the C-ABI.

Same method names, argument meaning and error behaviour as the reference classes
(discretizations/interfaces.hh:28-115, base.hh:151-178,240-367, swipdg.hh:159-512, block-swipdg.hh:553-690) and as
their pybindgen projection consumed by pyMOR (examples/linearelliptic/thermalblock_bindings_generator.py:34-58).
Exceptions carry the reference's exception names.
"""
import ctypes as C

import numpy as np

from . import capi
from .capi import HddError


class wrong_input_given(HddError): pass
class you_are_using_this_wrong(HddError): pass
class wrong_parameter_type(HddError): pass
class index_out_of_range(HddError): pass
class NotImplemented_(HddError): pass
class requirements_not_met(HddError): pass
class internal_error(HddError): pass
class device_error(HddError): pass
class linear_solver_failed(HddError): pass


_EXC = {capi.HDD_ERR_WRONG_INPUT: wrong_input_given, capi.HDD_ERR_USING_THIS_WRONG: you_are_using_this_wrong,
        capi.HDD_ERR_WRONG_PARAMETER_TYPE: wrong_parameter_type, capi.HDD_ERR_INDEX_OUT_OF_RANGE: index_out_of_range,
        capi.HDD_ERR_NOT_IMPLEMENTED: NotImplemented_, capi.HDD_ERR_REQUIREMENTS_NOT_MET: requirements_not_met,
        capi.HDD_ERR_INTERNAL: internal_error, capi.HDD_ERR_DEVICE: device_error,
        capi.HDD_ERR_NOT_CONVERGED: linear_solver_failed}


def _check(status):
    if status != capi.HDD_OK:
        raise _EXC.get(status, HddError)(status, capi.lib().hdd_last_error().decode())


def _owned_vector(disc, v, what="vector", writable=False):
    """the reference throws shapes_do_not_match for a vector of the wrong size (wrong_input_given here); the C-ABI copies
    num_owned_dofs() doubles from / to the pointer, so the check has to happen before the call"""
    n = disc.num_owned_dofs()
    if writable:
        if not (isinstance(v, np.ndarray) and v.dtype == np.float64 and v.flags.c_contiguous and v.flags.writeable):
            raise wrong_input_given(capi.HDD_ERR_WRONG_INPUT, "%s has to be a writable C-contiguous float64 array" % what)
        a = v
    else:
        a = capi.as_f64(v)
    if a.ndim != 1 or a.shape[0] != n:
        raise wrong_input_given(capi.HDD_ERR_WRONG_INPUT, "%s has shape %s, the space has %d (owned) DoFs" % (what, a.shape, n))
    return a


def _mu_array(mu):
    if mu is None:
        return None, 0
    if isinstance(mu, dict):
        if len(mu) == 0:
            return None, 0
        mu = list(mu.values())[0]
    a = np.atleast_1d(np.asarray(mu, dtype=np.float64)).copy()
    return a, a.shape[0]


class _Parts:
    """AffinelyDecomposedContainer view: num_components(), component(q), coefficient(q), affine_part()."""

    def __init__(self, disc, which):
        self._d, self._which = disc, which

    def num_components(self):
        n, a = C.c_int(), C.c_int()
        _check(capi.lib().hdd_num_components(self._d._h, self._which, C.byref(n), C.byref(a)))
        return n.value

    def has_affine_part(self):
        n, a = C.c_int(), C.c_int()
        _check(capi.lib().hdd_num_components(self._d._h, self._which, C.byref(n), C.byref(a)))
        return bool(a.value)

    def parametric(self):
        return self.num_components() > 0

    def coefficient(self, q):
        s = C.c_char_p()
        _check(capi.lib().hdd_component_coefficient(self._d._h, self._which, q, C.byref(s)))
        return s.value.decode()

    def _values(self, q):
        p, n = C.POINTER(C.c_double)(), C.c_int64()
        _check(capi.lib().hdd_component_values(self._d._h, self._which, q, C.byref(p), C.byref(n)))
        out = np.empty(n.value)
        _check(capi.lib().hdd_copy_to_host(self._d._h, capi.ptr(out), p, C.c_size_t(out.nbytes)))
        return out

    def component(self, q):
        """values of component q (matrix: CSR value array over pattern(); vector: entries), copied to the host"""
        return self._values(q)

    def affine_part(self):
        return self._values(-1)

    def device_pointer(self, q):
        p, n = C.POINTER(C.c_double)(), C.c_int64()
        _check(capi.lib().hdd_component_values(self._d._h, self._which, q, C.byref(p), C.byref(n)))
        return C.cast(p, C.c_void_p).value, n.value

    def freeze_parameter(self, mu=None):
        """sum_q theta_q(mu) component_q + affine_part, on the host (tests / scipy consumers)"""
        mu_a, ms = _mu_array(mu)
        n = self.num_components()
        theta = np.zeros(max(n, 1))
        _check(capi.lib().hdd_evaluate_coefficients(self._d._h, self._which, capi.ptr(mu_a), ms, capi.ptr(theta)))
        out = self.affine_part().copy() if self.has_affine_part() else 0.0
        for q in range(n):
            out = out + theta[q] * self.component(q)
        return out


class _ProductParts(_Parts):
    """get_product(id): the affinely decomposed product matrix (discretizations/base.hh:281-291)."""

    def __init__(self, disc, pid):
        self._d, self._id = disc, pid.encode()

    def _info(self):
        n, a, v = C.c_int(), C.c_int(), C.c_int()
        _check(capi.lib().hdd_product_num_components(self._d._h, self._id, C.byref(n), C.byref(a), C.byref(v)))
        return n.value, bool(a.value), bool(v.value)

    def num_components(self):
        return self._info()[0]

    def has_affine_part(self):
        return self._info()[1]

    def volume_pattern(self):
        """True: values follow SWIPDG.pattern_volume() (one dense block per cell), False: SWIPDG.pattern()"""
        return self._info()[2]

    def coefficient(self, q):
        s = C.c_char_p()
        _check(capi.lib().hdd_product_coefficient(self._d._h, self._id, q, C.byref(s)))
        return s.value.decode()

    def _values(self, q):
        p, n = C.POINTER(C.c_double)(), C.c_int64()
        _check(capi.lib().hdd_product_values(self._d._h, self._id, q, C.byref(p), C.byref(n)))
        out = np.empty(n.value)
        _check(capi.lib().hdd_copy_to_host(self._d._h, capi.ptr(out), p, C.c_size_t(out.nbytes)))
        return out

    def device_pointer(self, q):
        p, n = C.POINTER(C.c_double)(), C.c_int64()
        _check(capi.lib().hdd_product_values(self._d._h, self._id, q, C.byref(p), C.byref(n)))
        return C.cast(p, C.c_void_p).value, n.value

    def freeze_parameter(self, mu=None):
        mu_a, ms = _mu_array(mu)
        out = self.affine_part().copy() if self.has_affine_part() else 0.0
        for q in range(self.num_components()):
            theta = np.zeros(1)
            _check(capi.lib().hdd_expression_evaluate(self.coefficient(q).encode(),
                                                      (self._d.problem.parameter_name or "mu").encode(),
                                                      capi.ptr(mu_a), ms, capi.ptr(theta)))
            out = out + theta[0] * self.component(q)
        return out

    def apply2(self, u, v, mu=None):
        """u^T P(mu) v on the device"""
        mu_a, ms = _mu_array(mu)
        r = C.c_double()
        u, v = _owned_vector(self._d, u, "u"), _owned_vector(self._d, v, "v")
        _check(capi.lib().hdd_product_apply2(self._d._h, self._id, capi.ptr(mu_a), ms, capi.ptr(u), capi.ptr(v), C.byref(r)))
        return r.value

    def induced_norm(self, u, mu=None):
        return float(np.sqrt(max(self.apply2(u, u, mu), 0.0)))


class SWIPDG:
    """Discretizations::SWIPDG(grid_provider, boundary_info_cfg, problem, level, only_these_products)."""

    def __init__(self, grid, problem, boundary_info=None, polorder=1, device=0, cell_range=None, comm=None,
                 only_these_products=()):
        L = capi.lib()
        self.grid, self.problem = grid, problem
        self.polorder = polorder
        self.n_loc = ((polorder + 1) * (polorder + 2) // 2) if grid.kind == capi.HDD_SIMPLEX2D else (polorder + 1) ** 2
        self._h = None
        self._mesh = C.c_void_p()
        self._cache = {}
        cb, ce = (0, grid.n_cells) if cell_range is None else cell_range
        bt = None if boundary_info is None else np.ascontiguousarray(boundary_info, dtype=np.uint8)
        from .grids import CubeProvider
        if isinstance(grid, CubeProvider) and bt is None:
            # the grid provider's three vectors go to the device, the grid tables are written there
            _check(L.hdd_mesh_create_cube(C.c_int64(grid.nx), C.c_int64(grid.ny), C.c_double(grid.lower_left[0]),
                                          C.c_double(grid.upper_right[0]), C.c_double(grid.lower_left[1]),
                                          C.c_double(grid.upper_right[1]), grid.partitions[0], grid.partitions[1],
                                          C.c_int64(cb), C.c_int64(ce), device, C.byref(self._mesh)))
        else:
            if isinstance(grid, CubeProvider):
                grid = self.grid = grid.materialize()
            _check(L.hdd_mesh_create(grid.kind, C.c_int64(grid.n_cells), C.c_int64(grid.n_verts), capi.ptr(grid.xy),
                                     capi.ptr(grid.cell_verts, C.c_int32), capi.ptr(grid.cell_neigh, C.c_int32),
                                     capi.ptr(grid.cell_subdomain, C.c_int32), capi.ptr(bt, C.c_uint8), C.c_int64(cb),
                                     C.c_int64(ce), device, C.byref(self._mesh)))
        self._comm = comm  # keep the communicator alive as long as the mesh
        if comm is not None:
            _check(L.hdd_mesh_attach_comm(self._mesh, comm.handle))
        self._cproblem = problem.to_c()
        h = C.c_void_p()
        try:
            _check(L.hdd_swipdg_create(self._mesh, polorder, C.byref(self._cproblem), C.byref(h)))
        except Exception:
            L.hdd_mesh_destroy(self._mesh)
            self._mesh = None
            raise
        self._h = h
        self.cell_range = (cb, ce)
        if only_these_products:
            ids = (C.c_char_p * len(only_these_products))(*[p.encode() for p in only_these_products])
            _check(L.hdd_swipdg_only_these_products(self._h, ids, len(only_these_products)))

    def __del__(self):
        try:
            L = capi.lib()
            if self._h:
                L.hdd_swipdg_destroy(self._h)
            if self._mesh:
                L.hdd_mesh_destroy(self._mesh)
        except Exception:
            pass

    @staticmethod
    def static_id():
        return "hdd.linearelliptic.discretizations.swipdg"

    # ---- lifecycle -------------------------------------------------------------------------------------
    def init(self):
        _check(capi.lib().hdd_swipdg_init(self._h))

    def assemble(self):
        """re-runs system_assembler.walk() on the device and returns its device time in seconds"""
        t = C.c_double()
        _check(capi.lib().hdd_swipdg_assemble(self._h, C.byref(t)))
        return t.value

    # ---- space / containers --------------------------------------------------------------------------
    def num_dofs(self):
        g, o = C.c_int64(), C.c_int64()
        _check(capi.lib().hdd_num_dofs(self._h, C.byref(g), C.byref(o)))
        return g.value

    def num_owned_dofs(self):
        g, o = C.c_int64(), C.c_int64()
        _check(capi.lib().hdd_num_dofs(self._h, C.byref(g), C.byref(o)))
        return o.value

    def create_vector(self):
        return np.zeros(self.num_owned_dofs())

    def pattern(self):
        """(rowptr int64, col int32) of the owned rows, global column indices"""
        L = capi.lib()
        n, nnz = C.c_int64(), C.c_int64()
        rp, cl = C.POINTER(C.c_int64)(), C.POINTER(C.c_int32)()
        _check(L.hdd_pattern(self._h, C.byref(n), C.byref(nnz), C.byref(rp), C.byref(cl)))
        rowptr = np.empty(n.value + 1, np.int64)
        col = np.empty(nnz.value, np.int32)
        _check(L.hdd_copy_to_host(self._h, capi.ptr(rowptr, C.c_int64), rp, C.c_size_t(rowptr.nbytes)))
        _check(L.hdd_copy_to_host(self._h, capi.ptr(col, C.c_int32), cl, C.c_size_t(col.nbytes)))
        return rowptr, col

    def system_matrix(self):
        return _Parts(self, capi.HDD_LHS)

    def rhs(self):
        return _Parts(self, capi.HDD_RHS)

    get_operator = system_matrix
    get_rhs = rhs

    # ---- products (discretizations/base.hh:272-291) ---------------------------------------------------
    def available_products(self):
        t, n = C.POINTER(C.c_char_p)(), C.c_int()
        _check(capi.lib().hdd_products_available(self._h, C.byref(t), C.byref(n)))
        return [t[i].decode() for i in range(n.value)]

    def get_product(self, id):
        p = _ProductParts(self, id)
        p._info()  # raises like the reference: no products at all / unknown id
        return p

    def pattern_volume(self):
        """(rowptr, col) of the volume-pattern products: one dense n_loc x n_loc block per owned cell"""
        L = capi.lib()
        n, nnz = C.c_int64(), C.c_int64()
        rp, cl = C.POINTER(C.c_int64)(), C.POINTER(C.c_int32)()
        _check(L.hdd_pattern_volume(self._h, C.byref(n), C.byref(nnz), C.byref(rp), C.byref(cl)))
        rowptr = np.empty(n.value + 1, np.int64)
        col = np.empty(nnz.value, np.int32)
        _check(L.hdd_copy_to_host(self._h, capi.ptr(rowptr, C.c_int64), rp, C.c_size_t(rowptr.nbytes)))
        _check(L.hdd_copy_to_host(self._h, capi.ptr(col, C.c_int32), cl, C.c_size_t(col.nbytes)))
        return rowptr, col

    def error_norms(self, exact, exact_dx, exact_dy, vector=None, order=5, mu=None):
        """{L2, H1_semi, energy} norms of vector - exact (test/linearelliptic-swipdg.hh:267-290), on the device"""
        mu_a, ms = _mu_array(mu)
        out = np.zeros(3)
        v = None if vector is None else _owned_vector(self, vector)
        _check(capi.lib().hdd_error_norms(self._h, capi.ptr(v), exact.encode(), exact_dx.encode(), exact_dy.encode(),
                                          int(order), capi.ptr(mu_a), ms, capi.ptr(out)))
        return {"L2": out[0], "H1_semi": out[1], "energy": out[2]}

    def prolong(self, coarse, vector, father=None):
        """Operators::Prolongation(self.grid_view).apply(coarse function, .) (test/linearelliptic.hh:168-176): the DG
        function ``vector`` of the discretization ``coarse`` evaluated at the Lagrange nodes of this (finer) one, on the
        device.  father: grids.fathers(coarse.grid, self.grid), computed if not given."""
        from . import grids
        if father is None:
            father = grids.fathers(coarse.grid, self.grid)
        father = capi.as_i32(father)[self.cell_range[0]:self.cell_range[1]]
        u = capi.as_f64(vector)
        if u.shape[0] != coarse.num_dofs():
            raise wrong_input_given(capi.HDD_ERR_WRONG_INPUT, "vector has %d entries, the coarse space %d" % (u.shape[0], coarse.num_dofs()))
        out = self.create_vector()
        _check(capi.lib().hdd_prolong(coarse._h, capi.ptr(u), self._h, capi.ptr(np.ascontiguousarray(father), C.c_int32),
                                      capi.ptr(out)))
        return out

    def visualize(self, vector, filename, name="solution", mu=None):
        """visualize(vector, filename, name[, mu]) (discretizations/base.hh:125-147): <filename>.vtu with the DG function
        as point data.  SWIPDG has no "dirichlet" shift vector (the boundary values are imposed weakly), so the vector is
        written as it is.  Needs the whole vector (single process); host code (dune_hdd_b200/vtk.py)."""
        from . import vtk
        if self.cell_range != (0, self.grid.n_cells):
            raise requirements_not_met(capi.HDD_ERR_REQUIREMENTS_NOT_MET, "visualize needs the whole grid on this process")
        grid = self.grid.materialize() if hasattr(self.grid, "materialize") else self.grid
        return vtk.write_vtu(filename, grid, self.polorder, {name: vector})

    def parametric(self):
        return self.problem.parametric()

    def parameter_type(self):
        return self.problem.parameter_type()

    def apply(self, x, mu=None):
        """get_operator().freeze_parameter(mu).apply(x)"""
        mu_a, ms = _mu_array(mu)
        x = _owned_vector(self, x, "x")
        y = np.empty_like(x)
        _check(capi.lib().hdd_apply(self._h, capi.ptr(mu_a), ms, capi.ptr(x), capi.ptr(y)))
        return y

    def residual(self, mu=None, with_floor=False):
        """||rhs(mu) - system_matrix(mu) x|| / ||rhs(mu)|| of the solution the last solve left on the device, recomputed
        there with one SpMV (global over all ranks; collective).  with_floor: also the fp64 rounding level of that
        quantity, 2^-53 (1 + n_faces) n_loc max|A| ||x|| / ||b||"""
        mu_a, ms = _mu_array(mu)
        r, f = C.c_double(), C.c_double()
        _check(capi.lib().hdd_residual(self._h, capi.ptr(mu_a), ms, C.byref(r), C.byref(f)))
        return (r.value, f.value) if with_floor else r.value

    # ---- solve --------------------------------------------------------------------------------------------
    def solver_types(self):
        t, n = C.POINTER(C.c_char_p)(), C.c_int()
        _check(capi.lib().hdd_solver_types(C.byref(t), C.byref(n)))
        return [t[i].decode() for i in range(n.value)]

    def solver_options(self, type=""):
        type = type or self.solver_types()[0]
        if type.split(".lower")[0].split(".upper")[0] not in ["cg", "cg.jacobi", "cg.blockjacobi"] + self.solver_types():
            raise wrong_input_given(capi.HDD_ERR_WRONG_INPUT, "solver type '%s' is not one of solver_types()" % type)
        return {"type": type, "precision": 1e-10, "max_iter": 100000}

    def solve(self, options=None, mu=None, return_info=False):
        """CachedDefault::solve(options, vector, mu): cache lookup on (options, mu), else uncached_solve."""
        if isinstance(options, str):
            options = self.solver_options(options)
        options = dict(self.solver_options() if options is None else options)
        mu_a, ms = _mu_array(mu)
        key = (tuple(sorted(options.items())), None if mu_a is None else tuple(mu_a))
        if key not in self._cache:
            self._cache[key] = self.uncached_solve(options, mu, return_info=True)
        x, info = self._cache[key]
        return (x.copy(), dict(info)) if return_info else x.copy()

    def uncached_solve(self, options=None, mu=None, return_info=False, copy_to_host=True, out=None):
        """out: optional preallocated float64 array of num_owned_dofs() entries (e.g. page-locked, capi.pinned_empty)"""
        options = dict(self.solver_options() if options is None else options)
        mu_a, ms = _mu_array(mu)
        x = (self.create_vector() if out is None else _owned_vector(self, out, "out", writable=True)) if copy_to_host else None
        info = capi.hdd_solve_info()
        _check(capi.lib().hdd_solve(self._h, options.get("type", "").encode(), C.c_double(options.get("precision", 1e-10)),
                                    int(options.get("max_iter", 100000)), capi.ptr(mu_a), ms, capi.ptr(x),
                                    C.byref(info)))
        d = {"iterations": info.iterations, "converged": bool(info.converged),
             "relative_residual": info.relative_residual, "seconds": info.seconds,
             "seconds_per_iteration": info.seconds_per_iteration, "peer_memory": bool(info.peer_memory)}
        return (x, d) if return_info else x

    # ---- estimators ---------------------------------------------------------------------------------------
    def _parameters(self, parameters):
        p = capi.hdd_parameters()
        keep = []
        size = 0
        for key in ("mu", "mu_hat", "mu_bar", "parameter_range_min", "parameter_range_max"):
            if parameters and key in parameters and parameters[key] is not None:
                a, n = _mu_array(parameters[key])
                keep.append(a)
                setattr(p, key, capi.ptr(a))
                size = n
        p.mu_size = size
        p._keep = keep
        return p

    def available_estimators(self):
        t, n = C.POINTER(C.c_char_p)(), C.c_int()
        _check(capi.lib().hdd_estimators_available(self._h, C.byref(t), C.byref(n)))
        return [t[i].decode() for i in range(n.value)]

    def estimate(self, vector, type, parameters=None):
        p = self._parameters(parameters)
        eta = C.c_double()
        v = None if vector is None else _owned_vector(self, vector)
        _check(capi.lib().hdd_estimate(self._h, type.encode(), capi.ptr(v), C.byref(p), C.byref(eta), None))
        return eta.value

    def estimate_local(self, vector, type, parameters=None):
        p = self._parameters(parameters)
        eta = C.c_double()
        v = None if vector is None else _owned_vector(self, vector)
        n = self.num_subdomains() if "OS2014" in type else self.num_owned_dofs() // self.n_loc
        out = np.zeros(n)
        _check(capi.lib().hdd_estimate(self._h, type.encode(), capi.ptr(v), C.byref(p), C.byref(eta), capi.ptr(out)))
        return out

    def indicators(self, vector, parameters=None):
        """all squared per-cell indicators of one device pass (dict of arrays), for parity tests"""
        p = self._parameters(parameters)
        v = None if vector is None else _owned_vector(self, vector)
        n = self.num_owned_dofs() // self.n_loc
        out = np.zeros((8, n))
        _check(capi.lib().hdd_indicators(self._h, capi.ptr(v), C.byref(p), capi.ptr(out)))
        names = ["nc2", "res2", "r2", "df2", "dfstar2", "rstar2", "amin", "resstar2"]
        return {k: out[i] for i, k in enumerate(names)}

    # ---- BlockSWIPDG views (also valid on a plain SWIPDG: one subdomain) ----------------------------------
    def num_subdomains(self):
        n = C.c_int()
        _check(capi.lib().hdd_num_subdomains(self._h, C.byref(n)))
        return n.value

    def subdomain_offsets(self):
        p = C.POINTER(C.c_int64)()
        _check(capi.lib().hdd_subdomain_offsets(self._h, C.byref(p)))
        return np.array([p[i] for i in range(self.num_subdomains() + 1)], dtype=np.int64)

    def neighbouring_subdomains(self, ss):
        p, n = C.POINTER(C.c_int32)(), C.c_int()
        _check(capi.lib().hdd_neighbouring_subdomains(self._h, ss, C.byref(p), C.byref(n)))
        return [int(p[i]) for i in range(n.value)]

    def _block(self, ss, nn, q):
        import scipy.sparse as sp
        m = capi.hdd_csr()
        _check(capi.lib().hdd_block_extract(self._h, ss, nn, q, C.byref(m)))
        try:
            rowptr = np.ctypeslib.as_array(m.rowptr, shape=(m.n_rows + 1,)).copy()
            col = np.ctypeslib.as_array(m.col, shape=(max(m.nnz, 1),))[:m.nnz].copy()
            val = np.ctypeslib.as_array(m.val, shape=(max(m.nnz, 1),))[:m.nnz].copy()
            return sp.csr_matrix((val, col, rowptr), shape=(m.n_rows, m.n_cols))
        finally:
            capi.lib().hdd_csr_free(C.byref(m))


class BlockSWIPDG(SWIPDG):
    """Discretizations::BlockSWIPDG(ms_grid_provider, cfg, problem): the grid carries cell_subdomain (subdomain-major
    numbering), boundary info is forced to AllDirichlet (discretizations/block-swipdg.hh:110,237)."""

    def __init__(self, grid, problem, polorder=1, device=0, cell_range=None, comm=None, only_these_products=()):
        if grid.cell_subdomain is None:
            raise wrong_input_given(capi.HDD_ERR_WRONG_INPUT, "BlockSWIPDG needs a grid with subdomains")
        super().__init__(grid, problem, None, polorder, device, cell_range, comm, only_these_products)

    @staticmethod
    def static_id():
        return "hdd.linearelliptic.discretizations.block-swipdg"

    def get_local_operator(self, ss, q=-1):
        """diagonal block of affine part q: volume + inner faces of ss + its share of the coupling faces
        (discretizations/block-swipdg.hh:625-632, :376-380)"""
        return self._block(ss, ss, q)

    def get_coupling_operator(self, ss, nn, q=-1):
        """off-diagonal coupling block (ss, nn) (discretizations/block-swipdg.hh:634-660)"""
        return self._block(ss, nn, q)

    def get_local_functional(self, ss, q=-1):
        """rhs part q restricted to subdomain ss (discretizations/block-swipdg.hh:662-669)"""
        off = self.subdomain_offsets()
        if ss < 0 or ss >= len(off) - 1:
            raise index_out_of_range(capi.HDD_ERR_INDEX_OUT_OF_RANGE,
                                     "0 <= ss < num_subdomains() = %d is not true for ss = %d!" % (len(off) - 1, ss))
        part = self.rhs().affine_part() if q == -1 else self.rhs().component(q)
        r0 = self.cell_range[0] * self.n_loc
        return np.array(part[off[ss] - r0:off[ss + 1] - r0])

    def get_local_product(self, ss, id, q=-1):
        """product of the local discretization on subdomain ss (discretizations/block-swipdg.hh:612-618) for the
        volume-pattern products l2 / h1_semi / elliptic, whose local matrix is the diagonal sub-block of the global
        one: scipy CSR with subdomain-local indices"""
        import scipy.sparse as sp
        off = self.subdomain_offsets()
        if ss < 0 or ss >= len(off) - 1:
            raise index_out_of_range(capi.HDD_ERR_INDEX_OUT_OF_RANGE,
                                     "0 <= ss < num_subdomains() = %d is not true for ss = %d!" % (len(off) - 1, ss))
        if id not in ("l2", "h1_semi", "elliptic"):
            raise wrong_input_given(capi.HDD_ERR_WRONG_INPUT, id)
        P = self.get_product(id)
        vals = P.affine_part() if q == -1 else P.component(q)
        nl, r0 = self.n_loc, self.cell_range[0] * self.n_loc
        a, b = off[ss] - r0, off[ss + 1] - r0
        blocks = vals.reshape(-1, nl, nl)[a // nl:b // nl]
        return sp.block_diag(list(blocks), format="csr") if len(blocks) else sp.csr_matrix((0, 0))

    def localize_vector(self, global_vector, ss):
        """discretizations/block-swipdg.hh:567-583"""
        off = self.subdomain_offsets()
        if ss < 0 or ss >= len(off) - 1:
            raise index_out_of_range(capi.HDD_ERR_INDEX_OUT_OF_RANGE,
                                     "0 <= ss < num_subdomains() = %d is not true for ss = %d!" % (len(off) - 1, ss))
        if global_vector.shape[0] != off[-1]:
            raise index_out_of_range(capi.HDD_ERR_INDEX_OUT_OF_RANGE, "The size() of global_vector does not match!")
        return np.array(global_vector[off[ss]:off[ss + 1]])

    def globalize_vectors(self, local_vectors):
        """discretizations/block-swipdg.hh:585-600"""
        off = self.subdomain_offsets()
        if len(local_vectors) != len(off) - 1:
            raise index_out_of_range(capi.HDD_ERR_INDEX_OUT_OF_RANGE, "wrong number of local vectors")
        for ss, v in enumerate(local_vectors):
            if v.shape[0] != off[ss + 1] - off[ss]:
                raise index_out_of_range(capi.HDD_ERR_INDEX_OUT_OF_RANGE, "local vector %d has the wrong size" % ss)
        return np.concatenate(local_vectors)
