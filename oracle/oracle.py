"""ctypes binding of the CPU oracle (oracle/swipdg_oracle.cpp).

TEST INFRASTRUCTURE - only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs may import this module.  The product never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

SIMPLEX, CUBE = 0, 1
FN_ONE, FN_CELLWISE, FN_ESV_FORCE, FN_OS_SIN, FN_ESV_EXACT, FN_X, FN_Y, FN_XY = range(8)


class OFn(C.Structure):
    _fields_ = [("n", C.c_int), ("order", C.c_int), ("kind", C.c_int * 4), ("coef", C.c_double * 4),
                ("cell", C.POINTER(C.c_double) * 4)]


def fn(terms, order=0):
    """terms: list of (coef, kind[, cellwise ndarray])."""
    f = OFn()
    f.n = len(terms)
    f.order = order
    f._keep = []
    for k, t in enumerate(terms):
        f.coef[k] = float(t[0])
        f.kind[k] = int(t[1])
        if len(t) > 2:
            arr = np.ascontiguousarray(t[2], dtype=np.float64)
            f._keep.append(arr)
            f.cell[k] = arr.ctypes.data_as(C.POINTER(C.c_double))
    return f


def const(c):
    return fn([(c, FN_ONE)], 0)


def cellwise(values):
    return fn([(1.0, FN_CELLWISE, values)], 0)


def esv2007_force():
    """problems/ESV2007.hh:78 with integration_order 3 (testcases/ESV2007.hh:64)."""
    return fn([(1.0, FN_ESV_FORCE)], 3)


def esv2007_exact():
    return fn([(1.0, FN_ESV_EXACT)], 2)


def os2014_factor(mu):
    """problems/OS2014.hh:63-74: [1 + 0.75 s] + mu * [-0.75 s], s = sin(4 pi (x + y/2)), order 3."""
    return fn([(1.0, FN_ONE), (0.75 * (1.0 - mu), FN_OS_SIN)], 3)


def os2014_affine():
    return fn([(1.0, FN_ONE), (0.75, FN_OS_SIN)], 3)


def os2014_component():
    return fn([(-0.75, FN_OS_SIN)], 3)


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "swipdg_oracle.cpp")
    if force or not os.path.exists(so) or (os.path.exists(src) and os.path.getmtime(src) > os.path.getmtime(so)):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.or_fn_eval.restype = C.c_double
        _LIB.or_fn_eval.argtypes = [C.POINTER(OFn), C.c_int, C.c_double, C.c_double]
    return _LIB


def set_threads(n):
    """host threads of the assembly walk / SpMV / CG (default 1 = the reference's serial walk; more only for the
    'all cores' CPU baseline of bench.py)"""
    lib().or_set_threads(int(n))


def _p(a, t=C.c_double):
    if a is None:
        return None
    return a.ctypes.data_as(C.POINTER(t))


def _pf(f):
    return None if f is None else C.byref(f)


def n_local(kind, polorder):
    return (polorder + 1) * (polorder + 2) // 2 if kind == SIMPLEX else (polorder + 1) ** 2


class Mesh:
    def __init__(self, kind, xy, cv, nb, polorder=1):
        self.kind = kind
        self.p = polorder
        self.xy = np.ascontiguousarray(xy, dtype=np.float64)
        self.cv = np.ascontiguousarray(cv, dtype=np.int32)
        self.nb = np.ascontiguousarray(nb, dtype=np.int32)
        self.nl = n_local(kind, polorder)
        self.nc = self.cv.shape[0]
        self.nv = self.xy.shape[0]

    def with_polorder(self, polorder):
        return Mesh(self.kind, self.xy, self.cv, self.nb, polorder)

    @property
    def n_dofs(self):
        return self.nl * self.nc

    def args(self):
        return (self.kind, self.p, self.nc, self.nv, _p(self.xy), _p(self.cv, C.c_int32), _p(self.nb, C.c_int32))


def mesh_cube(nx, ny, x0, x1, y0, y1):
    xy = np.empty(((nx + 1) * (ny + 1), 2))
    cv = np.empty((nx * ny, 4), np.int32)
    nb = np.empty((nx * ny, 4), np.int32)
    lib().or_mesh_cube(nx, ny, C.c_double(x0), C.c_double(x1), C.c_double(y0), C.c_double(y1), _p(xy),
                       _p(cv, C.c_int32), _p(nb, C.c_int32))
    return Mesh(CUBE, xy, cv, nb)


def mesh_bisect(n, x0, x1, bisections):
    nc = 2 * n * n * 2 ** bisections
    maxv = 4 * nc + 16
    xy = np.empty((maxv, 2))
    cv = np.empty((nc, 3), np.int32)
    nb = np.empty((nc, 3), np.int32)
    nv = lib().or_mesh_bisect(n, C.c_double(x0), C.c_double(x1), bisections, _p(xy), _p(cv, C.c_int32),
                              _p(nb, C.c_int32), maxv)
    assert nv > 0
    return Mesh(SIMPLEX, xy[:nv].copy(), cv, nb)


def line_rule(order):
    x = np.empty(64); w = np.empty(64)
    n = lib().or_line_rule(order, _p(x), _p(w))
    return x[:n].copy(), w[:n].copy()


def element_rule(kind, order):
    x = np.empty(1024); y = np.empty(1024); w = np.empty(1024)
    n = lib().or_element_rule(kind, order, _p(x), _p(y), _p(w))
    return x[:n].copy(), y[:n].copy(), w[:n].copy()


def pattern(mesh):
    n = mesh.n_dofs
    rowptr = np.empty(n + 1, np.int64)
    lib().or_pattern(mesh.kind, mesh.p, mesh.nc, _p(mesh.nb, C.c_int32), _p(rowptr, C.c_int64), None)
    col = np.empty(rowptr[-1], np.int32)
    lib().or_pattern(mesh.kind, mesh.p, mesh.nc, _p(mesh.nb, C.c_int32), _p(rowptr, C.c_int64), _p(col, C.c_int32))
    return rowptr, col


def pattern_volume(mesh):
    """block-diagonal pattern of the volume-only products (l2, h1_semi, elliptic, boundary_l2)"""
    n = mesh.n_dofs
    rowptr = np.empty(n + 1, np.int64)
    col = np.empty(n * mesh.nl, np.int32)
    lib().or_pattern_volume(mesh.kind, mesh.p, mesh.nc, _p(rowptr, C.c_int64), _p(col, C.c_int32))
    return rowptr, col


PRODUCTS = {"l2": 0, "h1_semi": 1, "elliptic": 2, "boundary_l2": 3, "penalty": 4}


def assemble_product(mesh, which, rowptr, col, factor=None, tensor=None, bnd_dirichlet=None):
    val = np.zeros(col.shape[0])
    t = None if tensor is None else np.ascontiguousarray(tensor, dtype=np.float64)
    bd = None if bnd_dirichlet is None else np.ascontiguousarray(bnd_dirichlet, dtype=np.uint8)
    factor = factor or const(1.0)
    lib().or_assemble_product(*mesh.args(), PRODUCTS[which], _pf(factor), _p(t), _p(bd, C.c_uint8),
                              _p(rowptr, C.c_int64), _p(col, C.c_int32), _p(val))
    return val


def assemble_lhs(mesh, factor, tensor, rowptr, col, bnd_dirichlet=None):
    val = np.zeros(col.shape[0])
    t = None if tensor is None else np.ascontiguousarray(tensor, dtype=np.float64)
    bd = None if bnd_dirichlet is None else np.ascontiguousarray(bnd_dirichlet, dtype=np.uint8)
    lib().or_assemble_lhs(*mesh.args(), _pf(factor), _p(t), _p(bd, C.c_uint8), _p(rowptr, C.c_int64),
                          _p(col, C.c_int32), _p(val))
    return val


def assemble_rhs(mesh, force, factor=None, dirichlet=None, tensor=None, neumann=None, bnd_type=None):
    """bnd_type: uint8 [nc, nf], 1 Dirichlet / 2 Neumann; None = AllDirichlet"""
    b = np.zeros(mesh.n_dofs)
    t = None if tensor is None else np.ascontiguousarray(tensor, dtype=np.float64)
    bt = None if bnd_type is None else np.ascontiguousarray(bnd_type, dtype=np.uint8)
    lib().or_assemble_rhs(*mesh.args(), _pf(force), _pf(factor), _pf(dirichlet), _pf(neumann), _p(t),
                          _p(bt, C.c_uint8), _p(b))
    return b


def spmv(rowptr, col, val, x):
    y = np.empty_like(x)
    lib().or_spmv(C.c_int64(x.shape[0]), _p(rowptr, C.c_int64), _p(col, C.c_int32), _p(val), _p(x), _p(y))
    return y


def cg(rowptr, col, val, b, precond=1, rtol=1e-10, maxit=100000, x0=None, history=False):
    n = b.shape[0]
    x = np.zeros(n) if x0 is None else np.array(x0, dtype=np.float64)
    relres = C.c_double(0.0)
    hist = np.zeros(maxit + 1) if history else None
    lib().or_cg.restype = C.c_int
    it = lib().or_cg(C.c_int64(n), _p(rowptr, C.c_int64), _p(col, C.c_int32), _p(val), _p(b), _p(x), precond,
                     C.c_double(rtol), maxit, C.byref(relres), _p(hist))
    if history:
        return x, it, relres.value, hist[:it + 1]
    return x, it, relres.value


def oswald(mesh, u):
    iu = np.empty_like(u)
    lib().or_oswald(mesh.kind, mesh.nc, mesh.nv, _p(mesh.cv, C.c_int32), _p(mesh.nb, C.c_int32), _p(u), _p(iu))
    return iu


def indicators(mesh, u, force, a_mu, a_hat=None, a_bar=None, a_cut=None, a_min=None, a_max=None, tensor=None):
    """Per-cell squared indicators (dict of ndarrays), see or_indicators."""
    a_hat = a_hat or a_mu
    a_bar = a_bar or a_mu
    a_cut = a_cut or a_mu
    a_min = a_min or a_mu
    a_max = a_max or a_mu
    names = ["nc2", "res2", "r2", "df2", "dfstar2", "rstar2", "amin", "resstar2"]
    out = {k: np.zeros(mesh.nc) for k in names}
    t = None if tensor is None else np.ascontiguousarray(tensor, dtype=np.float64)
    u = np.ascontiguousarray(u, dtype=np.float64)
    lib().or_indicators(mesh.nc, mesh.nv, _p(mesh.xy), _p(mesh.cv, C.c_int32), _p(mesh.nb, C.c_int32), _p(u),
                        _pf(a_mu), _pf(a_hat), _pf(a_bar), _pf(a_cut), _pf(a_min), _pf(a_max), _p(t), _pf(force),
                        *[_p(out[k]) for k in names])
    return out


def error_norms(mesh, u, exact, factor=None, tensor=None, order=5):
    out = np.zeros(3)
    t = None if tensor is None else np.ascontiguousarray(tensor, dtype=np.float64)
    u = np.ascontiguousarray(u, dtype=np.float64)
    lib().or_error_norms(mesh.kind, mesh.p, mesh.nc, mesh.nv, _p(mesh.xy), _p(mesh.cv, C.c_int32), _p(u), _pf(exact),
                         _pf(factor), _p(t), order, _p(out))
    return {"L2": out[0], "H1_semi": out[1], "energy": out[2]}


def to_scipy(rowptr, col, val):
    import scipy.sparse as sp
    n = rowptr.shape[0] - 1
    return sp.csr_matrix((val, col, rowptr), shape=(n, n))


# ---- prolongation of the convergence studies (test/linearelliptic.hh:168-176) ----------------------------------
def reference_nodes(kind, p):
    """Lagrange nodes in DoF order: vertices for p = 1; p = 2 lexicographic ((0,0),(1/2,0),(1,0),(0,1/2),(1/2,1/2),(0,1)
    on the triangle, i + 3 j on the square), as or_basis in swipdg_oracle.cpp."""
    if kind == SIMPLEX:
        pts = [(i / p, j / p) for j in range(p + 1) for i in range(p + 1 - j)]
    else:
        pts = [(i / p, j / p) for j in range(p + 1) for i in range(p + 1)]
    return np.array(pts)


def basis_values(kind, p, xi, eta):
    """nodal Lagrange basis at reference points (arrays) -> [n_points, n_local]"""
    xi, eta = np.asarray(xi, dtype=np.float64), np.asarray(eta, dtype=np.float64)
    if kind == SIMPLEX:
        l0, l1, l2 = 1.0 - xi - eta, xi, eta
        if p == 1:
            return np.stack([l0, l1, l2], axis=-1)
        return np.stack([l0 * (2 * l0 - 1), 4 * l0 * l1, l1 * (2 * l1 - 1), 4 * l0 * l2, 4 * l1 * l2, l2 * (2 * l2 - 1)], axis=-1)
    if p == 1:
        lx, ly = [1.0 - xi, xi], [1.0 - eta, eta]
    else:
        lag = lambda t: [(1 - t) * (1 - 2 * t), 4 * t * (1 - t), t * (2 * t - 1)]
        lx, ly = lag(xi), lag(eta)
    return np.stack([lx[i] * ly[j] for j in range(p + 1) for i in range(p + 1)], axis=-1)


def _local_coordinates(mesh, cells, pts):
    """reference coordinates of pts[k] in cell cells[k]"""
    v = mesh.xy[mesh.cv[cells]]
    if mesh.kind == SIMPLEX:
        e1, e2, d = v[:, 1] - v[:, 0], v[:, 2] - v[:, 0], pts - v[:, 0]
        det = e1[:, 0] * e2[:, 1] - e2[:, 0] * e1[:, 1]
        return (d[:, 0] * e2[:, 1] - e2[:, 0] * d[:, 1]) / det, (e1[:, 0] * d[:, 1] - d[:, 0] * e1[:, 1]) / det
    h = v[:, 3] - v[:, 0]
    return (pts[:, 0] - v[:, 0, 0]) / h[:, 0], (pts[:, 1] - v[:, 0, 1]) / h[:, 1]


def fathers(coarse, fine):
    """for every fine cell the coarse cell containing its centre (the reference: ALUGrid father() /
    Stuff::Grid::EntityInlevelSearch, test/linearelliptic-swipdg.hh:186-194, -block-swipdg.hh:169-177): candidates from
    a k-d tree over the coarse centres, containment by local coordinates; independent of the product's bucket search"""
    from scipy.spatial import cKDTree
    cc = coarse.xy[coarse.cv].mean(axis=1)
    cf = fine.xy[fine.cv].mean(axis=1)
    k = min(12, coarse.nc)
    _, cand = cKDTree(cc).query(cf, k=k)
    cand = cand.reshape(fine.nc, k)
    best, depth = np.full(fine.nc, -1, np.int64), np.full(fine.nc, -1e-9)
    for j in range(k):
        xi, eta = _local_coordinates(coarse, cand[:, j], cf)
        d = np.minimum(np.minimum(xi, eta), 1 - xi - eta) if coarse.kind == SIMPLEX else \
            np.minimum(np.minimum(xi, 1 - xi), np.minimum(eta, 1 - eta))
        better = d > depth
        best[better], depth[better] = cand[better, j], d[better]
    assert (best >= 0).all(), "a fine cell centre lies in no coarse cell"
    return best


def prolong(coarse, u, fine, father=None):
    """fine DG vector = the coarse DG function at the fine Lagrange nodes (GDT::Operators::Prolongation)"""
    father = fathers(coarse, fine) if father is None else np.asarray(father)
    nodes = reference_nodes(fine.kind, fine.p)
    v = fine.xy[fine.cv]
    out = np.empty((fine.nc, fine.nl))
    uc = np.asarray(u).reshape(coarse.nc, coarse.nl)[father]
    for i, (xi, eta) in enumerate(nodes):
        if fine.kind == SIMPLEX:
            pts = v[:, 0] + xi * (v[:, 1] - v[:, 0]) + eta * (v[:, 2] - v[:, 0])
        else:
            pts = v[:, 0] + np.array([xi, eta]) * (v[:, 3] - v[:, 0])
        a, b = _local_coordinates(coarse, father, pts)
        out[:, i] = (basis_values(coarse.kind, coarse.p, a, b) * uc).sum(axis=1)
    return out.reshape(-1)
