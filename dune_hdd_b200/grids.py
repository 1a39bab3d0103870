"""Grid providers: host-side stand-ins for ``Stuff::Grid::Providers::Cube<GridType>`` and the multiscale grid
(``grid::Multiscale``) the reference's test cases use (testcases/ESV2007.hh:123-163, testcases/spe10.hh:262-307).

A ``Grid`` is just the flat arrays the C-ABI takes; a DUNE-equipped host would fill them from its own grid view.
"""
import ctypes as C

import numpy as np

from . import capi

SIMPLEX2D, CUBE2D = capi.HDD_SIMPLEX2D, capi.HDD_CUBE2D


class Grid:
    def __init__(self, kind, xy, cell_verts, cell_neigh, cell_subdomain=None, partitions=(1, 1)):
        self.kind = kind
        # np.ascontiguousarray keeps arrays that already have the right type and layout (e.g. page-locked ones) as they are
        self.xy = capi.as_f64(xy)
        self.cell_verts = capi.as_i32(cell_verts)
        self.cell_neigh = capi.as_i32(cell_neigh)
        self.cell_subdomain = None if cell_subdomain is None else capi.as_i32(cell_subdomain)
        self.partitions = tuple(partitions)
        self.n_loc = 3 if kind == SIMPLEX2D else 4

    @property
    def n_cells(self):
        return self.cell_verts.shape[0]

    @property
    def n_verts(self):
        return self.xy.shape[0]

    @property
    def n_dofs(self):
        return self.n_loc * self.n_cells

    @property
    def n_subdomains(self):
        return 1 if self.cell_subdomain is None else int(self.cell_subdomain.max()) + 1

    def centers(self):
        return self.xy[self.cell_verts].mean(axis=1)

    def subdomain_cell_offsets(self):
        if self.cell_subdomain is None:
            return np.array([0, self.n_cells], dtype=np.int64)
        counts = np.bincount(self.cell_subdomain, minlength=self.n_subdomains)
        return np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)


class CubeProvider:
    """``Stuff::Grid::Providers::Cube(lower_left, upper_right, num_elements)`` + the ``[px py 1]`` multiscale partition as a
    description only (testcases/ESV2007.hh:123-127, :150-163): the discretization hands the three vectors to
    ``hdd_mesh_create_cube`` and the grid tables are written on the device.  Same grid and numbering as ``cube(...)``;
    ``materialize()`` gives the flat host arrays (for VTK output or the oracle)."""
    kind = CUBE2D
    n_loc = 4

    def __init__(self, nx, ny=None, lower_left=(-1.0, -1.0), upper_right=(1.0, 1.0), partitions=(1, 1)):
        self.nx, self.ny = int(nx), int(nx if ny is None else ny)
        self.lower_left, self.upper_right = tuple(map(float, lower_left)), tuple(map(float, upper_right))
        self.partitions = (int(partitions[0]), int(partitions[1]))
        self._grid = None

    n_cells = property(lambda self: self.nx * self.ny)
    n_verts = property(lambda self: (self.nx + 1) * (self.ny + 1))
    n_dofs = property(lambda self: 4 * self.nx * self.ny)
    n_subdomains = property(lambda self: self.partitions[0] * self.partitions[1])
    cell_subdomain = property(lambda self: True)  # "has subdomains" for BlockSWIPDG; the array lives in materialize()

    @staticmethod
    def _starts(n, parts):
        b = np.clip(((np.arange(n) + 0.5) / n * parts).astype(np.int64), 0, parts - 1)
        return np.concatenate([np.searchsorted(b, np.arange(parts)), [n]]).astype(np.int64)

    def subdomain_cell_offsets(self):
        X, Y = self._starts(self.nx, self.partitions[0]), self._starts(self.ny, self.partitions[1])
        sizes = (np.diff(Y)[:, None] * np.diff(X)[None, :]).ravel()
        return np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)

    def materialize(self):
        if self._grid is None:
            self._grid = cube(self.nx, self.ny, self.lower_left, self.upper_right, self.partitions)
        return self._grid


def cube(nx, ny=None, lower_left=(-1.0, -1.0), upper_right=(1.0, 1.0), partitions=(1, 1), pinned=False):
    """SGrid<2,2> via Providers::Cube(lower_left, upper_right, num_elements): nx*ny axis-parallel cells, x fastest.
    pinned: allocate the arrays in page-locked host memory (hdd_host_alloc) so that the upload runs at link speed."""
    ny = nx if ny is None else ny
    L = capi.lib()
    nc, nv = C.c_int64(), C.c_int64()
    capi.check(L.hdd_grid_cube_sizes(C.c_int64(nx), C.c_int64(ny), C.byref(nc), C.byref(nv)))
    empty = capi.pinned_empty if pinned else np.empty
    xy = empty((nv.value, 2), np.float64)
    cv = empty((nc.value, 4), np.int32)
    nb = empty((nc.value, 4), np.int32)
    sub = empty((nc.value,), np.int32)
    capi.check(L.hdd_grid_cube(C.c_int64(nx), C.c_int64(ny), C.c_double(lower_left[0]), C.c_double(upper_right[0]),
                               C.c_double(lower_left[1]), C.c_double(upper_right[1]), int(partitions[0]),
                               int(partitions[1]), capi.ptr(xy), capi.ptr(cv, C.c_int32), capi.ptr(nb, C.c_int32),
                               capi.ptr(sub, C.c_int32)))
    return Grid(CUBE2D, xy, cv, nb, sub, partitions)


def simplex(squares_per_side, lower_left=(-1.0, -1.0), upper_right=(1.0, 1.0), partitions=(1, 1)):
    """ALUGrid<2,2,simplex,conforming> ladder member: squares_per_side^2 squares of 8 right triangles each.

    The ESV2007 ladder (testcases/ESV2007.hh:50-59,123-134; testcases/base.hh:92-103) is Cube(-1,1,4) refined by
    2 + 2*level bisections, i.e. ``squares_per_side = 4 * 2**level`` (128, 512, 2048, 8192 cells; reference 32768).
    """
    L = capi.lib()
    s = int(squares_per_side)
    nc, nv = C.c_int64(), C.c_int64()
    capi.check(L.hdd_grid_simplex_sizes(C.c_int64(s), C.byref(nc), C.byref(nv)))
    xy = np.empty((nv.value, 2))
    cv = np.empty((nc.value, 3), np.int32)
    nb = np.empty((nc.value, 3), np.int32)
    sub = np.empty(nc.value, np.int32)
    capi.check(L.hdd_grid_simplex(C.c_int64(s), C.c_double(lower_left[0]), C.c_double(upper_right[0]),
                                  C.c_double(lower_left[1]), C.c_double(upper_right[1]), int(partitions[0]),
                                  int(partitions[1]), capi.ptr(xy), capi.ptr(cv, C.c_int32), capi.ptr(nb, C.c_int32),
                                  capi.ptr(sub, C.c_int32)))
    return Grid(SIMPLEX2D, xy, cv, nb, sub, partitions)


def fathers(coarse, fine):
    """father[k] = the cell of ``coarse`` containing the centre of cell k of ``fine`` (hdd_grid_fathers): the hierarchy
    information ALUGrid's ``father()`` / ``Stuff::Grid::EntityInlevelSearch`` give the reference's studies
    (test/linearelliptic-swipdg.hh:186-194, test/linearelliptic-block-swipdg.hh:169-177)."""
    if coarse.kind != fine.kind:
        raise ValueError("coarse and fine grid have different element types")
    out = np.empty(fine.n_cells, np.int32)
    capi.check(capi.lib().hdd_grid_fathers(coarse.kind, C.c_int64(coarse.n_cells), C.c_int64(coarse.n_verts),
                                           capi.ptr(coarse.xy), capi.ptr(coarse.cell_verts, C.c_int32),
                                           C.c_int64(fine.n_cells), C.c_int64(fine.n_verts), capi.ptr(fine.xy),
                                           capi.ptr(fine.cell_verts, C.c_int32), capi.ptr(out, C.c_int32)))
    return out
