"""VTK output of discontinuous Lagrange functions: what ``discretization.visualize(vector, filename, name)``
(discretizations/base.hh:125-147 -> GDT::ConstDiscreteFunction::visualize -> Dune::VTKWriter, upstream) leaves on disk
for ParaView - an UnstructuredGrid ``<filename>.vtu`` with point data.  Host code.

A DG function jumps across faces, so every cell carries its own copy of its nodes (Dune's VTK::nonconforming mode).
polOrder 1 writes linear triangles / quads, polOrder 2 quadratic triangles / biquadratic quads, which represent the
P2 / Q2 function exactly (Dune's writer would subsample instead)."""
import numpy as np

# reference nodes in DoF order (device.cuh: Elem<KIND, P>) and the VTK node order / cell type of each element
_SIMPLEX, _CUBE = 0, 1
_NODES = {
    (_SIMPLEX, 1): [(0, 0), (1, 0), (0, 1)],
    (_SIMPLEX, 2): [(0, 0), (.5, 0), (1, 0), (0, .5), (.5, .5), (0, 1)],
    (_CUBE, 1): [(0, 0), (1, 0), (0, 1), (1, 1)],
    (_CUBE, 2): [(i / 2, j / 2) for j in range(3) for i in range(3)],
}
_VTK = {  # (cell type, permutation: VTK node k = DoF perm[k])
    (_SIMPLEX, 1): (5, [0, 1, 2]),
    (_SIMPLEX, 2): (22, [0, 2, 5, 1, 4, 3]),              # corners, then the midpoints of edges 01, 12, 20
    (_CUBE, 1): (9, [0, 1, 3, 2]),                        # VTK_QUAD runs counter-clockwise
    (_CUBE, 2): (28, [0, 2, 8, 6, 1, 5, 7, 3, 4]),        # corners, edge midpoints (bottom, right, top, left), centre
}


def node_coordinates(grid, polorder):
    """[n_cells, n_local, 2] physical coordinates of the Lagrange nodes"""
    ref = np.array(_NODES[(grid.kind, polorder)], dtype=np.float64)
    v = grid.xy[grid.cell_verts]
    if grid.kind == _SIMPLEX:
        return v[:, None, 0] + ref[None, :, 0, None] * (v[:, None, 1] - v[:, None, 0]) + ref[None, :, 1, None] * (v[:, None, 2] - v[:, None, 0])
    return v[:, None, 0] + ref[None, :, :] * (v[:, None, 3] - v[:, None, 0])


def _array(name, type, data, components=None):
    comp = "" if components is None else ' NumberOfComponents="%d"' % components
    fmt = "%d" if type.startswith(("Int", "UInt")) else "%.17g"
    body = " ".join(fmt % x for x in np.asarray(data).reshape(-1))
    return '<DataArray type="%s" Name="%s"%s format="ascii">\n%s\n</DataArray>\n' % (type, name, comp, body)


def write_vtu(filename, grid, polorder, point_data=None, cell_data=None):
    """point_data: {name: DG vector (n_cells * n_local)}; cell_data: {name: one value per cell}.  -> the file name"""
    if not filename.endswith(".vtu"):
        filename += ".vtu"
    cell_type, perm = _VTK[(grid.kind, polorder)]
    nl, nc = len(perm), grid.n_cells
    xyz = np.zeros((nc, nl, 3))
    xyz[:, :, :2] = node_coordinates(grid, polorder)[:, perm]
    out = ['<?xml version="1.0"?>\n<VTKFile type="UnstructuredGrid" version="0.1" byte_order="LittleEndian">\n<UnstructuredGrid>\n',
           '<Piece NumberOfPoints="%d" NumberOfCells="%d">\n' % (nc * nl, nc)]
    names = list((point_data or {}).keys())
    out.append('<PointData%s>\n' % (' Scalars="%s"' % names[0] if names else ""))
    for name, vec in (point_data or {}).items():
        vec = np.asarray(vec, dtype=np.float64)
        if vec.shape != (nc * nl,):
            raise ValueError("'%s' has %s entries, the space %d" % (name, vec.shape, nc * nl))
        out.append(_array(name, "Float64", vec.reshape(nc, nl)[:, perm]))
    out.append("</PointData>\n")
    names = list((cell_data or {}).keys())
    out.append('<CellData%s>\n' % (' Scalars="%s"' % names[0] if names else ""))
    for name, vec in (cell_data or {}).items():
        vec = np.asarray(vec, dtype=np.float64)
        if vec.shape != (nc,):
            raise ValueError("'%s' has %s entries, the grid %d cells" % (name, vec.shape, nc))
        out.append(_array(name, "Float64", vec))
    out.append("</CellData>\n<Points>\n" + _array("Coordinates", "Float64", xyz, 3) + "</Points>\n<Cells>\n")
    out.append(_array("connectivity", "Int64", np.arange(nc * nl)))
    out.append(_array("offsets", "Int64", nl * np.arange(1, nc + 1)))
    out.append(_array("types", "UInt8", np.full(nc, cell_type)))
    out.append("</Cells>\n</Piece>\n</UnstructuredGrid>\n</VTKFile>\n")
    with open(filename, "w") as f:
        f.write("".join(out))
    return filename
