"""The configuration-file driver and the VTK writer (discreteproblem.hh:44-440,
examples/linearelliptic/swipdg_main.cc, discretizations/base.hh:125-147) - host code, no device needed."""
import os
import subprocess
import sys
import xml.etree.ElementTree as ET

import numpy as np
import pytest

from dune_hdd_b200 import discreteproblem as dp
from dune_hdd_b200 import grids, problems, vtk

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_configuration_tree():
    cfg = dp.Configuration.from_string("""
        top = 1            # comment
        [a]
        flag = true
        vec = [0.5 1 2]
        [a.b]
        name = some text
        [c]
        n = 7
    """)
    assert cfg.get("top", type=int) == 1 and cfg.get("a.flag", False) is True and cfg.get("a.vec", type="vector") == [0.5, 1.0, 2.0]
    assert cfg.has_sub("a") and cfg.has_sub("a.b") and not cfg.has_sub("top") and cfg.sub("a").get("b.name") == "some text"
    assert cfg.sub("a").get_value_keys() == ["flag", "vec"] and cfg.get("c.missing", 3) == 3 and cfg.get("c.n", 0.5) == 7.0
    with pytest.raises(dp.configuration_error):
        cfg.get("c.missing")
    with pytest.raises(dp.configuration_error):
        cfg.sub("nope")
    with pytest.raises(dp.configuration_error):
        cfg.get("a.b.name", type=int)
    again = dp.Configuration.from_string(cfg.report())
    assert again._d == cfg._d
    both = dp.Configuration({"x": [1, 2]})
    both.add(cfg.sub("a"), "sub")
    assert both["x"] == "[1 2]" and both.get("sub.b.name") == "some text"


def test_default_config_file_and_discrete_problem(tmp_path):
    """write_config (discreteproblem.hh:63-84) and the constructor (:86-152)"""
    cls = dp.LinearellipticExampleSWIPDG
    assert cls.static_id() == "linearelliptic.swipdg"  # examples/linearelliptic/swipdg.hh:35-38
    fn = tmp_path / (cls.static_id() + ".cfg")
    cls.write_config_file(str(fn))
    text = fn.read_text()
    for needle in ("[linearelliptic.swipdg]", "gridprovider = stuff.grid.provider.cube", "boundaryinfo = ", "problem = ",
                   "[logging]", "info  = true", "debug = true", "file  = false", "visualize = true", "[parameter]",
                   "0.diffusion_factor = [0.1 0.1 1.0 1.0]", "1.diffusion_factor = [1.0 1.0 0.1 0.1]",
                   "[hdd.linearelliptic.problem.ESV2007]", "integration_order = 3"):
        assert needle in text, needle
    log = []
    p = dp.DiscreteProblem(cls.static_id(), [str(tmp_path)], out=log.append)
    assert p.grid_provider().n_cells == 64 and p.grid_provider().kind == grids.CUBE2D and p.problem().name == "ESV2007"
    assert p.filename() == cls.static_id() and p.debug_logging() and p.boundary_types() is None
    assert p.boundary_info().get("type") == "stuff.grid.boundaryinfo.alldirichlet"
    assert p.parameters() == [{"diffusion_factor": [0.1, 0.1, 1.0, 1.0]}, {"diffusion_factor": [1.0, 1.0, 0.1, 0.1]}]
    assert "has 64 elements" in "".join(log)
    (tmp_path / "some.other.id.cfg").write_text(text)  # the file exists but has no [some.other.id] (discreteproblem.hh:99-101)
    with pytest.raises(dp.configuration_error, match="Missing sub 'some.other.id'"):
        dp.DiscreteProblem("some.other.id", [str(tmp_path)])


def _edit(path, **replacements):
    text = path.read_text()
    for old, new in replacements.values():
        assert old in text, old
        text = text.replace(old, new)
    path.write_text(text)


def test_problems_and_grids_from_the_config(tmp_path):
    cls = dp.LinearellipticExampleSWIPDG
    fn = tmp_path / (cls.static_id() + ".cfg")
    cls.write_config_file(str(fn))
    _edit(fn, problem=("problem = hdd.linearelliptic.problem.ESV2007", "problem = hdd.linearelliptic.problem.OS2014.parametricESV2007"),
          ll=("[stuff.grid.provider.cube]\nlower_left = [0.0 0.0]", "[stuff.grid.provider.cube]\nlower_left = [-1 -1]"),
          ne=("num_elements = [8 8]", "num_elements = [4 4]"), nr=("num_refinements = 0", "num_refinements = 1"),
          bi=("boundaryinfo = stuff.grid.boundaryinfo.alldirichlet\n               stuff.grid.boundaryinfo.allneumann",
              "boundaryinfo = stuff.grid.boundaryinfo.allneumann"))
    p = dp.DiscreteProblem(cls.static_id(), [str(tmp_path)], grid_type="alu")
    g = p.grid_provider()
    assert g.kind == grids.SIMPLEX2D and g.n_cells == 128 and g.xy.min() == -1.0 and g.xy.max() == 1.0  # config 1, level 0
    assert p.problem().parametric() and p.problem().parameter_type() == {"mu": 1}
    assert (p.boundary_types() == 2).all() and p.boundary_types().shape == g.cell_neigh.shape
    q = dp.DiscreteProblem(cls.static_id(), [str(tmp_path)], grid_type="sgrid")
    assert q.grid_provider().n_cells == 64  # 4 x 4 refined once
    # thermalblock: the checkerboard indicators are a partition of unity, one coefficient per block
    _edit(fn, problem=("problem = hdd.linearelliptic.problem.OS2014.parametricESV2007", "problem = hdd.linearelliptic.problem.thermalblock"),
          ll=("lower_left = [-1 -1]", "lower_left = [0 0]"))
    t = dp.DiscreteProblem(cls.static_id(), [str(tmp_path)]).problem()
    assert t.parameter_type() == {"diffusion_factor": 4} and t.diffusion_factor.coefficients[3] == "diffusion_factor[3]"
    ind = np.array([c.cell_values for c in t.diffusion_factor.components])
    assert ind.shape == (4, 64) and (ind.sum(axis=0) == 1.0).all() and (ind.sum(axis=1) == 16.0).all()
    assert not t.diffusion_factor.has_affine_part() and t.force.affine.value == 1.0
    # spe10: the permeability file is read when it exists, the synthetic field stands in otherwise
    perm = np.linspace(1.0, 2.0, 2000)
    (tmp_path / "perm.dat").write_text("\n".join(" ".join("%.17g" % v for v in perm[i:i + 6]) for i in range(0, 2000, 6)))
    _edit(fn, problem=("problem = hdd.linearelliptic.problem.thermalblock", "problem = hdd.linearelliptic.problem.spe10.model1"),
          f=("filename = perm_case1.dat", "filename = %s" % (tmp_path / "perm.dat")), ur=("[stuff.grid.provider.cube]\nlower_left = [0 0]\nupper_right = [1.0 1.0]",
             "[stuff.grid.provider.cube]\nlower_left = [0 0]\nupper_right = [5 1]"), ne=("num_elements = [4 4]", "num_elements = [100 20]"),
          nr=("num_refinements = 1", "num_refinements = 0"))
    s = dp.DiscreteProblem(cls.static_id(), [str(tmp_path)])
    assert s.grid_provider().n_cells == 2000 and np.array_equal(s.problem().diffusion_tensor[:, 0], perm)
    assert np.array_equal(problems.read_spe10_model1(str(tmp_path / "perm.dat")).reshape(-1), perm)


def test_block_problem(tmp_path):
    cls = dp.LinearellipticExampleBlockSWIPDG
    assert cls.static_id() == "linearelliptic.block-swipdg"  # examples/linearelliptic/block-swipdg.hh:28-31
    fn = tmp_path / (cls.static_id() + ".cfg")
    cls.write_config_file(str(fn))
    text = fn.read_text()
    assert "gridprovider = grid.multiscale.provider.cube" in text and "boundaryinfo" not in text and "debug = false" in text
    _edit(fn, parts=("num_partitions = [2 2]", "num_partitions = [4 2]"))
    p = dp.DiscreteBlockProblem(cls.static_id(), [str(tmp_path)])
    g = p.grid_provider()
    assert g.n_subdomains == 8 and (np.diff(g.cell_subdomain) >= 0).all() and p.boundary_types() is None
    assert p.parameters() == []


@pytest.mark.parametrize("kind", ["alu", "sgrid"])
@pytest.mark.parametrize("polorder", [1, 2])
def test_vtu_writer(tmp_path, kind, polorder):
    g = grids.simplex(2) if kind == "alu" else grids.cube(3, 2, (0.0, 0.0), (3.0, 1.0))
    X = vtk.node_coordinates(g, polorder)
    f = lambda x, y: 1 + 2 * x - y + (x * y - 0.5 * y * y if polorder == 2 else 0)
    u = f(X[..., 0], X[..., 1]).reshape(-1)
    name = vtk.write_vtu(str(tmp_path / "out"), g, polorder, {"u": u}, {"subdomain": np.arange(g.n_cells)})
    assert name.endswith("out.vtu")
    piece = ET.parse(name).getroot().find("UnstructuredGrid/Piece")
    nl = X.shape[1]
    assert int(piece.attrib["NumberOfCells"]) == g.n_cells and int(piece.attrib["NumberOfPoints"]) == g.n_cells * nl
    pts = np.array(piece.find("Points/DataArray").text.split(), float).reshape(g.n_cells, nl, 3)
    val = np.array(piece.find("PointData/DataArray").text.split(), float).reshape(g.n_cells, nl)
    assert np.abs(val - f(pts[..., 0], pts[..., 1])).max() < 1e-14 and (pts[..., 2] == 0).all()
    arrays = {a.attrib["Name"]: np.array(a.text.split(), float) for a in piece.find("Cells").findall("DataArray")}
    assert np.array_equal(arrays["connectivity"], np.arange(g.n_cells * nl)) and np.array_equal(arrays["offsets"], nl * np.arange(1, g.n_cells + 1))
    assert set(arrays["types"]) == {{("alu", 1): 5, ("alu", 2): 22, ("sgrid", 1): 9, ("sgrid", 2): 28}[(kind, polorder)]}
    # VTK node order: corners counter-clockwise, then (quadratic cells) the edge midpoints in edge order, then the centre
    nv = 3 if kind == "alu" else 4
    c = pts[:, :nv, :2]
    area = 0.5 * sum(c[:, i, 0] * c[:, (i + 1) % nv, 1] - c[:, (i + 1) % nv, 0] * c[:, i, 1] for i in range(nv))
    assert (np.abs(area) > 0).all() and (np.sign(area) == np.sign(area[0])).all()
    if kind == "sgrid":
        assert (area > 0).all()
    if polorder == 2:
        for e in range(nv):
            assert np.abs(pts[:, nv + e, :2] - 0.5 * (c[:, e] + c[:, (e + 1) % nv])).max() < 1e-15
        if kind == "sgrid":
            assert np.abs(pts[:, 8, :2] - c.mean(axis=1)).max() < 1e-15
    cell = np.array(piece.find("CellData/DataArray").text.split(), float)
    assert np.array_equal(cell, np.arange(g.n_cells))
    with pytest.raises(ValueError):
        vtk.write_vtu(str(tmp_path / "bad"), g, polorder, {"u": u[:-1]})


def test_example_driver_writes_its_config_and_needs_a_device(tmp_path):
    """examples/swipdg_main.py = examples/linearelliptic/swipdg_main.cc: first run writes the config, second run runs"""
    script = os.path.join(ROOT, "examples", "swipdg_main.py")
    r = subprocess.run([sys.executable, script, str(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 0 and "Please review the configuration and start me again!" in r.stdout
    assert (tmp_path / "linearelliptic.swipdg.cfg").exists()
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    r = subprocess.run([sys.executable, script, str(tmp_path)], capture_output=True, text=True)
    assert r.returncode != 0 and "no CUDA device" in r.stderr and "has 64 elements" in r.stdout
