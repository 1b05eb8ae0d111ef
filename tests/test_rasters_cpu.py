"""CPU-only: the oracle restatement of the reference's Rasters extension (`compute_d8_flow_direction`,
`HydroModelsRastersExt.jl:32-109`) against hand-worked cases, and the host-side helpers (positions, NetCDF
loaders, `prepare_netcdf_forcing`) against files written here."""
import os
import sys

import numpy as np
import pytest

import hydromodels_jl_b200 as hm
from hydromodels_jl_b200 import rasters as rs

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "oracle"))
import hydro_oracle as ho  # noqa: E402


def test_d8_oracle_on_the_tilted_plane_of_the_reference_script():
    """`ext/HydroModelsRastersExt/test_basic.jl:27-33`: elevation falls 0.5 per cell to the east and to the south, so
    the south-east neighbour is steepest (1/sqrt 2 > 0.5) wherever it exists; the last column drains south, the last
    row east; the outlet corner is a pit whose least-uphill neighbours are W and N (-0.5 each; NW is -0.707) — W comes
    first in the reference's order and a later neighbour only wins on a strictly greater slope."""
    dem = np.array([[10.0 - 0.5 * (i + j) for j in range(5)] for i in range(5)])
    f = ho.compute_d8_flow_direction(dem)
    assert np.all(f[:4, :4] == 2) and np.all(f[:4, 4] == 4) and np.all(f[4, :4] == 1) and f[4, 4] == 16
    assert set(np.unique(f)) <= {0, 1, 2, 4, 8, 16, 32, 64, 128}


def test_d8_oracle_nodata_cellsize_and_ties():
    nan = np.nan
    dem = np.array([[5.0, nan, 5.0], [5.0, 4.0, nan], [nan, 3.0, 9.0]])
    f = ho.compute_d8_flow_direction(dem)
    assert f[0, 1] == 0 and f[1, 2] == 0 and f[2, 0] == 0               # NoData cells
    assert f[1, 1] == 4                                                 # the only lower neighbour is south
    assert f[0, 0] == 2 and f[0, 2] == 8 and f[2, 2] == 16              # NaN neighbours are skipped
    # anisotropic cells: orthogonal distance is max(cx, cy) for BOTH axes (reference :93-97)
    dem = np.array([[3.0, 2.0], [1.9, 0.0]])
    assert ho.compute_d8_flow_direction(dem, (1.0, 1.0))[0, 0] == 2     # 3/sqrt2 = 2.12 > 1.1
    assert ho.compute_d8_flow_direction(dem, (1.0, 3.0))[0, 0] == 2     # 3/sqrt10 = 0.95 > 1.1/3
    # a flat DEM: every slope is 0, the first existing neighbour in E, SE, S, ... order wins
    f = ho.compute_d8_flow_direction(np.zeros((3, 3)))
    assert f.tolist() == [[1, 1, 4], [1, 1, 4], [1, 1, 16]]
    assert ho.compute_d8_flow_direction(np.zeros((1, 1)))[0, 0] == 0    # no neighbour at all


def test_positions_and_route_from_flow_directions():
    dem = np.array([[3.0, np.nan, 2.0], [2.0, 1.0, 0.5]])
    assert rs.extract_positions_from_raster(dem) == [(1, 1), (1, 3), (2, 1), (2, 2), (2, 3)]
    assert rs.extract_positions_from_raster(dem, lambda x: x > 1.5) == [(1, 1), (1, 3), (2, 1)]
    flw = ho.compute_d8_flow_direction(dem)
    aggr = hm.build_aggr_func(flw, rs.extract_positions_from_raster(dem))
    down = aggr.downstream()                                            # node each node drains into (-1: leaves the node set)
    assert down.shape == (5,) and np.all(down < 5) and down[4] != 4


def _write_nc(path, arrays):
    from scipy.io import netcdf_file
    with netcdf_file(path, "w") as f:
        R, Cc, T = next(iter(arrays.values())).shape
        f.createDimension("row", R)
        f.createDimension("col", Cc)
        f.createDimension("time", T)
        for nm, a in arrays.items():
            v = f.createVariable(nm, "d", ("row", "col", "time"))
            v[:] = a


def test_netcdf_loaders_and_forcing_assembly(tmp_path):
    rng = np.random.default_rng(0)
    R, Cc, T = 4, 5, 7
    prcp, temp = rng.uniform(0, 10, (R, Cc, T)), rng.normal(5, 5, (R, Cc, T))
    temp[0, 0, :] = np.nan
    path = str(tmp_path / "forcing.nc")
    _write_nc(path, {"pr": prcp, "tas": temp})
    pos = rs.extract_positions_from_netcdf(path, "tas", threshold=-1e9)
    assert (1, 1) not in pos and len(pos) == R * Cc - int(np.sum(~(temp[:, :, 0] > -1e9)))
    data = rs.load_netcdf_data(path, {"prcp": "pr", "temp": "tas"}, pos)
    assert data["prcp"].shape == (T, len(pos))
    r, c = pos[3]
    assert np.array_equal(data["temp"][:, 3], temp[r - 1, c - 1, :])
    forcing = rs.prepare_netcdf_forcing(data, ["temp", "prcp"])                 # reference layout [T, V, Npos]
    assert forcing.shape == (T, 2, len(pos)) and np.array_equal(forcing[:, 1, :], data["prcp"])
    inp = rs.model_input(data, ["temp", "prcp"])                                # functor layout [V, N, T]
    assert inp.flags.f_contiguous and np.array_equal(inp, ho.gather_forcing([temp, prcp], pos))
    simple = rs.load_netcdf_data_simple(path, {"prcp": "pr"})
    assert np.allclose(simple["prcp"], prcp.mean(axis=(0, 1)))
    with pytest.raises(KeyError, match="not found in NetCDF file"):
        rs.load_netcdf_data(path, {"x": "nope"}, pos)
    with pytest.raises(KeyError):
        rs.prepare_netcdf_forcing(data, ["temp", "lday"])
