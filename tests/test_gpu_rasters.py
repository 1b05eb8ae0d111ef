"""GPU parity of the raster ingest kernels against the oracle restatement of
`/root/reference/ext/HydroModelsRastersExt`: D8 flow directions bit-exact (integer codes), packed forcing bit-exact
(a pure gather), and a DEM -> flow directions -> grid model run end to end."""
import os
import sys

import numpy as np
import pytest

import hydromodels_jl_b200 as hm
from hydromodels_jl_b200 import rasters as rs
from hydromodels_jl_b200 import synthetic as sy
from hydromodels_jl_b200.engine import DeviceBuffer

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "oracle"))
import hydro_oracle as ho  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("R,Cc,cell", [(5, 5, (1.0, 1.0)), (37, 70, (30.0, 30.0)), (64, 33, (1.0, 2.5)), (1, 1, (1.0, 1.0)),
                                       (9, 1, (2.0, 1.0)), (130, 257, (90.0, 60.0))])
def test_d8_flow_direction_matches_oracle_bit_for_bit(R, Cc, cell):
    rng = np.random.default_rng(R * 1000 + Cc)
    if R == 5:
        dem = np.array([[10.0 - 0.5 * (i + j) for j in range(5)] for i in range(5)])      # test_basic.jl:27-33
    else:
        dem = np.round(rng.normal(500, 50, (R, Cc)), 1)                                     # rounded: plenty of exact ties
        dem[rng.random((R, Cc)) < 0.05] = np.nan
        if R > 20:
            dem[10:14, 5:20] = 480.0                                                        # a flat lake
    got = rs.compute_d8_flow_direction(dem, cell)
    ref = ho.compute_d8_flow_direction(dem, cell)
    assert got.dtype == np.int64 and np.array_equal(got, ref)
    if R == 5:
        assert np.all(got[:4, :4] == 2) and got[4, 4] == 16


def test_d8_large_tilted_plane_property():
    """4096 x 4096 (BASELINE config 5's grid): a plane falling to the south-east drains SE everywhere it can."""
    n = 4096
    i = np.arange(n, dtype=np.float64)
    dem = 1e4 - 0.5 * (i[:, None] + i[None, :])
    f = rs.compute_d8_flow_direction(dem)
    assert np.all(f[:-1, :-1] == 2) and np.all(f[:-1, -1] == 4) and np.all(f[-1, :-1] == 1) and f[-1, -1] == 16


@pytest.mark.parametrize("layout", ["trc", "tcr"])
def test_pack_forcing_is_an_exact_gather(layout):
    rng = np.random.default_rng(3)
    R, Cc, T, V = 23, 41, 17, 3
    cubes = [rng.normal(size=(R, Cc, T)) for _ in range(V)]                                  # reference indexing [row, col, time]
    pos = [(r + 1, c + 1) for r in range(R) for c in range(Cc) if rng.random() < 0.7]
    mem = [np.ascontiguousarray(a.transpose(2, 0, 1)) if layout == "trc" else np.ascontiguousarray(a.transpose(2, 1, 0)) for a in cubes]
    dev = [DeviceBuffer.from_host(m) for m in mem]
    out = rs.device_forcing(dev, pos, T, R, Cc, layout=layout)
    got = out.to_host((V, len(pos), T), order="F")                                           # memory [T][N][V]
    assert np.array_equal(got, ho.gather_forcing(cubes, pos))
    with pytest.raises(IndexError):
        rs.device_forcing(dev, [(R + 1, 1)], T, R, Cc, layout=layout)


@pytest.mark.parametrize("layout", ["trc", "tcr"])
@pytest.mark.parametrize("R,Cc", [(70, 45), (32, 64), (5, 3)])
def test_pack_forcing_dense_grid_takes_the_tiled_kernel(layout, R, Cc):
    """Every cell in row-major order -> `hb_pack_forcing_dense_device` (32 x 32 tiles through shared memory); ragged edge
    tiles, both raster memory orders, and equality with the position-list kernel."""
    rng = np.random.default_rng(R * Cc)
    T, V = 9, 3
    cubes = [rng.normal(size=(R, Cc, T)) for _ in range(V)]
    pos = [(r + 1, c + 1) for r in range(R) for c in range(Cc)]
    mem = [np.ascontiguousarray(a.transpose(2, 0, 1)) if layout == "trc" else np.ascontiguousarray(a.transpose(2, 1, 0)) for a in cubes]
    dev = [DeviceBuffer.from_host(m) for m in mem]
    got = rs.device_forcing(dev, pos, T, R, Cc, layout=layout).to_host((V, R * Cc, T), order="F")
    assert np.array_equal(got, ho.gather_forcing(cubes, pos))
    if R * Cc > 1:                                                   # the same cells in another order: position-list kernel
        perm = rng.permutation(R * Cc)
        got2 = rs.device_forcing(dev, [pos[i] for i in perm], T, R, Cc, layout=layout).to_host((V, R * Cc, T), order="F")
        assert np.array_equal(got2, got[:, perm, :])


@pytest.mark.parametrize("layout", ["trc", "tcr"])
@pytest.mark.parametrize("dense", [True, False])
def test_streamed_ingest_from_host_rasters(layout, dense, monkeypatch):
    """`hb_ingest_forcing`: host rasters -> `[T][N][V]` on the device through two chunk slots (copy of chunk c + 1 under the
    gather of chunk c).  Ordinary numpy arrays (pageable: pinned staging by the library's threads) and pinned ones, a chunk
    size that cuts T into many ragged chunks and the one-chunk default — all equal to the exact gather."""
    from hydromodels_jl_b200.engine import check, lib
    rng = np.random.default_rng(11)
    R, Cc, T, V = 37, 50, 23, 3
    cubes = [rng.normal(size=(R, Cc, T)) for _ in range(V)]                                  # reference indexing [row, col, time]
    pos = [(r + 1, c + 1) for r in range(R) for c in range(Cc) if dense or rng.random() < 0.6]
    mem = [np.ascontiguousarray(a.transpose(2, 0, 1)) if layout == "trc" else np.asfortranarray(a) for a in cubes]
    ref = ho.gather_forcing(cubes, pos)
    for chunk in (None, str(V * R * Cc * 8 * 4), str(1)):                                    # default | 4 steps per chunk | 1 step
        if chunk is None:
            monkeypatch.delenv("HB_INGEST_CHUNK_BYTES", raising=False)
        else:
            monkeypatch.setenv("HB_INGEST_CHUNK_BYTES", chunk)
        out = rs.ingest_forcing(mem, None if dense else pos, layout=layout)
        assert np.array_equal(out.to_host((V, len(pos), T), order="F"), ref)
        out.free()
    # pinned sources take the direct path (no staging)
    import ctypes as C
    n_bytes = mem[0].nbytes
    pinned, keep = [], []
    for m in mem:
        ptr = C.c_void_p()
        check(lib().hb_malloc_host(C.byref(ptr), n_bytes))
        keep.append(ptr)
        a = np.ctypeslib.as_array((C.c_double * m.size).from_address(ptr.value)).reshape(m.shape, order="C" if layout == "trc" else "F")
        a[...] = m
        pinned.append(a)
    monkeypatch.setenv("HB_INGEST_CHUNK_BYTES", str(V * R * Cc * 8 * 5))
    out = rs.ingest_forcing(pinned, None if dense else pos, layout=layout)
    assert np.array_equal(out.to_host((V, len(pos), T), order="F"), ref)
    out.free()
    for ptr in keep:
        check(lib().hb_free_host(ptr))
    with pytest.raises(ValueError):
        rs.ingest_forcing([m.transpose(0, 2, 1) for m in mem], layout=layout)                # not contiguous in the stated order
    with pytest.raises(IndexError):
        rs.ingest_forcing(mem, [(R + 1, 1)], layout=layout)


def test_dem_to_routed_grid_run():
    """DEM -> D8 on the GPU -> `exphydro_grid` on those flow directions: same result as with the oracle's directions,
    and the outlet of the plane collects more flow than the ridge (`test/base/run_spatial_model.jl:140`)."""
    R, Cc, T = 24, 32, 40
    i, j = np.arange(R, dtype=np.float64), np.arange(Cc, dtype=np.float64)
    dem = 100.0 - 0.5 * i[:, None] - 0.3 * j[None, :]
    info = rs.create_hydroroute_from_dem(dem, (1.0, 1.0))
    assert np.array_equal(info["flow_dir"], ho.compute_d8_flow_direction(dem))
    N = len(info["positions"])
    model = hm.models.exphydro_grid(info["flow_dir"], np.asarray(info["positions"]))
    p = {"params": dict(sy.parameters("exphydro", 0, N), area_coef=np.ones(N), lag=np.full(N, 0.5))}
    init = dict(snowpack=np.zeros(N), soilwater=np.full(N, 1303.0), s_river=np.zeros(N))
    inp = sy.forcing("exphydro", 0, N, T)
    out = model(inp, p, initstates=init)
    ref = ho.run_model(model, inp, p, initstates=init)
    assert np.allclose(out, ref, rtol=1e-10, atol=1e-12)
    q = out[12].mean(axis=1).reshape(R, Cc)
    assert q[-1, -1] > q[0, 0]
